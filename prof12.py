import math, torch
from snfft_b200 import get_plan
n=11
plan=get_plan(n); nf=math.factorial(n)
x=torch.randn(1,nf,dtype=torch.float64,device='cuda'); spec=torch.empty_like(x); work=torch.empty_like(x)
for _ in range(2): plan.forward(x,'coset',out=spec,work=work)
torch.cuda.synchronize()
