# quick GPU sanity check (scratch; superseded by tests/)
import sys, math, time
sys.path.insert(0,'oracle')
import numpy as np, torch
import snfft_oracle as O
from snfft_b200 import Plan
dev=0
for n in range(1,9):
    plan=Plan(n,dev)
    rng=np.random.default_rng(n); B=3
    f=rng.standard_normal((B,math.factorial(n)))
    ref=O.pack(O.fft_all(f),n)
    ft=torch.from_numpy(f).cuda()
    got=plan.forward(ft,'lex').cpu().numpy()
    e1=np.abs(got-ref).max()/np.abs(ref).max()
    back=plan.inverse(torch.from_numpy(ref).cuda(),'lex').cpu().numpy()
    e2=np.abs(back-f).max()
    # level-by-level with streaming kernel
    c2l=O.coset_major_to_lex(n)
    x=torch.from_numpy(f[:,c2l]).cuda()
    for k in range(2,n+1): x=plan.level(k,x)
    e3=np.abs(x.cpu().numpy()-ref).max()/np.abs(ref).max()
    y=torch.from_numpy(ref).cuda()
    for k in range(n,1,-1): y=plan.level(k,y,inverse=True)
    e4=np.abs(y.cpu().numpy()-f[:,c2l]).max()
    print(n,'fwd',e1,'inv',e2,'lvl fwd',e3,'lvl inv',e4, flush=True)
n=8; plan=Plan(n,dev); B=1024
f=torch.randn(B,math.factorial(n),dtype=torch.float64,device='cuda')
out=torch.empty_like(f); work=torch.empty_like(f); back=torch.empty_like(f)
for it in range(3):
    torch.cuda.synchronize(); t=time.time()
    plan.forward(f,'coset',out=out,work=work); torch.cuda.synchronize(); t1=time.time()
    plan.inverse(out,'coset',out=back,work=work); torch.cuda.synchronize(); t2=time.time()
    print('S8x1024 fwd %.3f ms inv %.3f ms'%((t1-t)*1e3,(t2-t1)*1e3), 'roundtrip', (back-f).abs().max().item())
