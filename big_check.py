import math, time, torch, sys
from snfft_b200 import get_plan, timing_enable, timing_read
for n in (9,10,11,12):
    t0=time.time(); plan=get_plan(n); tb=time.time()-t0
    nf=math.factorial(n); B = 8 if n<=10 else 1
    x=torch.randn(B,nf,dtype=torch.float64,device='cuda')
    spec=torch.empty_like(x); work=torch.empty_like(x); back=torch.empty_like(x)
    plan.forward(x,'coset',out=spec,work=work); plan.inverse(spec,'coset',out=back,work=work)
    torch.cuda.synchronize()
    err=((back-x).norm()/x.norm()).item()
    parts,dims,offs=plan.layout()
    lhs=sum(d*spec[:,o:o+d*d].pow(2).sum(dim=1) for d,o in zip(dims,offs)); rhs=nf*x.pow(2).sum(dim=1)
    pl=((lhs-rhs).abs()/rhs).max().item()
    timing_enable(True); timing_read()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.forward(x,'coset',out=spec,work=work); plan.inverse(spec,'coset',out=back,work=work)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/3
    kinds=timing_read(); timing_enable(False)
    print('n=%d B=%d plan %.1fs tables %.1f MB  roundtrip %.2e plancherel %.2e  step %.3f ms  %.2f Gel/s  frac %.3f'%(n,B,tb,plan.device_bytes()/1e6,err,pl,ms,2*B*nf/ms/1e6, 32*B*nf/ms/1e6/6540.5), {k:(round(v[0]/3,3),v[1]//3) for k,v in kinds.items() if v[1]}, flush=True)
    del x,spec,work,back
