import torch
from rawaudiovae_kelsey_b200 import ops
dev="cuda"; n=5772800
p,g,m,v = (torch.randn(n,device=dev) for _ in range(4)); v.abs_()
st=torch.zeros((),device=dev); sh=torch.empty(n,dtype=torch.bfloat16,device=dev)
def t(fn,name,bytes_):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f"{name:28s} {ms*1e3:7.1f} us  {bytes_/ms/1e6:7.1f} GB/s")
t(lambda: ops.adam_step(p,g,m,v,st,1e-4,shadow_hi=sh,increment_step=False),"adam+shadow",n*30)
t(lambda: ops.adam_step(p,g,m,v,st,1e-4,increment_step=False),"adam",n*28)
pp=p.clone().requires_grad_(); pp.grad=g.clone()
o=torch.optim.Adam([pp],lr=1e-4,fused=True)
t(lambda: o.step(),"torch fused adam",n*28)
a=torch.empty(n*4,device=dev); b=torch.empty(n*4,device=dev)
t(lambda: b.copy_(a),"copy 92MB",n*32)
t(lambda: ops.split_bf16(p, False),"split_bf16 (alloc+kernel)",n*6)
t(lambda: ops.randn((8192,256),1,0),"randn 8MB",8192*256*4)
x=torch.randn(8192,2048,device=dev).to(torch.bfloat16)
t(lambda: ops.colsum(x),"colsum 8192x2048",8192*2048*2)
