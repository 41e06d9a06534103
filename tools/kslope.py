"""Per-k-block cost of the GEMM mainloop: time vs K at fixed tile count (slope = time per k-block, intercept = fixed)."""
import os, torch
from rawaudiovae_kelsey_b200 import ops
dev="cuda"
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/10*1e3
tag=f"CG={os.environ.get('RVAE_CTA_GROUP','auto')}"
# wgrad (MN-major operands): dW[1024,2048] = dy[B,1024]^T x[B,2048], split-K 2 -> 128 units (CG1) / 64 pair units (CG2)
prev=None
for B in (8192,16384,32768,65536):
    dy=torch.randn(B,1024,device=dev).to(torch.bfloat16); x=torch.randn(B,2048,device=dev).to(torch.bfloat16)
    out=torch.zeros(1024,2048,device=dev)
    t=timeit(lambda: ops.wgrad(dy,x,out=out,k_splits=2))
    kb=B//64//2
    s="" if prev is None else f"  slope {(t-prev[0])/(kb-prev[1])*1e3:.0f} ns/k-block"
    print(f"{tag} wgrad K={B:6d} ({kb:4d} k-blocks/unit): {t:7.1f} us  {2*B*1024*2048/t/1e6:7.1f} TFLOP/s{s}")
    prev=(t,kb)
# forward (K-major): y[M=18944,256] ... use M = 148*128 rows x N=256 -> exactly 148 tiles (CG1) / 74 pair tiles, vary K
prev=None
for K in (2048,4096,8192,16384):
    M=148*128
    x=torch.randn(M,K,device=dev).to(torch.bfloat16); w=torch.randn(256,K,device=dev).to(torch.bfloat16)
    t=timeit(lambda: ops.linear_act_fwd(x,w,None,ops.ACT_NONE))
    kb=K//64
    s="" if prev is None else f"  slope {(t-prev[0])/(kb-prev[1])*1e3:.0f} ns/k-block"
    print(f"{tag} linear K={K:6d} ({kb:4d} k-blocks/tile): {t:7.1f} us  {2*M*256*K/t/1e6:7.1f} TFLOP/s{s}")
    prev=(t,kb)
