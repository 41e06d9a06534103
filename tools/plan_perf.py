"""Per-GEMM device timing of one fused training step (default.ini dims, B=8192) through the plan (prepared GEMMs,
CUDA events recorded in C around each launch) - no Python/ctypes/tensor-map-encode overhead in the numbers."""
import os
import sys
import torch
from rawvae.model import VAE, FusedTrainStep
from rawaudiovae_kelsey_b200.optim import Adam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
S, H, L = 1024, 2048, 256
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = VAE(S, H, L).to(dev)
opt = Adam(model.parameters(), lr=1e-4)
step = FusedTrainStep(model, opt, 1e-4, graph=os.environ.get('STEP_GRAPH', '0') == '1')
PIPE = os.environ.get('STEP_PIPE', '0') == '1'
if PIPE:   # frames gathered from a device-resident corpus; the next batch is prefetched by the current step
    from rawvae.model import FrameBatch
    audio = torch.rand(32 * 30 * 44100, device=dev) * 2 - 1
    nfr = (audio.numel() - S) // 128 + 1
    idx = torch.randint(0, nfr, (64, B), device=dev)
    fbs = [FrameBatch(audio, B, 128, S, frame_idx=idx[i]) for i in range(64)]
    k = [0]
    _step = step
    def step(_x=None):
        i = k[0] % 64
        k[0] += 1
        return _step(fbs[i], next_data=fbs[(i + 1) % 64])
    x = None
else:
    x = torch.rand(B, S, device=dev) * 2 - 1
for _ in range(12):
    step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
e0.record()
for _ in range(n):
    step(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
tag = " ".join(f"{k}={os.environ[k]}" for k in ("RVAE_CTA_GROUP", "RVAE_BLOCK_N", "RVAE_DEBUG", "STEP_PIPE", "STEP_GRAPH") if k in os.environ)
print(f"[{tag}] step {ms*1e3:.1f} us  -> {B/ms/1e3:.2f} M frames/s  ({30408704*B/ms/1e9:.0f} TFLOP/s whole step)")
plan = model._plan_for(B)
plan.enable_timing(True)
for _ in range(20):
    step(x)
torch.cuda.synchronize()
tm = plan.read_timing()
tot, aux = 0.0, 0.0
for k, (t, c, f) in tm.items():
    if not c:
        continue
    us = 1e3 * t / 20          # per step
    if f > 0:
        tot += us
        print(f"  {k:8s} {us:7.1f} us  {f*c/20/us/1e6:7.1f} TFLOP/s")
    else:
        aux += us
        print(f"  {k:8s} {us:7.1f} us  ({c//20} launches/step)")
print(f"  GEMM total {tot:.1f} us -> {30408704*B/max(tot,1e-9)/1e6:.0f} TFLOP/s chain; other kernels {aux:.1f} us; "
      f"untimed step {ms*1e3:.1f} us")
# pure host cost of enqueueing a step (few enough launches not to fill the launch queue)
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"  host enqueue {1e5*(t1-t0):.1f} us/step")
