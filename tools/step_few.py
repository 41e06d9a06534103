"""A handful of fused training steps (default.ini dims, B=8192) for ncu launch lists."""
import sys
import time
import torch
from rawvae.model import VAE, FusedTrainStep
from rawaudiovae_kelsey_b200.optim import Adam
from rawaudiovae_kelsey_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B, S, H, L = 8192, 1024, 2048, 256
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = VAE(S, H, L).to(dev)
opt = Adam(model.parameters(), lr=1e-4)
step = FusedTrainStep(model, opt, 1e-4)
x = torch.rand(B, S, device=dev) * 2 - 1
torch.cuda.synchronize()
l0 = ops.launch_count()
t0 = time.perf_counter()
for _ in range(n):
    loss = step(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6*(t1-t0)/n:.1f} us/step, wall {1e6*(t2-t0)/n:.1f} us/step, launches/step {(ops.launch_count()-l0)/n:.1f}, loss {float(loss):.5f}")
