"""torchrun --nproc-per-node W tools/dp_perf.py : DP step timing + host enqueue cost."""
import os, sys, time
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rawaudiovae_kelsey_b200 import dist as rdist
from rawaudiovae_kelsey_b200.optim import Adam
from rawvae.model import VAE, FusedTrainStep
rank, world, local = rdist.init_from_env("nccl")
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
S, H, L, B = 1024, 2048, 256, 8192
torch.manual_seed(0)
model = VAE(S, H, L).to(dev); opt = Adam(model.parameters(), lr=1e-4)
step = rdist.DataParallelTrainStep(model, opt, 1e-4, global_batch=B * world, graph=os.environ.get("DP_GRAPH", "1") == "1")
x = torch.rand(B, S, device=dev) * 2 - 1
for _ in range(10): step(x)
# raw all-reduce cost of the bucket sizes on an otherwise idle GPU (torch's NCCL communicator)
for n in (2048 * 1024, 2048 * 256, 512 * 2048):
    t = torch.zeros(n, device=dev)
    for _ in range(3): dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(20): dist.all_reduce(t)
    a1.record(); torch.cuda.synchronize()
    if rank == 0: print(f'  all_reduce {4*n/1e6:.1f} MB: {a0.elapsed_time(a1)/20*1e3:.1f} us')
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 100
e0.record()
for _ in range(n): step(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(10): step(x)
t1 = time.perf_counter(); torch.cuda.synchronize()
if rank == 0:
    print(f"W={world} NUM_SMS={os.environ.get('RVAE_NUM_SMS','all')} step {ms*1e3:.1f} us -> {B*world/ms/1e3:.2f} M frames/s; host enqueue {(t1-t0)*1e5:.1f} us/step")
dist.barrier(); dist.destroy_process_group()
