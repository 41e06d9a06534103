"""torchrun --nproc-per-node W tools/p2p_bench.py : the peer-memory all-reduce kernel alone (idle GPUs) - result check
against torch.distributed, per-phase timing of the slowest CTA, and NCCL's time for the same sizes."""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rawaudiovae_kelsey_b200 import dist as rdist, ops, _lib
from rawvae.model import VAE

rank, world, local = rdist.init_from_env("nccl")
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
model = VAE(1024, 2048, 256).to(dev)
rdist.init_native_comm(dev)
assert rdist.adopt_symmetric_grads(model)
flat = model._flat
lib = _lib.load()
g = flat.grads
for n in (5632, 2048 * 256, 512 * 2048, 2048 * 1024):
    torch.manual_seed(rank)
    g[:n] = torch.randn(n, device=dev)
    ref = g[:n].clone()
    dist.all_reduce(ref)
    aux = torch.zeros(64, 8, dtype=torch.int64, device=dev); aux[:, 0] = 2 ** 62
    torch.cuda.synchronize(); dist.barrier()
    _lib.check(lib.rvae_dp_allreduce(ops.ctx(dev), g.data_ptr(), n, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = float((g[:n] - ref).abs().max())
    ops.set_aux_trace(aux, 64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        _lib.check(lib.rvae_dp_allreduce(ops.ctx(dev), g.data_ptr(), n, 0, torch.cuda.current_stream().cuda_stream))
    e1.record(); torch.cuda.synchronize()
    ops.set_aux_trace(None)
    a = aux.cpu().numpy()[:20].astype(float)
    t = torch.zeros(n, device=dev)
    for _ in range(3): dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record()
    for _ in range(20): dist.all_reduce(t)
    n1.record(); torch.cuda.synchronize()
    if rank == 0:
        d = a[5:, 4:8].mean(0) / 1e3
        print(f"{4*n/1e6:6.2f} MB: p2p {e0.elapsed_time(e1)/20*1e3:6.1f} us/launch (kernel {((a[5:,1]-a[5:,0]).mean())/1e3:5.1f}: barrier1 {d[0]:.1f} reduce+push {d[1]:.1f} "
              f"barrier2 {d[2]:.1f}) max err {err:.2e} | nccl {n0.elapsed_time(n1)/20*1e3:6.1f} us")
dist.barrier(); dist.destroy_process_group()
