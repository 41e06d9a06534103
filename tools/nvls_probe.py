"""Probe: does this box give NVLS multicast mappings through torch's symmetric memory? (torchrun, >= 2 ranks)"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
try:
    from torch._C._distributed_c10d import _SymmetricMemory
    print(rank, "has_multicast_support:", _SymmetricMemory.has_multicast_support(torch.device("cuda").type and __import__("torch").distributed.distributed_c10d.DeviceType.CUDA if False else torch._C._autograd.DeviceType.CUDA, local))
except Exception as e:
    print(rank, "has_multicast_support probe failed:", repr(e)[:200])
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok: multicast_ptr", hex(hdl.multicast_ptr), "buffers", [hex(p) for p in hdl.buffer_ptrs][:4],
          "signal pads", len(hdl.signal_pad_ptrs), "world", hdl.world_size, flush=True)
    t.fill_(rank + 1.0)
    dist.barrier()
    torch.cuda.synchronize()
    if hdl.multicast_ptr:
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(rank, "multimem_all_reduce_ ->", float(t[0]), "expected", world * (world + 1) / 2, flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symmetric memory failed:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
