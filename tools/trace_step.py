"""Timeline of ONE fused training step (default.ini dims, B=8192): when each GEMM's CTAs enter / leave (globaltimer),
so the gaps between kernels, the overlap of forked kernels and the cost of the non-GEMM kernels become visible.
   python tools/trace_step.py      (needs a B200; debug aid, see rvae_debug_set_trace)"""
import os
import numpy as np
import torch
from rawvae.model import VAE, FusedTrainStep
from rawaudiovae_kelsey_b200.optim import Adam
from rawaudiovae_kelsey_b200 import ops

B, S, H, L = 8192, 1024, 2048, 256
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1:   # torchrun: trace the data-parallel step (rank 0 prints)
    from rawaudiovae_kelsey_b200 import dist as rdist
    RANK, WORLD, local = rdist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
else:
    dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = VAE(S, H, L).to(dev)
opt = Adam(model.parameters(), lr=1e-4)
if WORLD > 1:
    step = rdist.DataParallelTrainStep(model, opt, 1e-4, global_batch=B * WORLD)
else:
    step = FusedTrainStep(model, opt, 1e-4)
if RANK != 0:
    import builtins
    builtins.print = lambda *a, **k: None
if os.environ.get('STEP_PIPE', '0') == '1':
    from rawvae.model import FrameBatch
    torch.manual_seed(1 + RANK)
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
for _ in range(5):
    step(x)
torch.cuda.synchronize()
NAMES = ["F1", "F2", "F3", "F4", "B4d", "B4w", "B3d", "B3w", "B2d", "B2w", "B1w"]
if os.environ.get("RVAE_DUAL_PAIRS", "64") != "0":   # backward stages 0..2 are fused dgrad + wgrad launches
    NAMES = ["F1", "F2", "F3", "F4", "B4d+B4w", "B3d+B3w", "B2d+B2w", "B1w"]
    if os.environ.get("RVAE_MERGE_B3W", "0") != "0":      # experiments: stage 1 = latent dgrad + latent kernel; B3w rides in stage 2's launch
        NAMES = ["F1", "F2", "F3", "F4", "B4d+B4w", "B3d", "B2d+B2w+B3w", "B1w"]
    elif os.environ.get("RVAE_SPLIT_STAGE1", "1") != "0":   # stage 1: latent dgrad, then latent kernel || fc3 weight gradient
        NAMES = ["F1", "F2", "F3", "F4", "B4d+B4w", "B3d", "B3w", "B2d+B2w", "B1w"]
    if os.environ.get("RVAE_FUSE_FORWARD", "0") != "0":   # fc1 + head and fc3 + fc4 are chained launches
        NAMES = ["F1>F2>F3>F4", "B4d+B4w", "B3d+B3w", "B2d+B2w", "B1w"]
NSTEP = 3
W, HDR = ops.TRACE_WORDS_PER_CTA, ops.TRACE_HEADER
nsm = ops.num_sms()
nl = len(NAMES) * NSTEP
buf = torch.zeros(nl * W * nsm, dtype=torch.int64, device=dev)
AUXN = 24 * NSTEP
aux = torch.zeros(AUXN, 8, dtype=torch.int64, device=dev)
aux[:, 0] = 2 ** 62
ops.set_aux_trace(aux, AUXN)
ops.set_trace(buf, nl)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(NSTEP):
    step(x)
e1.record()
torch.cuda.synchronize()
ops.set_trace(None)
ops.set_aux_trace(None)
print(f"{NSTEP} traced steps: {1e3 * e0.elapsed_time(e1) / NSTEP:.1f} us/step")
t = buf.cpu().numpy().reshape(nl, nsm, W)
AUXK = {1: "gather", 2: "randn", 3: "latent", 4: "adam", 5: "allreduce"}
rows = []   # (start, text)
a = aux.cpu().numpy()
for j in range(AUXN):
    if a[j, 1] > 0:
        extra = ""
        if int(a[j, 2]) == 5:
            extra = "  [slowest CTA: barrier1 %.1f  reduce+push %.1f  barrier2 %.1f us]" % tuple(a[j, 4:7] / 1e3)
        rows.append((float(a[j, 0]), float(a[j, 1]), f"{AUXK.get(int(a[j, 2]), '?'):6s} blocks {int(a[j, 3]):4d}{extra}"))
t0 = None
prev_end = None
gemm_rows = []
for i in range(nl):
    hdr = t[i, :, :HDR].astype(np.float64)
    live = hdr[:, 1] > 0
    if not live.any():
        continue
    first, last_in = hdr[live, 0].min(), hdr[live, 0].max()
    first_out, last = hdr[live, 6].min(), hdr[live, 6].max()
    # time the CTAs spent waiting for the previous kernel (PDL wait) in cycles -> us using the CTA's own clock rate
    ghz = np.median((hdr[live, 5] - hdr[live, 1]) / np.maximum(hdr[live, 6] - hdr[live, 0], 1))
    pdl = np.median(hdr[live, 3] - hdr[live, 2]) / ghz / 1e3
    if i % len(NAMES) == 0:
        if t0 is not None:
            print(f"   ---- step period {(first - t0) / 1e3:.1f} us")
        t0 = first
    gap = "" if prev_end is None else f" gap {((first - prev_end) / 1e3):6.1f}"
    print(f"{NAMES[i % len(NAMES)]:11s} ctas {int(live.sum()):3d}  enter {((first - t0) / 1e3):7.1f}..{((last_in - t0) / 1e3):7.1f}  "
          f"exit {((first_out - t0) / 1e3):7.1f}..{((last - t0) / 1e3):7.1f}  span {((last - first) / 1e3):6.1f}  pdl-wait {pdl:5.1f}{gap}")
    prev_end = last
    gemm_rows.append((first, last, NAMES[i % len(NAMES)]))
base = gemm_rows[0][0]
print("---- merged timeline (us since the first GEMM of the first traced step)")
allr = [(f, l, f"GEMM {n}") for f, l, n in gemm_rows] + rows
for f, l, name in sorted(allr):
    print(f"  {(f - base) / 1e3:8.1f} -> {(l - base) / 1e3:8.1f}  ({(l - f) / 1e3:6.1f})  {name}")

if WORLD > 1:
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()

# per-tile role summary of launches of the last traced step (TRACE_LAUNCH = comma-separated indices within the step,
# default 0); TRACE_PAIRS = comma-separated CTA indices whose own unit-by-unit timeline is printed as well
NE, NT = ops.TRACE_EVENTS, ops.TRACE_TILES
split = int(os.environ.get("TRACE_SPLIT", "0"))
pairs = [int(v) for v in os.environ.get("TRACE_PAIRS", "").split(",") if v]
for li in [int(v) for v in os.environ.get("TRACE_LAUNCH", "0").split(",")]:
    i = (NSTEP - 1) * len(NAMES) + li
    hdr = t[i, :, :HDR].astype(np.float64)
    ev = t[i, :, HDR:].reshape(nsm, NT, NE).astype(np.float64)
    live = hdr[:, 1] > 0
    ghz = float(np.median((hdr[live, 5] - hdr[live, 1]) / np.maximum(hdr[live, 6] - hdr[live, 0], 1)))
    us = lambda c: c / ghz / 1e3
    print(f"---- roles of launch {li} ({NAMES[li]}), {int(live.sum())} CTAs, clock {ghz:.2f} GHz; times in us since the CTA passed its PDL wait; "
          f"CTA lifetime after the PDL wait: median {us(np.median(hdr[live, 5] - hdr[live, 3])):.1f} max {us(np.max(hdr[live, 5] - hdr[live, 3])):.1f}")
    groups = [("", live)]
    if split > 0:
        lo = live.copy(); lo[split:] = False
        hi = live.copy(); hi[:split] = False
        groups = [(f"ctas < {split}", lo), (f"ctas >= {split}", hi)]
    for gname, gl in groups:
        if gname:
            print(f" -- {gname}: exit {us(np.median(hdr[gl, 5] - hdr[gl, 3])):6.1f} us after the PDL wait (max {us(np.max(hdr[gl, 5] - hdr[gl, 3])):6.1f})")
        for it in range(NT):
            e = ev[gl, it, :]
            ran = e[:, 6] > 0
            if not ran.any():
                break
            e = e[ran]; base = hdr[gl, 3][ran]
            mm = e[:, 3] > 0
            med = lambda a: float(np.median(a)) if len(a) else float("nan")
            mx = lambda a: float(np.max(a)) if len(a) else float("nan")
            print(f"   tile {it:2d} ({int(ran.sum()):3d} CTAs): prod start {us(med(e[:,0]-base)):6.1f} | mma acc-wait {us(med(e[mm,3]-e[mm,2])):5.2f} "
                  f"data-wait med {us(med(e[mm,4]-e[mm,3])):5.2f} max {us(mx(e[mm,4]-e[mm,3])):5.2f} issue {us(med(e[mm,5]-e[mm,4])):5.2f} commit at {us(med(e[mm,5]-base[mm])):6.1f} (max {us(mx(e[mm,5]-base[mm])):6.1f}) | "
                  f"epi {us(med(e[:,6]-base)):6.1f}->{us(med(e[:,7]-base)):6.1f}"
                  + ("  ev10..15 after acc-ready: " + " ".join(f"{us(med(e[e[:,k]>0,k]-e[e[:,k]>0,6])):5.1f}" for k in range(10, 16))
                     if os.environ.get("TRACE_EPI_DETAIL") else ""))
    for c in pairs:
        if c >= nsm or hdr[c, 1] <= 0:
            continue
        b0 = hdr[c, 3]
        print(f"   CTA {c}: exits {us(hdr[c, 5] - b0):.1f} us after its PDL wait")
        for it in range(NT):
            e = ev[c, it, :]
            if e[6] <= 0:
                break
            f = lambda k: us(e[k] - b0) if e[k] > 0 else float('nan')
            print(f"      unit {it:2d}: prod {f(0):6.1f}..{f(1):6.1f} | mma wait-acc {f(2):6.1f} got {f(3):6.1f} data {f(4):6.1f} commit {f(5):6.1f} | "
                  f"epi team0 {f(6):6.1f}..{f(7):6.1f} team1 {f(8):6.1f}..{f(9):6.1f}")
