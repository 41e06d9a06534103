"""Launch each HBM-bound kernel of the path once warm + once measured, at the sizes of a default.ini training step
(B = 8192) and of the widened inference config, for an ncu capture of their DRAM traffic:

   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
       -k regex:'adam_kernel|frame_gather|latent_bwd|randn_kernel|overlap_add|loss_fwd|loss_bwd|reparam_kernel|tanh_bwd|split_bf16|colsum' \
       --csv --log-file gpurun_out/hbm_kernels.csv python tools/ncu_hbm_kernels.py --ncu

Without ncu it times the same launches with CUDA events (L2 flushed between launches) and writes
gpurun_out/hbm_kernels_events.json: {kernel: {us, algorithmic_bytes, GB/s}} - the algorithmic byte counts are the
ones DESIGN.md 4.2 states."""
import json
import os
import sys

import torch

from rawaudiovae_kelsey_b200 import ops

dev = "cuda"
torch.manual_seed(0)
B, S, H, L, hop = 8192, 1024, 2048, 256, 128
NP = 5772800
g = lambda *s: torch.randn(*s, device=dev)

audio = g(30 * 44100 * 32).clamp_(-1, 1)
audio16 = torch.round(audio * 32767).to(torch.int16)
nfr_total = (audio.numel() - S) // hop + 1
idx = torch.randint(0, nfr_total, (B,), device=dev, dtype=torch.int64)
p, gr, m, v = g(NP), g(NP), g(NP), g(NP).abs_()
step = torch.zeros((), device=dev)
shadow = torch.empty(NP, dtype=torch.bfloat16, device=dev)
da3 = g(B, H).to(torch.bfloat16)
w3 = g(H, L).to(torch.bfloat16)
eps, lv, mu = g(B, L), g(B, L) * 0.3, g(B, L)
xhat, x = torch.tanh(g(B, S)), g(B, S).clamp_(-1, 1)
frames_w = g(16384, 4096)
frames_d = g(B, S)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

CASES = [
    # name, callable, algorithmic bytes
    ("frame_gather (fp32 wav -> bf16 frames, random 8192 of the corpus)",
     lambda: ops.frame_gather(audio, B, hop, S, frame_idx=idx, out_f32=False, out_bf16=True), B * S * (4 + 2)),
    ("frame_gather pcm16 (int16 wav -> bf16 frames, random 8192 of the corpus)",
     lambda: ops.frame_gather(audio16, B, hop, S, frame_idx=idx, out_f32=False, out_bf16=True), B * S * (2 + 2)),
    ("randn (eps [8192,256] fp32)", lambda: ops.randn((B, L), 1, 0), B * L * 4),
    ("adam (5 772 800 params + bf16 shadow, gradient cleared)",
     lambda: ops.adam_step(p, gr, m, v, step, 1e-4, shadow_hi=shadow, increment_step=False), NP * 30),
    ("latent_bwd (via dgrad_latent: reads dz, mu, logvar, eps; writes d_ml bf16)",
     lambda: ops.dgrad_latent(da3, w3, eps, lv, mu, kl_grad_scale=1e-6), B * L * (16 + 4)),
    ("loss_fwd (API path: MSE + KL reductions, fp32 inputs)", lambda: ops.loss_fwd(xhat, x, mu, lv, 1e-4),
     B * S * 8 + B * L * 8),
    ("loss_bwd (API path)", lambda: ops.loss_bwd(xhat, x, mu, lv, 1e-4, None), B * S * 12 + B * L * 16),
    ("reparameterize (API path)", lambda: ops.reparameterize(mu, lv, eps), B * L * 16),
    ("overlap_add default (8192 x 1024 frames, hop 128)", lambda: ops.overlap_add(frames_d, hop),
     B * S * 4 + B * hop * 4),
    ("overlap_add widened (16384 x 4096 frames, hop 512)", lambda: ops.overlap_add(frames_w, 512),
     16384 * 4096 * 4 + 16384 * 512 * 4),
]

under_ncu = "--ncu" in sys.argv
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""
out = {}
for name, fn, nbytes in CASES:
    if only and only not in name:
        continue
    fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(1 if under_ncu else 9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(times)[len(times) // 2]
    out[name] = {"us_events_incl_launch": round(us, 2), "algorithmic_bytes": nbytes,
                 "GB_per_s_events": round(nbytes / us / 1e3, 1)}
    print(f"{name:80s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / us / 1e3:8.1f} GB/s", flush=True)
if not under_ncu and not only:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/hbm_kernels_events.json", "w"), indent=1)
print("done")
