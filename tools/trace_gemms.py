"""Role timeline of each GEMM of one default.ini training step (B=8192): where a CTA's time goes.
   python tools/trace_gemms.py [names...]      (needs a B200; debug aid, see rvae_debug_set_trace)"""
import sys
import numpy as np
import torch
from rawaudiovae_kelsey_b200 import ops

dev = "cuda"
torch.manual_seed(0)
g = lambda *s: torch.randn(*s, device=dev)
bf = lambda t: t.to(torch.bfloat16)
B, S_, H_, L_ = 8192, 1024, 2048, 256
x, h, z, ml = bf(g(B, S_)), bf(g(B, H_)), bf(g(B, L_)), bf(g(B, 2 * L_))
w1, w2, w3, w4 = bf(g(H_, S_)), bf(g(2 * L_, H_)), bf(g(H_, L_)), bf(g(S_, H_))
b1, b2, b4 = g(H_), g(2 * L_), g(S_)
eps, esh = g(B, L_), g(B, L_) * 0.3
acc = torch.zeros(2, dtype=torch.float64, device=dev)
dw4 = torch.zeros(S_, H_, device=dev); dw3 = torch.zeros(H_, L_, device=dev)
dw2 = torch.zeros(2 * L_, H_, device=dev); dw1 = torch.zeros(H_, S_, device=dev)
bg = torch.zeros(H_, device=dev)
GEMMS = {
    "F1": lambda: ops.linear_act_fwd(x, w1, b1, ops.ACT_RELU),
    "F2": lambda: ops.encode_head_fwd(h, w2, b2, eps, kl_acc=acc[1:]),
    "F3": lambda: ops.linear_act_fwd(z, w3, b1, ops.ACT_RELU),
    "F4": lambda: ops.out_tanh_mse_fwd(h, w4, b4, x, grad_scale=1e-6, tanh_approx=True, want_xhat=False, mse_acc=acc[:1], bias_grad=bg[:S_]),
    "B4w": lambda: ops.wgrad(x, h, out=dw4),
    "B4d": lambda: ops.dgrad_relu(x, w4, h, bias_grad=bg),
    "B3w": lambda: ops.wgrad(h, z, out=dw3),
    "B3d": lambda: ops.dgrad_latent(h, w3, eps, esh, esh, bias_grad=bg[:2 * L_]),
    "B2w": lambda: ops.wgrad(ml, h, out=dw2),
    "B2d": lambda: ops.dgrad_relu(ml, w2, h, bias_grad=bg),
    "B1w": lambda: ops.wgrad(h, x, out=dw1),
}
W, HDR, NT, NE = ops.TRACE_WORDS_PER_CTA, ops.TRACE_HEADER, ops.TRACE_TILES, ops.TRACE_EVENTS
nsm = ops.num_sms()
buf = torch.zeros(W * nsm, dtype=torch.int64, device=dev)
names = sys.argv[1:] or list(GEMMS)
for name in names:
    fn = GEMMS[name]
    ops.set_trace(None)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    ops.set_trace(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ops.set_trace(None)
    t = buf.cpu().numpy().reshape(nsm, W)
    hdr = t[:, :HDR].astype(np.float64)
    ev = t[:, HDR:].reshape(nsm, NT, NE).astype(np.float64)
    live = hdr[:, 1] > 0
    n = int(live.sum())
    gt0 = hdr[live, 0].min()
    span_ns = hdr[live, 6].max() - gt0
    cyc = (hdr[live, 5] - hdr[live, 1])
    ghz = float(np.median(cyc / np.maximum(hdr[live, 6] - hdr[live, 0], 1)))
    us = lambda c: c / ghz / 1e3
    print(f"== {name}: event-timed {1e3 * e0.elapsed_time(e1):.1f} us; {n} CTAs; first-entry -> last-exit {span_ns/1e3:.1f} us; "
          f"clock {ghz:.2f} GHz")
    print(f"   CTA entry skew {np.ptp(hdr[live,0])/1e3:.1f} us; setup {us(np.median(hdr[live,2]-hdr[live,1])):.2f} us; "
          f"PDL wait {us(np.median(hdr[live,3]-hdr[live,2])):.2f} us; CTA lifetime median {us(np.median(cyc)):.1f} max {us(cyc.max()):.1f} us; "
          f"exit skew {np.ptp(hdr[live,6])/1e3:.1f} us")
    # per tile iteration statistics over CTAs that ran that iteration
    for it in range(NT):
        e = ev[live, it, :]
        ran = e[:, 6] > 0
        if not ran.any():
            break
        e = e[ran]; base = hdr[live, 3][ran]
        rel = lambda k: us(np.median(e[:, k] - base))
        mm = e[:, 3] > 0   # leader CTAs only issue MMAs
        def med(a): return float(np.median(a)) if len(a) else float("nan")
        print(f"   tile {it:2d} ({int(ran.sum()):3d} CTAs): prod {rel(0):6.2f}->{rel(1):6.2f} | "
              f"mma free-wait {us(med(e[mm,3]-e[mm,2])):5.2f} data-wait {us(med(e[mm,4]-e[mm,3])):5.2f} issue {us(med(e[mm,5]-e[mm,4])):5.2f} (commit at {us(med(e[mm,5]-base[mm])):6.2f}) | "
              f"epi0 {rel(6):6.2f}->{rel(7):6.2f} ({us(np.median(e[:,7]-e[:,6])):5.2f}) epi1 {rel(8):6.2f}->{rel(9):6.2f} ({us(np.median(e[:,9]-e[:,8])):5.2f})")
        if (e[:, 10] > 0).any():
            d = lambda a, b: us(np.median(e[:, a] - e[:, b]))
            print(f"            team0 unit0: ready->acquired {d(10,6):5.2f} ->ld done {d(11,10):5.2f} ->staged {d(12,11):5.2f} ->committed {d(13,12):5.2f}"
                  f" | unit1: acquired +{d(14,13):5.2f} ->committed {d(15,14):5.2f} | ->released {d(7,15):5.2f}")
