"""GPU bring-up: each op of librvae_b200 against plain torch fp32 on the same inputs. Run one case per process
(`python tools/bringup.py <case>`) so a device-side trap cannot poison later cases; `all` spawns them."""
import subprocess
import sys
import time

import torch

CASES = ["linear", "linear_small", "head", "out", "dgrad", "dz", "wgrad", "wgrad_split", "linear_fp32", "elementwise"]


def rel(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def bf(t):
    return t.to(torch.bfloat16)


def main(case):
    from rawaudiovae_kelsey_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    g = lambda *s: torch.randn(*s, device=dev)
    if case in ("linear", "linear_small", "linear_fp32"):
        shapes = [(8192, 2048, 1024), (8192, 2048, 256), (1000, 1024, 2048)] if case != "linear_small" else [(128, 256, 64), (256, 128, 128), (130, 192, 320)]
        for (M, N, K) in shapes:
            x, w, b = g(M, K), g(N, K) * 0.05, g(N)
            if case == "linear_fp32":
                xp, wp = ops.split_bf16(x, True), ops.split_bf16(w, True)
                _, _, y = ops.linear_act_fwd(xp, wp, b, ops.ACT_RELU, out_bf16=False, out_f32=True)
                ref = torch.relu(x.double() @ w.double().T + b.double())
            else:
                xb, wb = bf(x), bf(w)
                yb, _, y = ops.linear_act_fwd(xb, wb, b, ops.ACT_RELU, out_bf16=True, out_f32=True)
                ref = torch.relu(xb.double() @ wb.double().T + b.double())
                torch.cuda.synchronize()
                print(case, (M, N, K), "bf16 stream rel/maxabs", rel(yb, ref), flush=True)
            torch.cuda.synchronize()
            print(case, (M, N, K), "rel/maxabs", rel(y, ref), flush=True)
        # timing
        M, N, K = 8192, 2048, 1024
        x, w, b = bf(g(M, K)), bf(g(N, K)), g(N)
        for _ in range(3):
            ops.linear_act_fwd(x, w, b, ops.ACT_RELU)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.linear_act_fwd(x, w, b, ops.ACT_RELU)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{case} 8192x2048x1024: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    elif case == "head":
        for (M, L, K) in [(8192, 256, 2048), (300, 64, 128)]:
            h, w2, b2, eps = g(M, K), g(2 * L, K) * 0.02, g(2 * L) * 0.1, g(M, L)
            hb, wb = bf(h), bf(w2)
            acc = torch.zeros(2, dtype=torch.float64, device=dev)
            mu, lv, (z, _) = ops.encode_head_fwd(hb, wb, b2, eps, kl_acc=acc[1:])
            torch.cuda.synchronize()
            ml = hb.double() @ wb.double().T + b2.double()
            rmu, rlv = ml[:, :L], ml[:, L:]
            sig = torch.exp(0.5 * rlv)
            print("head", (M, L, K), "mu", rel(mu, rmu), "lv", rel(lv, rlv), "z", rel(z, rmu + eps.double() * sig),
                  "kl", float(acc[1]), float((1 + rlv - rmu ** 2 - torch.exp(rlv)).sum()), flush=True)
            mu2, lv2, _ = ops.encode_head_fwd(hb, wb, b2, None, want_z=False)
            print("head (encode only)", rel(mu2, rmu), rel(lv2, rlv), flush=True)
    elif case == "out":
        for (M, S, K) in [(8192, 1024, 2048), (200, 128, 64)]:
            h, w, b, x = g(M, K), g(S, K) * 0.02, g(S) * 0.1, torch.rand(M, S, device=dev) * 2 - 1
            hb, wb, xb = bf(h), bf(w), bf(x)
            acc = torch.zeros(2, dtype=torch.float64, device=dev)
            c0 = 2.0 / (M * S)
            bg = torch.zeros(S, device=dev)
            xhat, (da, _) = ops.out_tanh_mse_fwd(hb, wb, b, xb, grad_scale=c0, mse_acc=acc[:1], bias_grad=bg)
            torch.cuda.synchronize()
            r = torch.tanh(hb.double() @ wb.double().T + b.double())
            d = r - xb.double()
            print("out", (M, S, K), "xhat", rel(xhat, r), "da", rel(da, c0 * d * (1 - r * r)), "mse", float(acc[0]), float((d * d).sum()),
                  "db", rel(bg, (c0 * d * (1 - r * r)).sum(0)), flush=True)
    elif case == "dgrad":
        for (M, N, Kd) in [(8192, 2048, 1024), (8192, 2048, 512), (200, 128, 64)]:
            dy, w, h = g(M, Kd), g(Kd, N) * 0.05, g(M, N)
            dyb, wb, hb = bf(dy), bf(w), bf(h)
            bg = torch.zeros(N, device=dev)
            dx, _ = ops.dgrad_relu(dyb, wb, hb, bias_grad=bg)
            torch.cuda.synchronize()
            ref = (dyb.double() @ wb.double()) * (hb.double() > 0)
            print("dgrad", (M, N, Kd), rel(dx, ref), "db", rel(bg, ref.sum(0)), flush=True)
    elif case == "dz":
        for (M, L, H) in [(8192, 256, 2048), (200, 64, 128)]:
            da3, w3, eps, lv, mu = g(M, H), g(H, L) * 0.05, g(M, L), g(M, L) * 0.3, g(M, L)
            db, wb = bf(da3), bf(w3)
            c0 = 1e-2
            dz = db.double() @ wb.double()
            sig = torch.exp(0.5 * lv.double())
            esh = 0.5 * eps.double() * sig
            bg = torch.zeros(2 * L, device=dev)
            dml, _ = ops.dgrad_latent(db, wb, eps, lv, mu, kl_grad_scale=c0, bias_grad=bg)
            torch.cuda.synchronize()
            ref = torch.cat([dz + c0 * mu.double(), dz * esh + 0.5 * c0 * (sig * sig - 1)], 1)
            print("dz fused", (M, L, H), rel(dml, ref), "db", rel(bg, ref.sum(0)), flush=True)
            gmu, glv = g(M, L), g(M, L)
            dml, _ = ops.dgrad_latent(db, wb, eps, lv, None, g_mu=gmu, g_logvar=glv)
            torch.cuda.synchronize()
            ref = torch.cat([dz + gmu.double(), dz * esh + glv.double()], 1)
            print("dz external", (M, L, H), rel(dml, ref), flush=True)
    elif case in ("wgrad", "wgrad_split"):
        for (B, M, N) in [(8192, 1024, 2048), (8192, 2048, 256), (8192, 512, 2048), (1000, 128, 64), (8192, 2048, 1024)]:
            dy, x = g(B, M), g(B, N)
            dyb, xb = bf(dy), bf(x)
            dw = ops.wgrad(dyb, xb, k_splits=(1 if case == "wgrad" else 0))
            torch.cuda.synchronize()
            ref = dyb.double().T @ xb.double()
            print(case, (B, M, N), rel(dw, ref), flush=True)
    elif case == "elementwise":
        # framing
        n = 100000
        audio = torch.rand(n, device=dev) * 2 - 1
        hop, S = 128, 1024
        P = -(-n // hop) * hop
        N = P // hop - S // hop + 1
        pad = torch.cat([audio, torch.zeros(P - n, device=dev)])
        ref = pad.unfold(0, S, hop)
        f32, hi, lo = ops.frame_gather(audio, N, hop, S, out_bf16=True, out_lo=True)
        print("frames exact", bool((f32 == ref).all()), "hi exact", bool((hi == ref.to(torch.bfloat16)).all()),
              "hi+lo", rel(hi.float() + lo.float(), ref))
        idx = torch.randperm(N, device=dev)[:700].contiguous()
        f32, _, _ = ops.frame_gather(audio, 700, hop, S, frame_idx=idx)
        print("gather exact", bool((f32 == ref[idx]).all()))
        a16 = (audio * 32767).to(torch.int16)
        f32, _, _ = ops.frame_gather(a16, N, hop, S)
        pad16 = torch.cat([a16.float() / 32768, torch.zeros(P - n, device=dev)])
        print("i16 exact", bool((f32 == pad16.unfold(0, S, hop)).all()))
        ola = ops.overlap_add(ref.contiguous(), hop)
        print("ola==pad", rel(ola, pad))
        e = ops.randn((1 << 20,), seed=1)
        print("randn mean/std", float(e.mean()), float(e.std()), "kurt", float((e ** 4).mean()))
        # loss
        B, S, L = 512, 1024, 256
        xh, x, mu, lv = torch.tanh(g(B, S)), torch.rand(B, S, device=dev), g(B, L) * 0.1, g(B, L) * 0.1
        l = ops.loss_fwd(xh, x, mu, lv, 1e-4)
        refl = torch.nn.functional.mse_loss(xh.double(), x.double()) + 1e-4 * (-0.5) * torch.mean(1 + lv.double() - mu.double() ** 2 - lv.double().exp())
        print("loss", float(l), float(refl))
        xr, mr, lr_ = xh.double().requires_grad_(), mu.double().requires_grad_(), lv.double().requires_grad_()
        (torch.nn.functional.mse_loss(xr, x.double()) + 1e-4 * (-0.5) * torch.mean(1 + lr_ - mr ** 2 - lr_.exp())).backward()
        gx, gm, gl = ops.loss_bwd(xh, x, mu, lv, 1e-4, None)
        print("loss_bwd", rel(gx, xr.grad), rel(gm, mr.grad), rel(gl, lr_.grad))
        da, _ = ops.tanh_bwd(gx, xh)
        print("tanh_bwd", rel(da, gx.double() * (1 - xh.double() ** 2)))
        m = bf(g(1000, 256))
        print("colsum", rel(ops.colsum(m), m.double().sum(0)))
        # adam
        n = 100003 * 4
        p, gr = g(n), g(n)
        pr = p.clone().requires_grad_()
        opt = torch.optim.Adam([pr], lr=1e-3)
        m_, v_, st = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros((), device=dev)
        sh = torch.empty(n, dtype=torch.bfloat16, device=dev)
        for i in range(3):
            pr.grad = gr * (i + 1)
            opt.step()
            ops.adam_step(p, gr * (i + 1), m_, v_, st, 1e-3, shadow_hi=sh)
        print("adam", rel(p, pr.detach()), "step", float(st), "shadow", rel(sh, p.to(torch.bfloat16)))
    elif case == "perf":
        import os
        tag = f"CG={os.environ.get('RVAE_CTA_GROUP','auto')} BN={os.environ.get('RVAE_BLOCK_N','auto')}"
        B, S_, H_, L_ = 8192, 1024, 2048, 256
        def timeit(fn, flops, name):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"{tag} {name:5s} {ms*1e3:7.1f} us {flops/ms/1e9:7.1f} TFLOP/s", flush=True)
            return ms
        x, h, z, ml = bf(g(B, S_)), bf(g(B, H_)), bf(g(B, L_)), bf(g(B, 2 * L_))
        w1, w2, w3, w4 = bf(g(H_, S_)), bf(g(2 * L_, H_)), bf(g(H_, L_)), bf(g(S_, H_))
        b1, b2, b4 = g(H_), g(2 * L_), g(S_)
        eps, esh = g(B, L_), g(B, L_) * 0.3
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        dw4 = torch.zeros(S_, H_, device=dev); dw3 = torch.zeros(H_, L_, device=dev)
        dw2 = torch.zeros(2 * L_, H_, device=dev); dw1 = torch.zeros(H_, S_, device=dev)
        tot = 0.0
        tot += timeit(lambda: ops.linear_act_fwd(x, w1, b1, ops.ACT_RELU), 2 * B * H_ * S_, "F1")
        tot += timeit(lambda: ops.encode_head_fwd(h, w2, b2, eps, kl_acc=acc[1:]), 2 * B * 2 * L_ * H_, "F2")
        tot += timeit(lambda: ops.linear_act_fwd(z, w3, b1, ops.ACT_RELU), 2 * B * H_ * L_, "F3")
        tot += timeit(lambda: ops.out_tanh_mse_fwd(h, w4, b4, x, grad_scale=1e-6, tanh_approx=True, want_xhat=False, mse_acc=acc[:1]), 2 * B * S_ * H_, "F4")
        tot += timeit(lambda: ops.wgrad(x, h, out=dw4), 2 * B * S_ * H_, "B4w")
        tot += timeit(lambda: ops.dgrad_relu(x, w4, h), 2 * B * S_ * H_, "B4d")
        tot += timeit(lambda: ops.wgrad(h, z, out=dw3), 2 * B * H_ * L_, "B3w")
        tot += timeit(lambda: ops.dgrad_latent(h, w3, eps, esh, esh), 2 * B * H_ * L_, "B3d")
        tot += timeit(lambda: ops.wgrad(ml, h, out=dw2), 2 * B * 2 * L_ * H_, "B2w")
        tot += timeit(lambda: ops.dgrad_relu(ml, w2, h), 2 * B * 2 * L_ * H_, "B2d")
        tot += timeit(lambda: ops.wgrad(h, x, out=dw1), 2 * B * H_ * S_, "B1w")
        print(f"{tag} TOTAL {tot*1e3:.1f} us -> {30408704*B/tot/1e9:.1f} TFLOP/s chain", flush=True)
    print("launches", ops.launch_count(), flush=True)


if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "all"
    if case == "all":
        rc = 0
        for c in CASES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, c], timeout=90)
                code = r.returncode
            except subprocess.TimeoutExpired:
                code = "TIMEOUT"
            print(f"=== {c}: exit {code} in {time.time()-t0:.1f}s", flush=True)
            rc |= int(code != 0)
        sys.exit(rc)
    main(case)
