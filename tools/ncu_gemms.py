"""Launch each GEMM of one default.ini training step (B=8192) twice (warm + measured) for an ncu capture:
   ncu --set full -k regex:gemm_kernel -s 11 -c 11 python tools/ncu_gemms.py"""
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
for rep in range(2):
    ops.linear_act_fwd(x, w1, b1, ops.ACT_RELU)                                       # F1
    ops.encode_head_fwd(h, w2, b2, eps, kl_acc=acc[1:])                 # F2
    ops.linear_act_fwd(z, w3, b1, ops.ACT_RELU)                                       # F3
    ops.out_tanh_mse_fwd(h, w4, b4, x, grad_scale=1e-6, tanh_approx=True, want_xhat=False, mse_acc=acc[:1])  # F4
    ops.wgrad(x, h, out=dw4)                                                          # B4w
    ops.dgrad_relu(x, w4, h)                                                          # B4d
    ops.wgrad(h, z, out=dw3)                                                          # B3w
    ops.dgrad_latent(h, w3, eps, esh, esh)                                            # B3d
    ops.wgrad(ml, h, out=dw2)                                                         # B2w
    ops.dgrad_relu(ml, w2, h)                                                         # B2d
    ops.wgrad(h, x, out=dw1)                                                          # B1w
    torch.cuda.synchronize()
print("done")
