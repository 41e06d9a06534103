"""Inference throughput of the widened VAE (BASELINE.json configs[4]: segment_length 4096, n_units 4096, latent 256):
wav -> frames (hop 512) -> encode -> reparameterize -> decode -> overlap-add, batches of 16 384 frames."""
import time
import torch
from rawvae.model import VAE
from rawaudiovae_kelsey_b200 import ops

S, H, L, hop, B = 4096, 4096, 256, 512, 16384
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = VAE(S, H, L).to(dev).eval()
audio = torch.rand((B - 1) * hop + S, device=dev) * 2 - 1
eps = torch.randn(B, L, device=dev)

def run():
    frames, _, _ = ops.frame_gather(audio, B, hop, S)
    mu, lv = model.encode(frames)
    z = model.reparameterize(mu, lv, eps=eps)
    xh = model.decode(z)
    return ops.overlap_add(xh, hop)

for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = 73400320 * B
print(f"widened VAE inference: {ms*1e3:.1f} us per {B}-frame batch -> {B/ms/1e3:.2f} M frames/s, "
      f"{flop/ms/1e9:.0f} TFLOP/s ({flop/ms/1e9/1644.9*100:.1f} % of bf16 burst peak); output {out.numel()} samples")
