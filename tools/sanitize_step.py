"""Small-dims fused training steps (eager, no CUDA graph) for compute-sanitizer:

   compute-sanitizer --tool racecheck  python tools/sanitize_step.py
   compute-sanitizer --tool synccheck  python tools/sanitize_step.py
   compute-sanitizer --tool memcheck   python tools/sanitize_step.py

Covers the persistent tcgen05 GEMMs (every epilogue kind), the fused dgrad+wgrad launches, latent backward, Adam,
frame gather, Philox noise and overlap-add at S=256, H=320, L=64, B=512 (ragged: 500)."""
import sys
import torch
from rawvae.model import VAE
from rawaudiovae_kelsey_b200.model import FrameBatch, FusedTrainStep
from rawaudiovae_kelsey_b200.optim import Adam
from rawaudiovae_kelsey_b200 import ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)
S, H, L, hop = 256, 320, 64, 64
for B in (512, 500):
    model = VAE(S, H, L).to(dev)
    model.eps_seed = 1
    step = FusedTrainStep(model, Adam(model.parameters(), lr=1e-3), 1e-4, graph=False)
    audio = (torch.rand(200_000, device=dev) * 2 - 1)
    n_frames = (audio.numel() - S) // hop + 1
    idx = torch.randint(0, n_frames, (4, B), device=dev, dtype=torch.int64)
    fbs = [FrameBatch(audio, B, hop, S, frame_idx=idx[k]) for k in range(4)]
    for i in range(3):
        loss = step(fbs[i % 4], next_data=fbs[(i + 1) % 4])
    torch.cuda.synchronize()
    print("B", B, "loss", float(loss))
    with torch.no_grad():
        xh, mu, lv = model(torch.rand(B, S, device=dev))
        y = ops.overlap_add(xh.float().contiguous(), hop)
    torch.cuda.synchronize()
print("done")
