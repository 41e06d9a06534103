"""CPU oracle for the rawaudiovae hot path.  TEST INFRASTRUCTURE - NOT A PRODUCT PATH.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
Nothing under rawaudiovae_kelsey_b200/ or rawvae/ imports it, and the product path has no CPU fallback.

What it is: a plain restatement of the reference's algorithm (kelseyicotton/rawaudiovae_kelsey) for the path
BASELINE.json names - VAE forward, loss, the backward pass autograd derives for them, torch.optim.Adam's update,
and the framing / resynthesis rules of rawvae.dataset - with every step written out explicitly (no nn.Module,
no autograd, no torch.optim), each function citing the reference file:line it follows. Arithmetic uses torch CPU
tensors (float64 for ground truth, float32 to mirror the reference's own precision): the reference's arithmetic
lives in PyTorch/ATen (third-party, unpinned; "tested with torch 2.0.1", README.md:3), which is the library
installed here (torch 2.11.0), so the same ATen GEMMs back both.

Pinning: the reference ships no tests or golden vectors. This oracle is pinned against
outputs of the reference ITSELF: oracle/gen_golden.py imports /root/reference/rawvae/{model,dataset}.py
unmodified, runs them on seeded inputs with injected eps, and commits the results under tests/golden/;
tests/test_oracle_vs_golden.py checks every function below against those fixtures. The overlap-add mode of
resynthesis has no reference implementation (tutorial.ipynb only concatenates, :543,932,1289) - that one function
is "parity unpinned" and is checked through its defining identity ola(frames(x)) == pad(x) instead.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

PARAM_NAMES = ("fc1.weight", "fc1.bias", "fc21.weight", "fc21.bias", "fc22.weight", "fc22.bias",
               "fc3.weight", "fc3.bias", "fc4.weight", "fc4.bias")


# ------------------------------------------------------------------------------------------------ model
def init_params(segment_length: int, n_units: int, latent_dim: int, seed: int = 0,
                dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Parameters in the reference's creation order and default nn.Linear init (rawvae/model.py:13-17):
    kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight, same bound for bias.
    Consumes torch's global RNG exactly as `torch.manual_seed(seed); VAE(S, H, L)` does."""
    torch.manual_seed(seed)
    shapes = [("fc1", n_units, segment_length), ("fc21", latent_dim, n_units), ("fc22", latent_dim, n_units),
              ("fc3", n_units, latent_dim), ("fc4", segment_length, n_units)]
    params = {}
    for name, out_f, in_f in shapes:
        w = torch.empty(out_f, in_f)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_f)
        b = torch.empty(out_f)
        torch.nn.init.uniform_(b, -bound, bound)
        params[name + ".weight"] = w.to(dtype)
        params[name + ".bias"] = b.to(dtype)
    return params


def bf16_operands(t: torch.Tensor) -> torch.Tensor:
    """Round to bfloat16 and back: models the "bf16 mode" of BASELINE.json, in which every GEMM operand (inputs,
    weights, activations and activation gradients) is carried as bf16 while accumulation, biases, the elementwise
    math and the master weights stay in higher precision."""
    return t.to(torch.bfloat16).to(t.dtype)


def _ident(t: torch.Tensor) -> torch.Tensor:
    return t


def forward(params: Dict[str, torch.Tensor], x: torch.Tensor, eps: torch.Tensor, q=_ident) -> Dict[str, torch.Tensor]:
    """VAE.forward with the noise made explicit (rawvae/model.py:19-35).
    encode :19-21, reparameterize :23-26 (eps replaces torch.randn_like), decode :28-30, view(-1, S) :33.
    `q` is applied to every GEMM operand: identity = the reference's arithmetic; bf16_operands = bf16 mode."""
    S = params["fc1.weight"].shape[1]
    x = q(x.reshape(-1, S))
    a1 = x @ q(params["fc1.weight"]).T + params["fc1.bias"]           # model.py:20  fc1
    h1 = q(torch.clamp_min(a1, 0))                                    # model.py:20  relu
    mu = h1 @ q(params["fc21.weight"]).T + params["fc21.bias"]        # model.py:21
    logvar = h1 @ q(params["fc22.weight"]).T + params["fc22.bias"]    # model.py:21
    std = torch.exp(0.5 * logvar)                                     # model.py:24
    z = q(mu + eps * std)                                             # model.py:26
    a3 = z @ q(params["fc3.weight"]).T + params["fc3.bias"]           # model.py:29
    h3 = q(torch.clamp_min(a3, 0))
    x_hat = torch.tanh(h3 @ q(params["fc4.weight"]).T + params["fc4.bias"])  # model.py:30
    return dict(x=x, a1=a1, h1=h1, mu=mu, logvar=logvar, std=std, eps=eps, z=z, a3=a3, h3=h3, x_hat=x_hat)


def loss_function(x_hat, x, mu, logvar, kl_beta: float, segment_length: int) -> torch.Tensor:
    """rawvae/model.py:38-46: mse_loss(mean over B*S) + kl_beta * (-0.5 * mean over B*L of 1+lv-mu^2-e^lv)."""
    x = x.reshape(-1, segment_length)
    recon = ((x_hat - x) ** 2).sum() / x_hat.numel()                   # model.py:39
    kld = -0.5 * (1 + logvar - mu ** 2 - torch.exp(logvar)).sum() / mu.numel()  # model.py:45
    return recon + kl_beta * kld                                      # model.py:46


def backward(params: Dict[str, torch.Tensor], act: Dict[str, torch.Tensor], kl_beta: float,
             grad_out: float = 1.0, q=_ident, masks=None) -> Dict[str, torch.Tensor]:
    """What loss.backward() (train.py:191, train_iterable.py:208) computes for the graph above, written out
    (SURVEY.md Appendix A): 5 weight gradients, 5 bias gradients, no gradient for x. Also returns the
    activation gradients the kernels materialise (da4, da3, dmu, dlv, da1). `q` as in forward().
    `masks=(m1, m3)` overrides the ReLU gates [h1 > 0], [h3 > 0] - used by the parity tests to evaluate the reference
    gradient at the implementation's gating pattern (the gate of a unit whose pre-activation is ~0 is ill-conditioned:
    any rounding flips it, and a flipped gate changes that unit's gradient by 100 %)."""
    m1 = (act["h1"] > 0) if masks is None else masks[0].to(act["h1"].dtype)
    m3 = (act["h3"] > 0) if masks is None else masks[1].to(act["h3"].dtype)
    x, h1, mu, lv, std, eps, z, h3, xh = (act[k] for k in ("x", "h1", "mu", "logvar", "std", "eps", "z", "h3", "x_hat"))
    B, S = xh.shape
    L = mu.shape[1]
    dxh = grad_out * 2.0 * (xh - x) / (B * S)                         # MseLossBackward
    da4 = q(dxh * (1 - xh * xh))                                      # TanhBackward
    g = {}
    g["fc4.weight"] = da4.T @ h3                                      # AddmmBackward (wgrad)
    g["fc4.bias"] = da4.sum(0)
    dh3 = da4 @ q(params["fc4.weight"])                               # AddmmBackward (dgrad)
    da3 = q(dh3 * m3)                                                 # ReluBackward (threshold_backward)
    g["fc3.weight"] = da3.T @ z
    g["fc3.bias"] = da3.sum(0)
    dz = da3 @ q(params["fc3.weight"])
    dmu = q(dz + grad_out * kl_beta * mu / (B * L))                   # reparam + KL branch
    dlv = q(dz * eps * std * 0.5 + grad_out * kl_beta * (torch.exp(lv) - 1) / (2 * B * L))
    g["fc21.weight"] = dmu.T @ h1
    g["fc21.bias"] = dmu.sum(0)
    g["fc22.weight"] = dlv.T @ h1
    g["fc22.bias"] = dlv.sum(0)
    dh1 = dmu @ q(params["fc21.weight"]) + dlv @ q(params["fc22.weight"])
    da1 = q(dh1 * m1)
    g["fc1.weight"] = da1.T @ x
    g["fc1.bias"] = da1.sum(0)
    g["_act"] = dict(da4=da4, da3=da3, dmu=dmu, dlv=dlv, da1=da1)
    return g


def adam_init(params: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    return {k: dict(step=0, exp_avg=torch.zeros_like(v), exp_avg_sq=torch.zeros_like(v)) for k, v in params.items()}


def adam_step(params, grads, state, lr: float, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """torch.optim.Adam defaults as constructed at train.py:163 / train_iterable.py:180 (betas (0.9, 0.999),
    eps 1e-8, no weight decay, no amsgrad), single-tensor update order lerp_/mul_+addcmul_/sqrt/div/add_/addcdiv_."""
    b1, b2 = betas
    for k in PARAM_NAMES:
        p, g, st = params[k], grads[k], state[k]
        st["step"] += 1
        t = st["step"]
        st["exp_avg"] += (1 - b1) * (g - st["exp_avg"])               # lerp_
        st["exp_avg_sq"].mul_(b2).add_((1 - b2) * g * g)              # mul_ + addcmul_
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = st["exp_avg_sq"].sqrt() / math.sqrt(bc2) + eps
        p -= (lr / bc1) * st["exp_avg"] / denom                       # addcdiv_


def train_step(params, state, x, eps, kl_beta: float, lr: float, q=_ident) -> float:
    """One iteration of the training-loop body (train_iterable.py:200-210): forward, loss, backward, Adam."""
    act = forward(params, x, eps, q)
    S = params["fc1.weight"].shape[1]
    loss = loss_function(act["x_hat"], act["x"], act["mu"], act["logvar"], kl_beta, S)
    grads = backward(params, act, kl_beta, 1.0, q)
    adam_step(params, grads, state, lr)
    return float(loss)


# ------------------------------------------------------------------------------------------------ framing
def pad_to_multiple(audio: np.ndarray, m: int) -> np.ndarray:
    """Right zero-pad to a multiple of m (rawvae/dataset.py:102-104, 141-143, 61-63)."""
    r = len(audio) % m
    if r != 0:
        audio = np.concatenate([audio, np.zeros(m - r, dtype=audio.dtype)])
    return audio


def audio_dataset_len(n_samples: int, segment_length: int, hop: int) -> int:
    """AudioDataset.__len__ (rawvae/dataset.py:120-121) for an n-sample input."""
    if segment_length % hop != 0:
        raise ValueError("segment_length {} is not a multiple of hop_size {}".format(segment_length, hop))
    padded = -(-n_samples // hop) * hop
    return padded // hop - segment_length // hop + 1


def audio_dataset_frames(audio: np.ndarray, segment_length: int, hop: int,
                         index: Optional[Sequence[int]] = None) -> np.ndarray:
    """Frames of AudioDataset (rawvae/dataset.py:92-121): frame i = pad[i*hop : i*hop + S]."""
    n = audio_dataset_len(len(audio), segment_length, hop)
    pad = pad_to_multiple(np.asarray(audio), hop)
    idx = range(n) if index is None else index
    return np.stack([pad[i * hop: i * hop + segment_length] for i in idx]) if len(idx) else \
        np.zeros((0, segment_length), dtype=pad.dtype)


def test_dataset_frames(audio: np.ndarray, segment_length: int) -> np.ndarray:
    """Frames of TestDataset (rawvae/dataset.py:135-160): pad to a multiple of S, non-overlapping frames."""
    pad = pad_to_multiple(np.asarray(audio), segment_length)
    return pad.reshape(-1, segment_length)


test_dataset_frames.__test__ = False  # not a pytest test


def iterable_file_frames(audio: np.ndarray, hop: int, segment_length: int = 1024) -> np.ndarray:
    """Frames IterableAudioDataset.process_data yields for one (mono, already resampled) file
    (rawvae/dataset.py:61-69): pad to a multiple of hop; starts 0, hop, ..., len - S. S is hard-coded to 1024
    in the reference (:66)."""
    pad = pad_to_multiple(np.asarray(audio), hop)
    starts = range(0, len(pad) - segment_length + 1, hop)
    return np.stack([pad[i:i + segment_length] for i in starts]) if len(starts) else \
        np.zeros((0, segment_length), dtype=pad.dtype)


def iterable_stream(files: Sequence[np.ndarray], hop: int, n_frames: int, segment_length: int = 1024) -> np.ndarray:
    """First n_frames of the endless stream chain(map(process_data, cycle(files))) (rawvae/dataset.py:77-78)."""
    out: List[np.ndarray] = []
    got = 0
    if not any(len(iterable_file_frames(f, hop, segment_length)) for f in files):
        raise ValueError("no file yields a frame")
    while got < n_frames:
        for f in files:
            fr = iterable_file_frames(f, hop, segment_length)
            out.append(fr)
            got += len(fr)
            if got >= n_frames:
                break
    return np.concatenate(out)[:n_frames]


def resynth_concat(frames: np.ndarray) -> np.ndarray:
    """frames.view(-1) (train_iterable.py:246; tutorial.ipynb:543,932,1289)."""
    return np.asarray(frames).reshape(-1)


def resynth_overlap_add(frames: np.ndarray, hop: int) -> np.ndarray:
    """PARITY UNPINNED (no reference implementation): sum of frames at stride hop divided by the per-sample
    overlap count. Defining identity: resynth_overlap_add(audio_dataset_frames(x, S, hop), hop) == pad(x)."""
    frames = np.asarray(frames, dtype=np.float64)
    n, S = frames.shape
    out = np.zeros((n - 1) * hop + S if n else 0)
    cnt = np.zeros_like(out)
    for i in range(n):
        out[i * hop:i * hop + S] += frames[i]
        cnt[i * hop:i * hop + S] += 1
    return out / np.maximum(cnt, 1)


# ------------------------------------------------------------------------------------------------ synthetic data
def synth_wav(rng: np.random.Generator, n_samples: int, sr: int = 44100) -> np.ndarray:
    """0.5*sin(2*pi*f*t + phi) + 0.05*N(0,1), f log-uniform in [55, 7040] Hz, clipped, as the
    float32 value of 16-bit PCM (int16 / 32768 - what torchaudio / librosa / soundfile all decode to)."""
    f = math.exp(rng.uniform(math.log(55.0), math.log(7040.0)))
    phi = rng.uniform(0, 2 * math.pi)
    t = np.arange(n_samples) / sr
    x = 0.5 * np.sin(2 * math.pi * f * t + phi) + 0.05 * rng.standard_normal(n_samples)
    pcm = np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)
    return (pcm.astype(np.float32) / 32768.0).astype(np.float32)
