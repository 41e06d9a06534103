"""wav decode / encode / resample for the host side of the data path.

The reference reads audio with librosa.load(sr=...) (train.py:118-126, rawvae/tests.py:28-36: mono = mean of
channels, resampled to `sr`) or torchaudio.load (rawvae/dataset.py:47: float32 [C, N]) and writes with
soundfile.write (train_iterable.py:247). None of those packages is required here: PCM wavs are decoded with
scipy.io.wavfile (int16 / 32768, the value all three agree on), optional packages are used when importable.
"""
from __future__ import annotations

from pathlib import Path
from typing import Tuple

import numpy as np
import torch


def load_wav_channels(path) -> Tuple[np.ndarray, int]:
    """float32 [channels, samples] in [-1, 1) and the file's sampling rate (torchaudio.load convention)."""
    import scipy.io.wavfile as wavfile
    sr, data = wavfile.read(str(path))
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    x = x[None, :] if x.ndim == 1 else np.ascontiguousarray(x.T)
    return x, int(sr)


def open_pcm16(path) -> Tuple[np.ndarray, int]:
    """Zero-copy view of a 16-bit PCM wav: (int16 memmap [samples] or [samples, channels], sampling rate), or
    (None, sr) when the file is not 16-bit PCM. The streaming ingest copies it straight into pinned memory - no
    float round trip (int16 / 32768 is exact in fp32 and the framing kernel applies it on the GPU)."""
    import scipy.io.wavfile as wavfile
    try:
        sr, data = wavfile.read(str(path), mmap=True)
    except Exception:
        return None, 0
    if data.dtype != np.int16:
        return None, int(sr)
    return data, int(sr)


def resample(audio: torch.Tensor, sr_in: int, sr_out: int) -> torch.Tensor:
    """[C, N] -> [C, N'] (torchaudio.functional.resample when available, polyphase otherwise)."""
    if sr_in == sr_out:
        return audio
    try:
        import torchaudio
        return torchaudio.functional.resample(audio, sr_in, sr_out)
    except Exception:  # pragma: no cover - torchaudio is present in the target image
        from math import gcd
        import scipy.signal
        g = gcd(sr_in, sr_out)
        y = scipy.signal.resample_poly(audio.numpy(), sr_out // g, sr_in // g, axis=-1)
        return torch.from_numpy(np.ascontiguousarray(y.astype(np.float32)))


def load_mono(path, sr: int) -> Tuple[np.ndarray, int]:
    """librosa.load(path, sr=sr) equivalent: mono = mean over channels, resampled to sr, float32 1-D."""
    x, file_sr = load_wav_channels(path)
    mono = x.mean(axis=0, dtype=np.float32) if x.shape[0] > 1 else x[0]
    if file_sr != sr:
        mono = resample(torch.from_numpy(mono)[None, :], file_sr, sr)[0].numpy()
    return np.ascontiguousarray(mono, dtype=np.float32), sr


def write_wav(path, data, sr: int) -> None:
    """soundfile.write(path, data, sr) equivalent for float input: 16-bit PCM wav."""
    import scipy.io.wavfile as wavfile
    x = np.asarray(data, dtype=np.float32).reshape(-1)
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    wavfile.write(str(path), int(sr), pcm)
