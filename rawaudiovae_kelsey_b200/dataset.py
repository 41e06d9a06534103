"""Drop-in for the reference's rawvae/dataset.py plus the GPU-resident fast path.

The four public classes keep the reference's constructor signatures and `__len__/__getitem__/__iter__` semantics
bit-exactly (rawvae/dataset.py:11-160), so they still work under torch.utils.data.DataLoader on CPU numpy - that
is indexing, not arithmetic, and it is what the reference's own callers do. The hot path does not go through
them: `GpuFrameLoader` / `GpuFrameStream` keep the wav samples in HBM and emit `FrameBatch` index descriptors that
the framing kernel (rvae_frame_gather) turns into fc1's bf16 operand directly.
"""
from __future__ import annotations

import pathlib
import random
from itertools import chain, cycle
from pathlib import Path
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import IterableDataset

from . import audio_io
from .model import FrameBatch


class IterableAudioDataset(IterableDataset):
    """Endless stream of 1024-sample frames over the wav files of a folder (rawvae/dataset.py:11-84).

    Only the file list is shuffled, per __iter__ (:38-42); within a file frames come in order at stride hop_size;
    the segment length is fixed at 1024 as in the reference (:66). Use with DataLoader(shuffle=False)."""

    def __init__(self, audio_folder, sampling_rate, hop_size, dtype, device, shuffle=True):
        self.sampling_rate = sampling_rate
        self.hop_size = hop_size
        self.dtype = dtype
        self.device = device
        self.shuffle = shuffle
        if isinstance(audio_folder, pathlib.PurePath):
            self.audio_folder = audio_folder
        else:
            self.audio_folder = Path(audio_folder)
        self.audio_file_list = [f for f in self.audio_folder.glob('*.wav')]
        self.num_files = len(self.audio_file_list)

    @property
    def shuffled_data_list(self):
        return random.sample(self.audio_file_list, len(self.audio_file_list))

    def load_file(self, audio_file) -> torch.Tensor:
        """Decode -> resample if needed -> channel 0 -> 1-D -> zero-pad to a multiple of hop (:47-63)."""
        audio, sr = audio_io.load_wav_channels(audio_file)           # float32 [C, N]
        audio = torch.from_numpy(audio)
        if sr != self.sampling_rate:
            audio = audio_io.resample(audio, sr, self.sampling_rate)
        if audio.shape[0] > 1:
            audio = audio[0:1, :]
        audio = audio.flatten()
        if len(audio) % self.hop_size != 0:
            num_zeros = self.hop_size - (len(audio) % self.hop_size)
            audio = torch.nn.functional.pad(audio, (0, num_zeros), 'constant')
        return audio

    def process_data(self, audio_file):
        audio = self.load_file(audio_file)
        segment_length = 1024
        on_cuda = getattr(self.device, "type", str(self.device)) == "cuda"
        if on_cuda:
            audio = audio.to(self.device)  # one H2D copy per file instead of one per frame (:72-73)
        for i in range(0, len(audio) - segment_length + 1, self.hop_size):
            yield audio[i:i + segment_length]

    def get_stream(self, audio_file_list):
        return chain.from_iterable(map(self.process_data, cycle(audio_file_list)))

    def __iter__(self):
        if self.shuffle:
            return self.get_stream(self.shuffled_data_list)
        else:
            return self.get_stream(self.audio_file_list)

    def gpu_stream(self, batch_size: int, device=None, pcm16: bool = False) -> "GpuFrameStream":
        """The same frame stream, batched on the GPU (see GpuFrameStream)."""
        return GpuFrameStream(self, batch_size, device or self.device, pcm16=pcm16)


class AudioDataset(torch.utils.data.Dataset):
    """Overlapping frames of one long array: frame i = audio[i*hop : i*hop + S] (rawvae/dataset.py:86-121)."""

    def __init__(self, audio_np, segment_length, sampling_rate, hop_size, transform=None):
        self.transform = transform
        self.sampling_rate = sampling_rate
        self.segment_length = segment_length
        self.hop_size = hop_size
        if segment_length % hop_size != 0:
            raise ValueError("segment_length {} is not a multiple of hop_size {}".format(segment_length, hop_size))
        if len(audio_np) % hop_size != 0:
            num_zeros = hop_size - (len(audio_np) % hop_size)
            audio_np = np.pad(audio_np, (0, num_zeros), 'constant', constant_values=(0, 0))
        self.audio_np = audio_np

    def __getitem__(self, index):
        seg_start = index * self.hop_size
        seg_end = (index * self.hop_size) + self.segment_length
        sample = self.audio_np[seg_start: seg_end]
        if self.transform:
            sample = self.transform(sample)
        return sample

    def __len__(self):
        return (len(self.audio_np) // self.hop_size) - (self.segment_length // self.hop_size) + 1

    def gpu_loader(self, batch_size: int, shuffle: bool, device="cuda", drop_last: bool = False) -> "GpuFrameLoader":
        return GpuFrameLoader(self.audio_np, len(self), self.hop_size, self.segment_length, batch_size, shuffle,
                              device, drop_last)


class ToTensor(object):
    """Convert ndarrays in sample to Tensors (rawvae/dataset.py:123-127)."""

    def __call__(self, sample):
        return torch.from_numpy(sample)


class TestDataset(torch.utils.data.Dataset):
    """Non-overlapping frames, zero-padded to a multiple of S (rawvae/dataset.py:129-160)."""
    __test__ = False  # not a pytest class

    def __init__(self, audio_np, segment_length, sampling_rate, transform=None):
        self.transform = transform
        self.sampling_rate = sampling_rate
        self.segment_length = segment_length
        if len(audio_np) % segment_length != 0:
            num_zeros = segment_length - (len(audio_np) % segment_length)
            audio_np = np.pad(audio_np, (0, num_zeros), 'constant', constant_values=(0, 0))
        self.audio_np = audio_np

    def __getitem__(self, index):
        seg_start = index * self.segment_length
        seg_end = (index * self.segment_length) + self.segment_length
        sample = self.audio_np[seg_start: seg_end]
        if self.transform:
            sample = self.transform(sample)
        return sample

    def __len__(self):
        return len(self.audio_np) // self.segment_length

    def gpu_loader(self, batch_size: int, device="cuda") -> "GpuFrameLoader":
        return GpuFrameLoader(self.audio_np, len(self), self.segment_length, self.segment_length, batch_size, False,
                              device, False)


# ---------------------------------------------------------------------------------------------------- GPU fast path
def sampler_seed() -> int:
    """The seed torch's RandomSampler would draw (DataLoader(shuffle=True) without a generator): the iterator first
    draws its _base_seed, then RandomSampler draws seed = int(torch.empty((), dtype=torch.int64).random_().item())
    from the global RNG."""
    torch.empty((), dtype=torch.int64).random_()  # the DataLoader iterator's _base_seed draw comes first
    return int(torch.empty((), dtype=torch.int64).random_().item())


def sampler_permutation(n: int, seed: Optional[int] = None) -> torch.Tensor:
    """torch.randperm(n, generator=Generator().manual_seed(seed)) with seed = sampler_seed(): same global seed => same
    batches as the reference's DataLoader (train.py:134)."""
    gen = torch.Generator()
    gen.manual_seed(sampler_seed() if seed is None else seed)
    return torch.randperm(n, generator=gen)


def shard_bounds(batch: int, rank: int, world: int):
    """Rows [lo, hi) of a global batch owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GpuFrameLoader:
    """Map-style batches on the GPU: the (padded) audio array lives in HBM once; every epoch draws the sampler
    permutation on the host exactly as DataLoader(shuffle=True) does, ships the indices (8 B/frame), and yields
    FrameBatch descriptors. Under data parallelism rank r takes rows shard_bounds(B, r, W) of every global batch."""

    def __init__(self, audio_np, n_frames: int, hop: int, segment_length: int, batch_size: int, shuffle: bool,
                 device="cuda", drop_last: bool = False, rank: int = 0, world: int = 1):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuFrameLoader needs a CUDA device (no CPU fallback)")
        a = np.ascontiguousarray(audio_np)
        if a.dtype == np.int16:
            self.audio = torch.from_numpy(a).to(self.device)
        else:
            self.audio = torch.from_numpy(a.astype(np.float32, copy=False)).to(self.device)
        self.n_frames, self.hop, self.segment_length = int(n_frames), int(hop), int(segment_length)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.rank, self.world = rank, world

    def __len__(self):
        if self.drop_last:
            return self.n_frames // self.batch_size
        return (self.n_frames + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[FrameBatch]:
        n, bs = self.n_frames, self.batch_size
        if self.shuffle:
            # every rank draws a seed (its global RNG advances like the reference's DataLoader would); under data
            # parallelism rank 0's seed wins (8 bytes travel, not the n indices): the shards below are slices of ONE
            # permutation - the single-process one - never overlapping, never missing a frame
            seed = sampler_seed()
            if self.world > 1:
                from . import dist as rdist
                seed = rdist.agree(seed)
            perm = sampler_permutation(n, seed).to(self.device, non_blocking=True)
        for b in range(len(self)):
            lo, hi = b * bs, min((b + 1) * bs, n)
            gb = hi - lo
            if gb < self.world:
                continue   # fewer frames than ranks: dropped on EVERY rank, so no rank skips a collective step alone
            s_lo, s_hi = shard_bounds(gb, self.rank, self.world)
            lo, hi = lo + s_lo, lo + s_hi
            rows = dict(global_row0=s_lo, global_batch=gb) if self.world > 1 else {}
            if self.shuffle:
                yield FrameBatch(self.audio, hi - lo, self.hop, self.segment_length, frame_idx=perm[lo:hi], **rows)
            else:
                yield FrameBatch(self.audio, hi - lo, self.hop, self.segment_length, first_frame=lo, **rows)


class GpuFrameStream:
    """IterableAudioDataset's endless stream, batched on the GPU: each file is decoded once, padded to a multiple
    of hop and uploaded (as PCM16 or float32); a batch is a list of (file buffer, first frame, count) runs that the
    framing kernel gathers into one [B, 1024] operand. Batches straddle file boundaries exactly as the reference's
    DataLoader(batch_size=B, shuffle=False) over the iterable dataset does (train_iterable.py:143-151)."""

    def __init__(self, dataset: IterableAudioDataset, batch_size: int, device, pcm16: bool = False,
                 rank: int = 0, world: int = 1, cache_files: bool = True):
        self.ds, self.batch_size, self.device = dataset, int(batch_size), torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuFrameStream needs a CUDA device (no CPU fallback)")
        self.pcm16, self.rank, self.world = pcm16, rank, world
        self.cache = {} if cache_files else None

    def _file_buffer(self, path) -> torch.Tensor:
        if self.cache is not None and path in self.cache:
            return self.cache[path]
        audio = self.ds.load_file(path)
        if self.pcm16:
            audio = torch.round(audio * 32768.0).clamp_(-32768, 32767).to(torch.int16)
        buf = audio.to(self.device)
        if self.cache is not None:
            self.cache[path] = buf
        return buf

    def __iter__(self) -> Iterator[List[FrameBatch]]:
        files = self.ds.shuffled_data_list if self.ds.shuffle else self.ds.audio_file_list
        if not files:
            raise RuntimeError("no wav files in {}".format(self.ds.audio_folder))
        if self.world > 1:   # ONE file order for the group (rank 0's): the ranks shard the same frame stream
            from . import dist as rdist
            files = [Path(f) for f in rdist.agree([str(f) for f in files])]
        S, hop, bs = 1024, self.ds.hop_size, self.batch_size
        pending: List[FrameBatch] = []
        have = 0
        for path in cycle(files):
            buf = self._file_buffer(path)
            n_frames = (buf.numel() - S) // hop + 1 if buf.numel() >= S else 0
            start = 0
            while start < n_frames:
                take = min(n_frames - start, bs - have)
                pending.append(FrameBatch(buf, take, hop, S, first_frame=start))
                have += take
                start += take
                if have == bs:
                    yield self._shard(pending)
                    pending, have = [], 0

    def _shard(self, runs: List[FrameBatch]) -> List[FrameBatch]:
        if self.world == 1:
            return runs
        lo, hi = shard_bounds(self.batch_size, self.rank, self.world)
        out, pos = [], 0
        for r in runs:
            a, b = max(lo, pos), min(hi, pos + r.n_frames)
            if b > a:
                out.append(FrameBatch(r.audio, b - a, r.hop, r.segment_length, first_frame=r.first_frame + (a - pos),
                                      global_row0=lo, global_batch=self.batch_size))
            pos += r.n_frames
        return out
