"""Drop-in for the reference's rawvae/dataset.py plus the GPU-resident fast path.

The four public classes keep the reference's constructor signatures and `__len__/__getitem__/__iter__` semantics
bit-exactly (rawvae/dataset.py:11-160), so they still work under torch.utils.data.DataLoader on CPU numpy - that
is indexing, not arithmetic, and it is what the reference's own callers do. The hot path does not go through
them: `GpuFrameLoader` / `GpuFrameStream` keep the wav samples in HBM and emit `FrameBatch` index descriptors that
the framing kernel (rvae_frame_gather) turns into fc1's bf16 operand directly.
"""
from __future__ import annotations

import os
import pathlib
import random
import threading
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

    def gpu_stream(self, batch_size: int, device=None, pcm16: bool = False, **kwargs) -> "GpuFrameStream":
        """The same frame stream, batched on the GPU (see GpuFrameStream; kwargs: rank, world, cache_bytes, lookahead)."""
        return GpuFrameStream(self, batch_size, device or self.device, pcm16=pcm16, **kwargs)


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


def resident_audio(audio_np, device) -> torch.Tensor:
    """The corpus as it lives in HBM. int16 input stays int16. A float array whose every sample is k / 32768 with k an
    int16 - what decoding a 16-bit PCM wav yields (librosa / torchaudio / soundfile agree) - is
    stored as that int16: half the footprint and half the bytes the framing kernel reads, and LOSSLESS, because the
    kernel's int16 * (1 / 32768) reproduces the float exactly (checked here, sample by sample, before it is relied on).
    Anything else (float wavs, resampled audio) stays float32. RVAE_PCM16_RESIDENT=0 turns the detection off."""
    a = np.ascontiguousarray(audio_np)
    if a.dtype == np.int16:
        return torch.from_numpy(a).to(device)
    a = a.astype(np.float32, copy=False)
    if os.environ.get("RVAE_PCM16_RESIDENT", "1") not in ("0", ""):
        q = a * np.float32(32768.0)
        k = np.rint(q)
        if a.size and np.array_equal(q, k) and float(k.min()) >= -32768.0 and float(k.max()) <= 32767.0:
            return torch.from_numpy(k.astype(np.int16)).to(device)
    return torch.from_numpy(a).to(device)


class GpuFrameLoader:
    """Map-style batches on the GPU: the (padded) audio array lives in HBM once; every epoch draws the sampler
    permutation on the host exactly as DataLoader(shuffle=True) does, ships the indices (8 B/frame), and yields
    FrameBatch descriptors. Under data parallelism rank r takes rows shard_bounds(B, r, W) of every global batch."""

    def __init__(self, audio_np, n_frames: int, hop: int, segment_length: int, batch_size: int, shuffle: bool,
                 device="cuda", drop_last: bool = False, rank: int = 0, world: int = 1):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuFrameLoader needs a CUDA device (no CPU fallback)")
        self.audio = resident_audio(audio_np, self.device)
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
    """IterableAudioDataset's endless stream, batched on the GPU through a bounded ingest ring.

    Host side: a worker thread decodes the next files (wav -> mono channel 0 -> resample if needed -> zero-pad to a
    multiple of hop, rawvae/dataset.py:47-63) into PINNED buffers while the GPU trains; each file is then copied
    (PCM16 on the wire when `pcm16`, else float32) on a copy stream into a device RING of `cache_bytes` - the only HBM
    the corpus ever occupies, so corpora larger than HBM (or than the ring) stream through, and a file that is
    still resident when the stream cycles back to it is not uploaded again. Device side: a batch is ONE FrameBatch
    over the ring - `first_frame` when its frames are one run of one file, else an int64 index list built on the
    device - so batches straddle file boundaries exactly as the reference's DataLoader(batch_size=B, shuffle=False)
    over the iterable dataset does (train_iterable.py:143-151), the framing kernel converts int16 -> bf16 while it
    gathers, and every batch has the same input signature (one CUDA graph serves the whole stream).

    A ring region is overwritten only after the steps that read it have run: every batch records which files it
    touches, an event is recorded on the consumer's stream each time the consumer comes back for another batch, and
    the copy stream waits for the event that covers the last reader of the region it is about to reuse."""

    LAG = 3   # a batch handed out at yield j has been enqueued by the consumer once yield j + LAG is requested

    def __init__(self, dataset: IterableAudioDataset, batch_size: int, device, pcm16: bool = False,
                 rank: int = 0, world: int = 1, cache_bytes: int = 2 << 30, lookahead: int = 2):
        self.ds, self.batch_size, self.device = dataset, int(batch_size), torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuFrameStream needs a CUDA device (no CPU fallback)")
        # pcm16="auto": 16-bit PCM on the wire and in the ring when EVERY file of the folder is 16-bit PCM at the target
        # rate (lossless: int16 / 32768 is what the float decode yields), float32 otherwise
        self.pcm16, self.rank, self.world = pcm16, rank, world
        if pcm16 == "auto":
            heads = [audio_io.open_pcm16(f) for f in dataset.audio_file_list]
            self.pcm16 = bool(heads) and all(d is not None and sr == dataset.sampling_rate for d, sr in heads)
        self.pcm16 = bool(self.pcm16)
        self.dtype = torch.int16 if self.pcm16 else torch.float32
        self.esize = 2 if self.pcm16 else 4
        self.capacity = max(int(cache_bytes) // self.esize // 1024 * 1024, 1024)   # samples
        self.lookahead = max(1, int(lookahead))
        self.ring: Optional[torch.Tensor] = None
        self.stats = {"files_uploaded": 0, "bytes_uploaded": 0, "resident_hits": 0, "ring_wraps": 0,
                      "pcm16_passthrough": 0, "pinned_allocs": 0}
        self._free = []                      # [(pinned buffer, event of the last H2D copy that read it or None)]
        self._free_lock = threading.Lock()

    # ---- host side
    def _pinned(self, n: int) -> torch.Tensor:
        """A pinned staging buffer of >= n samples from the pool (allocating pinned memory costs milliseconds; the
        pool holds lookahead + 2 buffers in steady state). A buffer is reused only after the copy that read it."""
        with self._free_lock:
            k = next((i for i, (b, _) in enumerate(self._free) if b.numel() >= n), None)
            buf, ev = self._free.pop(k) if k is not None else (None, None)
        if buf is None:
            buf = torch.empty(max(n, 1) + n // 8, dtype=self.dtype).pin_memory()
            self.stats["pinned_allocs"] += 1
        elif ev is not None:
            ev.synchronize()
        return buf

    def _release(self, buf: torch.Tensor, ev) -> None:
        with self._free_lock:
            self._free.append((buf, ev))
            if len(self._free) > self.lookahead + 3:          # drop the smallest when files of odd sizes pile up
                self._free.sort(key=lambda t: t[0].numel())
                self._free.pop(0)

    def _decode(self, path):
        """(view of a pinned staging buffer holding the file's samples in the wire format, zero-padded to a multiple
        of hop, rawvae/dataset.py:47-63; the buffer itself). 16-bit PCM files at the target rate are copied from the
        memory-mapped file into pinned memory as they are (channel 0 of multi-channel files, :54-55)."""
        hop = self.ds.hop_size
        if self.pcm16:
            data, sr = audio_io.open_pcm16(path)
            if data is not None and sr == self.ds.sampling_rate:
                n = data.shape[0]
                n_pad = (n + hop - 1) // hop * hop
                buf = self._pinned(n_pad)
                view = buf[:n_pad]
                dst = view.numpy()
                np.copyto(dst[:n], data if data.ndim == 1 else data[:, 0])
                dst[n:] = 0
                self.stats["pcm16_passthrough"] += 1
                return view, buf
        audio = self.ds.load_file(path)                       # float32, zero-padded to a multiple of hop
        if self.pcm16:
            audio = torch.round(audio * 32768.0).clamp_(-32768, 32767).to(torch.int16)
        buf = self._pinned(audio.numel())
        view = buf[:audio.numel()]
        view.copy_(audio)
        return view, buf

    # ---- device side
    def _ensure_ring(self, need: int) -> None:
        if need > self.capacity:
            raise RuntimeError(f"a file of {need} samples does not fit the ingest ring ({self.capacity} samples); "
                               f"raise cache_bytes to at least {need * self.esize}")
        if self.ring is None:
            self.ring = torch.zeros(self.capacity, dtype=self.dtype, device=self.device)

    def __iter__(self) -> Iterator[FrameBatch]:
        from concurrent.futures import ThreadPoolExecutor
        files = self.ds.shuffled_data_list if self.ds.shuffle else self.ds.audio_file_list
        if not files:
            raise RuntimeError("no wav files in {}".format(self.ds.audio_folder))
        if self.world > 1:   # ONE file order for the group (rank 0's): the ranks shard the same frame stream
            from . import dist as rdist
            files = [Path(f) for f in rdist.agree([str(f) for f in files])]
        S, hop, bs = 1024, self.ds.hop_size, self.batch_size
        dev = self.device
        copy_stream = torch.cuda.Stream(device=dev)
        pool = ThreadPoolExecutor(max_workers=2)   # decode / staging copies release the GIL
        order = cycle(files)
        queue = []          # [(path, decode future or None)] - the next files of the stream, decoded ahead
        regions = {}        # region id -> {off, n, path, last_yield, ready}: live pieces of the ring
        resident = {}       # path -> region id of its newest upload
        events = {}         # k -> event recorded on the consumer's stream when it came back after batch k
        state = {"write": 0, "next_id": 0, "yields": 0}

        def top_up():
            while len(queue) < self.lookahead + 1:
                path = next(order)
                queue.append((path, None if path in resident else pool.submit(self._decode, path)))

        def reclaim(off, n, strict):
            """Make ring[off : off + n) writable: retire the regions it overlaps once their last readers have run.
            Returns False (and changes nothing) when a region cannot be retired yet; raises instead when `strict`."""
            hit = [rid for rid, r in regions.items() if r["off"] < off + n and off < r["off"] + r["n"]]
            wait = []
            for rid in hit:
                r = regions[rid]
                need = r["last_yield"] + self.LAG
                covering = [k for k in events if k >= need]
                if any(t is r for t in touched) or (r["last_yield"] > 0 and not covering):
                    if strict:
                        raise RuntimeError(
                            "GpuFrameStream: the ingest ring is too small - a region would be overwritten while a batch "
                            "that reads it may still be pending; raise cache_bytes (now {} bytes)".format(
                                self.capacity * self.esize))
                    return False
                if r["last_yield"] > 0:
                    wait.append(events[min(covering)])
            for rid in hit:
                r = regions.pop(rid)
                if resident.get(r["path"]) == rid:
                    del resident[r["path"]]
            for ev in wait:
                copy_stream.wait_event(ev)
            return True

        def place(path, staged, strict=True):
            host, buf = staged
            n = host.numel()
            self._ensure_ring(n)
            off = 0 if state["write"] + n > self.capacity else state["write"]
            if not reclaim(off, n, strict):
                return None
            if off == 0 and state["write"] != 0:
                self.stats["ring_wraps"] += 1
            with torch.cuda.stream(copy_stream):
                self.ring[off:off + n].copy_(host, non_blocking=True)    # pinned host -> ring, off the step's stream
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            self._release(buf, ready)
            rid = state["next_id"]
            state["next_id"] += 1
            regions[rid] = {"off": off, "n": n, "path": path, "last_yield": 0, "ready": ready}
            resident[path] = rid
            state["write"] = off + n
            self.stats["files_uploaded"] += 1
            self.stats["bytes_uploaded"] += n * self.esize
            return regions[rid]

        def upload_ahead():
            """The next file's host->device copy is issued as soon as its decode is done and its ring space is free,
            while the batches of the current file are still being consumed."""
            if queue and queue[0][1] is not None and queue[0][1].done() and queue[0][0] not in resident:
                path, fut = queue[0]
                r = place(path, fut.result(), strict=False)
                if r is not None:
                    queue[0] = (path, None)
                    ahead[path] = r["ready"]

        runs, have, touched, waits, ahead = [], 0, [], [], {}
        try:
            while True:
                top_up()
                path, fut = queue.pop(0)
                rid = resident.get(path)
                if rid is not None:
                    r = regions[rid]                 # uploaded ahead, or still in the ring from an earlier pass
                    if path in ahead:
                        waits.append(ahead.pop(path))
                    else:
                        self.stats["resident_hits"] += 1
                else:
                    r = place(path, fut.result() if fut is not None else self._decode(path))
                    waits.append(r["ready"])
                n_frames = (r["n"] - S) // hop + 1 if r["n"] >= S else 0
                f0 = r["off"] // hop                  # files are padded to a multiple of hop: offsets stay hop-aligned
                start = 0
                while start < n_frames:
                    take = min(n_frames - start, bs - have)
                    runs.append((f0 + start, take))
                    touched.append(r)
                    have += take
                    start += take
                    if have == bs:
                        main = torch.cuda.current_stream(dev)
                        for ev in waits:              # the uploads this batch reads
                            main.wait_event(ev)
                        waits = []
                        state["yields"] += 1
                        for t in touched:
                            t["last_yield"] = state["yields"]
                        batch = self._batch(runs, hop, S)
                        runs, have, touched = [], 0, []
                        upload_ahead()
                        yield batch
                        # The consumer is back for more. With a one-batch lookahead (trainer._with_next) the gather of
                        # batch k has been enqueued by the time it returns after batch k + 1; LAG = 3 leaves margin.
                        ev = torch.cuda.Event()
                        ev.record(torch.cuda.current_stream(dev))
                        events[state["yields"]] = ev
                        events.pop(state["yields"] - 256, None)
        finally:
            pool.shutdown(wait=False, cancel_futures=True)

    def _batch(self, runs, hop: int, S: int) -> FrameBatch:
        """One FrameBatch over the ring for the global batch `runs` = [(first ring frame, count)], sharded by rank."""
        bs = self.batch_size
        lo, hi = shard_bounds(bs, self.rank, self.world) if self.world > 1 else (0, bs)
        rows = dict(global_row0=lo, global_batch=bs) if self.world > 1 else {}
        out, pos = [], 0
        for f0, cnt in runs:                          # this rank's rows [lo, hi) of the global batch
            a, b = max(lo, pos), min(hi, pos + cnt)
            if b > a:
                out.append((f0 + (a - pos), b - a))
            pos += cnt
        if len(out) == 1:
            return FrameBatch(self.ring, out[0][1], hop, S, first_frame=out[0][0], **rows)
        idx = torch.cat([torch.arange(f, f + c, dtype=torch.int64, device=self.device) for f, c in out])
        return FrameBatch(self.ring, hi - lo, hop, S, frame_idx=idx, **rows)
