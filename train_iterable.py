# -*- coding: utf-8 -*-
"""Streaming trainer - same CLI, ini schema and artefacts as the reference's train_iterable.py
(train_iterable.py:34-329): total_num_frames / batch_size batches over the endless IterableAudioDataset stream.

    python train_iterable.py --config ./kelsey_iterable.ini
    torchrun --nproc-per-node 8 train_iterable.py --config ./kelsey_iterable.ini

All arithmetic runs in the sm_100a kernels behind rawvae.model / rawvae.dataset (rawaudiovae_kelsey_b200)."""
import sys

from rawaudiovae_kelsey_b200.trainer import run_stream_trainer

if __name__ == "__main__":
    sys.exit(run_stream_trainer())
