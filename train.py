# -*- coding: utf-8 -*-
"""Epoch trainer - same CLI, ini schema and artefacts as the reference's train.py (train.py:32-307):

    python train.py --config ./default.ini
    torchrun --nproc-per-node 8 train.py --config ./default.ini      # data parallel, one process per GPU

All arithmetic runs in the sm_100a kernels behind rawvae.model / rawvae.dataset (rawaudiovae_kelsey_b200)."""
import sys

from rawaudiovae_kelsey_b200.trainer import run_epoch_trainer

if __name__ == "__main__":
    sys.exit(run_epoch_trainer())
