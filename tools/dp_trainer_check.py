"""torchrun --nproc-per-node 2 tools/dp_trainer_check.py --config x.ini [--slow 3] [--dump path]

train_iterable.py's drop-in under data parallelism with rank 0's checkpoint writes made artificially slow (torch.save
sleeps `--slow` seconds on rank 0 only): the other ranks must simply wait for it - in the host barrier after the
checkpoint block and, if they get that far, in the gradient exchange, which waits for minutes and never traps. Every
rank then dumps its final flat weights to <dump>.rank<r> so the caller can check the replicas stayed identical."""
import argparse
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rawaudiovae_kelsey_b200 import trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", required=True)
ap.add_argument("--slow", type=float, default=3.0)
ap.add_argument("--dump", default="")
ap.add_argument("--which", default="stream", choices=["stream", "epoch"])
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))

if rank == 0 and args.slow > 0:
    _save = torch.save

    def slow_save(*a, **k):
        time.sleep(args.slow)
        return _save(*a, **k)
    torch.save = slow_save

models = []
_VAE = trainer.VAE


def recording_vae(*a, **k):
    m = _VAE(*a, **k)
    models.append(m)
    return m


trainer.VAE = recording_vae
run = trainer.run_stream_trainer if args.which == "stream" else trainer.run_epoch_trainer
rc = run(["--config", args.config])
torch.cuda.synchronize()
if args.dump:
    torch.save(models[-1]._flat.params.detach().cpu(), f"{args.dump}.rank{rank}")
sys.exit(rc or 0)
