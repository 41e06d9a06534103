"""Drop-in for the reference's rawvae/model.py (same public names)."""
from rawaudiovae_kelsey_b200.model import VAE, loss_function, FusedTrainStep, FrameBatch  # noqa: F401
