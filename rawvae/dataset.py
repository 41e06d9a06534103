"""Drop-in for the reference's rawvae/dataset.py (same public names) plus the GPU loaders."""
from rawaudiovae_kelsey_b200.dataset import (AudioDataset, GpuFrameLoader, GpuFrameStream,  # noqa: F401
                                             IterableAudioDataset, TestDataset, ToTensor)
