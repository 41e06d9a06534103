"""rawaudiovae_kelsey_b200 - B200-native (sm_100a) hot path for the rawaudiovae frame-level VAE.

Host-side Python mirrors the reference's `rawvae` API; all arithmetic runs in hand-written CUDA kernels behind
the C ABI in include/rvae_b200.h (librvae_b200.so). There is no CPU fallback.
"""
__version__ = "0.1.0"
