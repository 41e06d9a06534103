"""Drop-in for the reference's rawvae/tests.py (a helper module, not a test file)."""
from rawaudiovae_kelsey_b200.testaudio import init_test_audio  # noqa: F401
