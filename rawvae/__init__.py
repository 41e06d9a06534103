"""`rawvae` - the reference's import surface (rawvae.model / rawvae.dataset / rawvae.tests), served by the
B200-native package rawaudiovae_kelsey_b200. The reference ships this as an implicit namespace package
(its rawvae/init.py is empty and mis-named); a real package here keeps `from rawvae.model import VAE` working.
"""
