"""scvae_b200: B200-native engine for the scVAE training/evaluation hot path.

The compute path lives in ``libscvae_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/scvae_b200.h``); this package is the thin host shell that mirrors the reference's
model interface (``VariationalAutoencoder`` / ``GaussianMixtureVariationalAutoencoder``).
"""

__version__ = "0.1.0"
