"""one_peace_b200 — B200-native (sm_100a) implementation of the ONE-PEACE encoder / contrastive /
optimizer hot path behind the reference's fairseq model / criterion / optimizer API.

Importing the package does not load the CUDA extension; the first kernel call does and fails loudly
if ``csrc/libonepeace_b200.so`` is missing.
"""
__version__ = "0.1.0"
