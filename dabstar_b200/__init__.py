"""dabstar_b200 — B200-native DAB Mode-I baseband decode path (IQ -> FIC/MSC bits) behind a C ABI.

``dabstar_b200.api`` mirrors the reference's class names over include/dabstar_b200.h; ``dabstar_b200.synth`` is the
bundled synthetic transmitter. Importing the package does not touch CUDA; creating an ``api.Context`` does and
raises when the CUDA library or a GPU is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"
