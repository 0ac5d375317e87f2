"""B200-native hot path of Long-Tail-GAN (ash-shar/Long-Tail-GAN): CUDA kernels behind a C ABI plus the
Python host layer that mirrors the reference's plugin surface (generator / discriminator / sample /
eval_functions / train / test)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
