"""scrabble-gan_b200: a from-scratch, B200-native (sm_100a) implementation of the ScrabbleGAN train-step hot path
behind the reference's `src/bigacgan` Python API.  Host code is Python; all arithmetic runs in hand-written CUDA
kernels reached through the C ABI in include/sgan.h (libsgan.so, bound with ctypes).  There is no CPU fallback.

The directory name contains a hyphen (it mirrors the reference repository name), so import it with
    import importlib; sg = importlib.import_module("scrabble-gan_b200")
or through the alias module `scrabble_gan_b200` at the repository root.
"""
__version__ = "0.1.0"

from . import _abi  # noqa: F401  (ctypes binding; loading the .so is deferred until first use)

__all__ = ["_abi", "__version__"]
