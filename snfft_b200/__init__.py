"""snfft_b200 -- B200-native S_n Fourier transform (drop-in for the hot path of horacepan/SnFFT).

Layout: ``csrc/`` CUDA kernels + plan builder + C ABI (``include/snfft_b200.h``); ``_native.py``
ctypes binding; ``plan.py`` tensor API; ``fft.py`` / ``yor.py`` / ``young_tableau.py`` /
``perm2.py`` / ``perm.py`` / ``utils.py`` mirror the reference's module names and signatures.
"""
from .plan import Plan, get_plan, launch_count  # noqa: F401

__all__ = ["Plan", "get_plan", "launch_count"]
