"""snfft_b200 -- B200-native S_n Fourier transform (drop-in for the hot path of horacepan/SnFFT).

Layout: ``csrc/`` CUDA kernels + plan builder + C ABI (``include/snfft_b200.h``); ``_native.py``
ctypes binding; ``plan.py`` tensor API; ``fft.py`` / ``yor.py`` / ``young_tableau.py`` /
``perm2.py`` / ``perm.py`` / ``utils.py`` mirror the reference's module names and signatures.

The native library is loaded on first use of ``Plan`` / ``get_plan`` (so that
``python -m snfft_b200.build`` can run before it exists); there is no fallback when it is missing.
"""
__all__ = ["Plan", "get_plan", "launch_count", "timing_enable", "timing_read"]


def __getattr__(name):
    if name in __all__:
        from . import plan as _plan
        return getattr(_plan, name)
    raise AttributeError(f"module 'snfft_b200' has no attribute {name!r}")
