"""B200-native embed -> attack -> extract hot path of image-in-speech watermarking.

All compute runs in hand-written CUDA (sm_100a) inside ``csrc/libwmk.so`` and is reached through
the C ABI declared in ``include/wmk.h``.  There is no CPU fallback: importing a compute module
without the built library, or calling it without a CUDA device, raises.
"""
__version__ = "0.1.0"
