"""sydr_b200 -- B200-native PCPS acquisition and E/P/L tracking behind SyDR's Python API.

Host side mirrors the reference's operator interface (sydr/dsp, sydr/signal, sydr/channel);
the arithmetic lives in libsydr_b200.so (hand-written sm_100a CUDA, C ABI in
include/sydr_b200.h).  There is no CPU fallback.
"""
from ._lib import SydrError, load  # noqa: F401

__version__ = "0.1.0"
