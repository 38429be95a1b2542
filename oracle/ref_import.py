"""Import the real reference (read-only at /root/reference) for fixture generation.

TEST INFRASTRUCTURE ONLY (see oracle/sydr_oracle.py).  Used by
``tests/golden/make_golden.py`` in the build container; ``/root/reference`` does not exist
on the GPU box, so nothing reachable from ``pytest -m gpu``, ``smoke()`` or ``bench.py``
imports this module.

Two modules the reference imports but never uses on this path are stubbed
(SURVEY.md §8c): ``matplotlib.pyplot`` (sydr/dsp/acquisition.py:3) and ``gps_time``
(sydr/utils/time.py:4).  No reference file is modified or copied.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SYDR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sydr", "dsp"))


def _stub_modules():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            p = types.ModuleType("matplotlib.pyplot")
            m.pyplot = p
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = p
    if "gps_time" not in sys.modules:
        try:
            import gps_time  # noqa: F401
        except Exception:
            g = types.ModuleType("gps_time")

            class GPSTime:  # minimal stand-in, never exercised on the DSP path
                def __init__(self, week_number=0, time_of_week=0.0):
                    self.week_number = week_number
                    self.time_of_week = time_of_week

                @classmethod
                def from_datetime(cls, dt):
                    return cls(0, 0.0)

            g.GPSTime = GPSTime
            sys.modules["gps_time"] = g


def load():
    """Returns a namespace with the reference's hot-path callables."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _stub_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    from sydr.dsp import acquisition as acq
    from sydr.dsp import tracking as trk
    from sydr.signal import gnsssignal as gs
    from sydr.signal import ca
    ns.PCPS = acq.PCPS
    ns.TwoCorrelationPeakComparison = acq.TwoCorrelationPeakComparison
    ns.EPL = trk.EPL
    ns.generateReplica = trk.generateReplica
    ns.getCorrelator = trk.getCorrelator
    ns.LoopFiltersCoefficients = trk.LoopFiltersCoefficients
    ns.DLL_NNEML = trk.DLL_NNEML
    ns.PLL_costa = trk.PLL_costa
    ns.BorreLoopFilter = trk.BorreLoopFilter
    ns.GenerateGPSGoldCode = gs.GenerateGPSGoldCode
    ns.UpsampleCode = gs.UpsampleCode
    ns.getSamplesPerCode = gs.getSamplesPerCode
    ns.ca = ca
    return ns


def load_channel():
    """The live Borre channel class + its collaborators, for in-process driving
    (buf.shift(1 ms); ch._processHandler()) without fork, GUI or database."""
    load()
    ns = types.SimpleNamespace()
    from sydr.channel.channel_l1ca_borre import ChannelL1CA
    from sydr.utils.circularbuffer import CircularBuffer
    from sydr.signal.rfsignal import RFSignal
    from sydr.utils.enumerations import ChannelMessage, ChannelState, TrackingFlags
    ns.TrackingFlags = TrackingFlags
    ns.ChannelL1CA = ChannelL1CA
    ns.CircularBuffer = CircularBuffer
    ns.RFSignal = RFSignal
    ns.ChannelMessage = ChannelMessage
    ns.ChannelState = ChannelState
    return ns


def load_channel_kaplan():
    """The Kaplan channel variant (sydr/channel/channel_l1ca_kaplan.py) for in-process driving."""
    ns = load_channel()
    from sydr.channel.channel_l1ca_kaplan import ChannelL1CA_Kaplan
    from sydr.utils.enumerations import LoopLockState, TrackingFlags
    ns.ChannelL1CA_Kaplan = ChannelL1CA_Kaplan
    ns.LoopLockState = LoopLockState
    ns.TrackingFlags = TrackingFlags
    return ns


def load_database():
    """The reference's SQLite result sink (sydr/io/database.py) for fixture generation.  Its module
    imports ``pymap3d`` through sydr/utils/coordinate.py without using it on this path: stubbed."""
    load()
    if "pymap3d" not in sys.modules:
        try:
            import pymap3d  # noqa: F401
        except Exception:
            sys.modules["pymap3d"] = types.ModuleType("pymap3d")
    ns = types.SimpleNamespace()
    from sydr.io.database import DatabaseHandler
    from sydr.utils.enumerations import ChannelMessage
    ns.DatabaseHandler = DatabaseHandler
    ns.ChannelMessage = ChannelMessage
    return ns
