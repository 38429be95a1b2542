"""Enumerations shared with the reference's result packets (sydr/utils/enumerations.py):
same member names and values, so packets produced here compare equal field by field."""
from __future__ import annotations

import sqlite3
from enum import Enum, IntEnum, unique


class _NamedEnum(Enum):
    def __str__(self):
        return str(self.name)

    def __conform__(self, protocol):      # sqlite adapter hook, as in the reference (enumerations.py:28-35)
        if protocol is sqlite3.PrepareProtocol:
            return str(self.name)


@unique
class GNSSSystems(_NamedEnum):
    UNKNOWN = 0
    GPS = 1
    GLONASS = 2
    GALILEO = 3
    BEIDOU = 4
    QZSS = 5
    IRNSS = 6
    SBAS = 7


@unique
class GNSSSignalType(_NamedEnum):
    GPS_L1_CA = 0

    def __str__(self):
        return str(self.name).replace("_", " ")


@unique
class ChannelState(_NamedEnum):
    """Channel state machine, sydr/utils/enumerations.py:100-112."""
    OFF = 0
    IDLE = 1
    ACQUIRING = 2
    TRACKING = 3


@unique
class ChannelMessage(_NamedEnum):
    """Packet types, sydr/utils/enumerations.py:117-129."""
    END_OF_PIPE = 0
    CHANNEL_UPDATE = 1
    ACQUISITION_UPDATE = 2
    TRACKING_UPDATE = 3
    DECODING_UPDATE = 4


@unique
class TrackingFlags(IntEnum):
    """Bit flags, sydr/utils/enumerations.py:134-153."""
    UNKNOWN = 0
    CODE_LOCK = 1
    BIT_SYNC = 2
    SUBFRAME_SYNC = 4
    TOW_DECODED = 8
    EPH_DECODED = 16
    TOW_KNOWN = 32
    EPH_KNOWN = 64
    FINE_LOCK = 128

    def __str__(self):
        return str(self.name)


@unique
class LoopLockState(IntEnum):
    """Carrier loop state of the Kaplan channel, sydr/utils/enumerations.py:143-150."""
    UNKNOWN = 0
    PULL_IN = 1
    WIDE_TRACK = 2
    NARROW_TRACK = 3

    def __str__(self):
        return str(self.name)
