"""Channel base class with the interface of sydr/channel/channel.py (Channel, ChannelStatus).

Same constructor signature, attributes and packet helpers.  The reference makes every channel a
`multiprocessing.Process` that the manager lock-steps with two Events per millisecond
(channel.py:21,121-160); a CUDA context does not survive `fork`, and the batched dispatcher
(`channelManager.ChannelManager`) runs all channels in one GPU launch anyway, so here a channel
is a plain in-process object: `start()` only marks it running, `run()` executes one tick.
"""
from __future__ import annotations

import threading
from abc import ABC, abstractmethod

from ..signal.rfsignal import RFSignal
from ..utils.circularbuffer import CircularBuffer
from ..utils.enumerations import ChannelMessage, ChannelState, TrackingFlags


class Channel(ABC):
    TIMEOUT = 100000          # kept for interface parity (channel.py:26)

    @abstractmethod
    def __init__(self, cid: int, sharedBuffer: CircularBuffer, resultQueue, rfSignal: RFSignal, configuration: dict):
        self.name = f'CID{cid}'
        self.daemon = True
        self.configuration = configuration
        self.channelID = cid
        self.channelState = ChannelState.IDLE
        self.satelliteID = 0
        self.rfBuffer = sharedBuffer
        self.resultQueue = resultQueue
        self.eventRun = threading.Event()
        self.eventDone = threading.Event()
        self.currentSample = 0
        self.rfSignal = rfSignal
        self.tow = 0
        self.week = 0
        self.codeSinceTOW = 0
        self._started = False

    # ---- multiprocessing.Process look-alikes (no fork) ----------------------------------------
    def start(self):
        self._started = True

    def is_alive(self):
        return self._started

    def join(self, timeout=None):
        return None

    # --------------------------------------------------------------------------------------------
    def setSatellite(self, satelliteID: int):
        """channel.py:104-119."""
        self.satelliteID = satelliteID
        self.channelState = ChannelState.ACQUIRING

    def run(self):
        """One tick of channel.py:121-160, in-process: process the buffer according to the channel
        state, append the channel update, hand the packets to the result queue (if any)."""
        results = self._processHandler()
        results.append(self.prepareChannelUpdate())
        if self.resultQueue is not None:
            self.resultQueue.put(results)
        self.eventDone.set()
        return results

    @abstractmethod
    def _processHandler(self):
        return

    def prepareResults(self):
        return {"cid": self.channelID}

    def prepareChannelUpdate(self):
        """channel.py:211-228."""
        _packet = self.prepareResults()
        _packet['type'] = ChannelMessage.CHANNEL_UPDATE
        _packet['state'] = self.channelState
        _packet['tracking_flags'] = self.trackFlags
        _packet['tow'] = self.tow
        _packet['time_since_tow'] = self.getTimeSinceTOW()
        _packet['unprocessed_samples'] = self.rfBuffer.getNbUnreadSamples(self.currentSample)
        _packet['code_since_tow'] = self.codeSinceTOW
        return _packet


class ChannelStatus(ABC):
    """channel.py:232-264."""

    def __init__(self, channelID: int, satelliteID: int):
        self.channelID = channelID
        self.satelliteID = satelliteID
        self.channelState = ChannelState.IDLE
        self.trackFlags = TrackingFlags.UNKNOWN
        self.week = 0
        self.tow = 0
        self.timeSinceTOW = 0
        self.subframeFlags = []
        self.unprocessedSamples = 0
        self.isTOWDecoded = False
