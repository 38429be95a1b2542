"""Ring-buffer bookkeeping with the semantics of sydr/utils/circularbuffer.py:54-148.

In the reference this object is the 100 ms complex128 ring in shared memory that every channel
process reads.  Here the samples live in HBM (see channel/channelManager.py); this class keeps
the same index arithmetic (idxWrite / size / full, wrap-around slices, unread-sample count) so
that `unprocessed_samples` and `time_since_tow` in the result packets are computed exactly as
the reference computes them.  A host copy of the data is optional (`store=True`, used by the
per-call channel path and by tests)."""
from __future__ import annotations

import numpy as np


class CircularBuffer:
    def __init__(self, size: int, dtype=complex, sharedMemory=None, store: bool = True):
        self.maxSize = int(size)
        self.dtype = dtype
        self.sharedMemory = sharedMemory
        if not store:
            self.buffer = None
        elif sharedMemory is None:
            self.buffer = np.ndarray((1, self.maxSize), dtype=dtype)
        else:
            self.buffer = np.ndarray((1, self.maxSize), dtype=dtype, buffer=sharedMemory.buf)
        self.full = False
        self.idxWrite = 0
        self.idxRead = 0
        self.size = 0

    def shift(self, data):
        """circularbuffer.py:54-80: append `data`; its length must divide the ring size."""
        n = len(data)
        if self.maxSize % n != 0:
            raise ValueError("Data shift need to be a multiple from the max buffer size.")
        if self.buffer is not None:
            self.buffer[:, self.idxWrite:self.idxWrite + n] = data
        self.shiftIdxWrite(n)

    def shiftIdxWrite(self, shift: int):
        """circularbuffer.py:84-100."""
        self.idxWrite += shift
        self.size = self.idxWrite
        if self.full:
            self.idxWrite %= self.maxSize
        else:
            if self.idxWrite >= self.maxSize:
                self.full = True
                self.idxWrite %= self.maxSize
            if self.size > self.maxSize:
                self.size = self.maxSize

    def shiftIdxRead(self, shift: int):
        self.idxRead = (self.idxRead + shift) % self.maxSize

    def getSlice(self, idxStart: int = None, samplesRequired: int = 0):
        """circularbuffer.py:113-137: (1, n) view, or a concatenated copy when the slice wraps."""
        if self.buffer is None:
            raise RuntimeError("this CircularBuffer keeps no host copy of the samples")
        if idxStart is None:
            idxStart = self.idxRead
        idxStop = (idxStart + samplesRequired) % self.maxSize
        if idxStop < idxStart:
            return np.concatenate((self.buffer[:, idxStart:], self.buffer[:, :idxStop]), axis=1)
        return self.buffer[:, idxStart:idxStop]

    def getNbUnreadSamples(self, currentSample: int):
        """circularbuffer.py:141-148."""
        if currentSample <= self.idxWrite:
            return self.idxWrite - currentSample
        return self.maxSize - currentSample + self.idxWrite
