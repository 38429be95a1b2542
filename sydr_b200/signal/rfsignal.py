"""IQ recording reader with the interface of sydr/signal/rfsignal.py (RFSignal).

Same constructor dictionary, attributes (`samplingFrequency`, `interFrequency`, `isComplex`,
`fileDataType`, `dtype`, `samplesPerMs`) and methods (`getMilliseconds`, `readFile`,
`readFileBySamples`, `closeFile`, `getCurrentSampleIndex`).  The difference is what happens to
the bytes: the reference widens every int8/int16 I,Q pair to complex128 on the host
(rfsignal.py:127-130, 16 B/sample).  Here the arrays handed out are `IQBlock`s -- complex128
views for callers that want the reference's values, carrying the untouched interleaved integers
in `.raw`, which is what `ChannelManager.addNewRFData` uploads (2 or 4 B/sample) so that the
samples stay integer all the way into the CUDA kernels.
"""
from __future__ import annotations

import numpy as np


class IQBlock(np.ndarray):
    """complex128 samples + `.raw`: the interleaved I,Q integers they were read from."""

    def __new__(cls, raw: np.ndarray):
        r = np.ascontiguousarray(raw)
        obj = (r[0::2] + 1j * r[1::2]).view(cls)          # rfsignal.py:127-130
        obj.raw = r
        return obj

    def __array_finalize__(self, obj):
        self.raw = None                                    # slices / ufunc results are plain complex data

    def block(self, start: int, stop: int) -> "IQBlock":
        """Samples [start, stop) with their raw integers attached."""
        out = self[start:stop]
        out.raw = self.raw[2 * start:2 * stop] if self.raw is not None else None
        return out


class RFSignal:
    CHUNCK_SIZE_MS = 120      # milliseconds per file read (rfsignal.py:6)

    def __init__(self, configuration: dict):
        self.filepath = str(configuration['filepath'])
        self.samplingFrequency = float(configuration['sampling_frequency'])
        self.isComplex = bool(configuration['is_complex'])       # the reference's cast, quirk included (rfsignal.py:30)
        self.interFrequency = float(configuration['intermediate_frequency'])
        dataSize = int(configuration['data_size'])
        if dataSize == 8:
            self.fileDataType = np.int8
        elif dataSize == 16:
            self.fileDataType = np.int16
        else:
            raise ValueError(f"Data type of {dataSize} bit(s) is not valid.")
        self.dtype = np.complex128 if self.isComplex else self.fileDataType
        self.file_id = None
        self.samplesPerMs = int(self.samplingFrequency * 1e-3)
        self.chunck = np.empty((1, self.CHUNCK_SIZE_MS * self.samplesPerMs))
        self.chunckMsCounter = self.CHUNCK_SIZE_MS

    # -------------------------------------------------------------------------------------------
    def getMilliseconds(self, nbMilliseconds: int):
        """rfsignal.py:58-88: next `nbMilliseconds` of data (a divisor of CHUNCK_SIZE_MS)."""
        if self.CHUNCK_SIZE_MS % nbMilliseconds:
            raise ValueError(f"The number of millisecond requested should be a multiple of the chunck size for "
                             f"optimal read ({nbMilliseconds} not multiple of {self.CHUNCK_SIZE_MS}).")
        if self.chunckMsCounter == self.CHUNCK_SIZE_MS:
            self.chunck = self.readFile(timeLength=self.CHUNCK_SIZE_MS, keep_open=True)
            self.chunckMsCounter = 0
        start = self.chunckMsCounter * self.samplesPerMs
        stop = start + self.samplesPerMs * nbMilliseconds
        self.chunckMsCounter += nbMilliseconds
        if isinstance(self.chunck, IQBlock):
            return self.chunck.block(start, stop)
        return self.chunck[start:stop]

    def _read(self, count: int, offset: int, keep_open: bool):
        fid = open(self.filepath, 'rb') if self.file_id is None else self.file_id
        data = np.fromfile(fid, self.fileDataType, offset=offset, count=count)
        if keep_open:
            self.file_id = fid
        else:
            fid.close()
        return IQBlock(data) if self.isComplex else data

    def readFile(self, timeLength, skip=0, keep_open=False):
        """rfsignal.py:92-132: `timeLength` milliseconds, skipping `skip` samples first."""
        itemsize = np.dtype(self.fileDataType).itemsize
        if self.isComplex:
            return self._read(int(2 * (timeLength * 1e-3) * self.samplingFrequency), int(itemsize * skip * 2), keep_open)
        return self._read(int((timeLength * 1e-3) * self.samplingFrequency), int(itemsize * skip), keep_open)

    def readFileBySamples(self, nb_values, skip=0, keep_open=False):
        """rfsignal.py:136-176."""
        itemsize = np.dtype(self.fileDataType).itemsize
        if self.isComplex:
            return self._read(int(2 * nb_values), int(itemsize * skip * 2), keep_open)
        return self._read(int(nb_values), int(itemsize * skip), keep_open)

    def closeFile(self):
        if self.file_id is None:
            raise Warning("File was already close.")
        self.file_id.close()
        self.file_id = None

    def getCurrentSampleIndex(self):
        if self.file_id is None:
            raise Warning("Signal file not open, cannot return current cursor position.")
        pos = self.file_id.tell()
        return int(pos / 2) if self.isComplex else int(pos)
