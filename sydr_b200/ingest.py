"""Streaming ingest: IQ file -> pinned host buffers -> sliding device window -> acquisition once,
then closed-loop tracking (and navigation bits) chunk by chunk.

Replaces the reference's read path: `RFSignal.readFile` / `getMilliseconds`
(sydr/signal/rfsignal.py:58-132) widens every int8/int16 I,Q pair to complex128 on the host
(16 B/sample) and `CircularBuffer.shift` (sydr/utils/circularbuffer.py:54-105) copies it into a
100 ms shared-memory ring that every channel process slices per millisecond
(sydr/receiver/receiver.py:120-139).  Here the file's own bytes (2 or 4 B/sample) are read by a
small pool of threads straight into pinned host memory, copied to the device on a copy stream
while the previous chunk is being tracked, and never change format.

Device memory is two windows of `tail + chunk` samples used alternately.  Window k holds the
recording's samples [k*chunk - tail, k*chunk + len_k): the last `tail` samples of the previous
window (the epochs still open at a chunk boundary) are copied device-to-device in front of the new
chunk, and the tracking kernel is told where the recording starts relative to the window
(`iq_base = tail - k*chunk`, possibly negative), so channel states and epoch records keep
recording-relative sample indices for any file length.
"""
from __future__ import annotations

import math
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib as L
from .engine import (IQ_PAD_BYTES, AcquisitionEngine, KaplanTrackingEngine, NavBitEngine, TrackingEngine,
                     make_kaplan_states, make_trk_states)


class FileChunkReader:
    """Background reader: consecutive chunks of an interleaved int8/int16 IQ file into rotating
    pinned host buffers.  Each chunk is read by `threads` parallel `os.preadv` calls (they release
    the GIL), which is what it takes to feed a PCIe 5 x16 link from the page cache."""

    def __init__(self, path: str, np_dtype, chunk_samples: int, skip_samples: int = 0, max_samples: int | None = None,
                 n_buffers: int = 3, threads: int = 4, buffers=None, pool=None):
        self.path = path
        self.itemsize = np.dtype(np_dtype).itemsize
        self.bps = 2 * self.itemsize                            # bytes per complex sample
        self.chunk = int(chunk_samples)
        total = os.path.getsize(path) // self.bps - int(skip_samples)
        if max_samples is not None:
            total = min(total, int(max_samples))
        self.total = max(total, 0)
        self.skip = int(skip_samples)
        self.n_chunks = -(-self.total // self.chunk) if self.total else 0
        tdt = torch.int8 if self.itemsize == 1 else torch.int16
        # pinned staging buffers and the read pool may be handed in by a long-lived owner (page-locking
        # hundreds of MB costs tens of milliseconds)
        self._bufs = buffers if buffers is not None else self.alloc_buffers(tdt, self.chunk, n_buffers)
        self._free = queue.Queue()
        for i in range(len(self._bufs)):
            self._free.put(i)
        self._ready = queue.Queue()
        self._threads = max(1, int(threads))
        self._own_pool = pool is None
        self._pool = ThreadPoolExecutor(self._threads) if pool is None else pool
        self._fd = os.open(path, os.O_RDONLY)
        self._stop = False
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    @staticmethod
    def alloc_buffers(tdt, chunk_samples: int, n_buffers: int = 3):
        return [torch.empty(2 * int(chunk_samples), dtype=tdt, pin_memory=torch.cuda.is_available())
                for _ in range(n_buffers)]

    def _read_into(self, view: memoryview, offset: int):
        done = 0
        while done < len(view):
            got = os.preadv(self._fd, [view[done:]], offset + done)
            if got <= 0:
                raise IOError(f"short read from {self.path} at byte {offset + done}")
            done += got

    def _run(self):
        try:
            for k in range(self.n_chunks):
                slot = self._free.get()
                if self._stop or slot is None:
                    return
                lo = k * self.chunk
                n = min(self.chunk, self.total - lo)
                raw = memoryview(self._bufs[slot].numpy()).cast("B")[:n * self.bps]
                off = (self.skip + lo) * self.bps
                step = -(-len(raw) // self._threads)
                step += (-step) % 4096
                futs = [self._pool.submit(self._read_into, raw[a:min(a + step, len(raw))], off + a)
                        for a in range(0, len(raw), step)]
                for f in futs:
                    f.result()
                self._ready.put((k, slot, n))
            self._ready.put(None)
        except BaseException as e:                               # surfaced to the consumer
            self._ready.put(e)

    def __iter__(self):
        while True:
            item = self._ready.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            k, slot, n = item
            yield k, slot, self._bufs[slot][:2 * n], n

    def release(self, slot: int):
        self._free.put(slot)

    def close(self):
        self._stop = True
        self._free.put(None)
        self._worker.join(timeout=5)
        if self._own_pool:
            self._pool.shutdown(wait=False)
        os.close(self._fd)


class StreamingReceiver:
    """File -> acquisition table, per-epoch tracking records and navigation bits, at any length.

    `rf` is an `RFSignal` (sydr_b200.signal.rfsignal, same configuration dictionary as the
    reference's).  The first chunk is searched for `search_prns`; the satellites found are handed
    over with the reference's scalars (channel_l1ca_borre.py:301-311) and tracked to the end of
    the file.  `run()` yields one result dictionary per chunk as soon as the chunk is done."""

    def __init__(self, rf, search_prns, n_channels, chunk_seconds=1.0, doppler_range=5000.0, doppler_step=250.0,
                 coh=1, noncoh=10, threshold=1.5, channel_cfg=None, want_records=True, want_bits=True,
                 reader_threads=8, device=None, cluster=0, threads=0, use_tma=True, loop="borre"):
        L.require_device()
        if not rf.isComplex:
            raise L.SydrError("StreamingReceiver needs interleaved I,Q samples (is_complex = true)")
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.rf = rf
        self.fs = float(rf.samplingFrequency)
        self.np_dtype = np.dtype(rf.fileDataType)
        self.nbits = 8 * self.np_dtype.itemsize
        self._tdt = torch.int8 if self.nbits == 8 else torch.int16
        self.n_channels, self.threshold, self.channel_cfg = int(n_channels), float(threshold), channel_cfg
        self.loop = str(loop)                  # "borre" or "kaplan": which loop closure the tracking kernel runs
        if self.loop not in ("borre", "kaplan"):
            raise L.SydrError(f"unknown loop closure '{loop}'")
        self.want_records, self.want_bits = bool(want_records), bool(want_bits)
        self.reader_threads = int(reader_threads)
        self.trk_cfg = dict(cluster=cluster, threads=threads, use_tma=use_tma)
        self.acq = AcquisitionEngine(self.fs, float(rf.interFrequency), doppler_range, doppler_step, coh, noncoh,
                                     list(search_prns), device=self.device)
        ms = int(self.fs * 1e-3)
        self.chunk = max(int(round(chunk_seconds * self.fs)) // 64 * 64, 64)
        self.tail = -(-(2 * ms + 64) // 64) * 64                 # > one epoch + vector slack, 16-byte phase kept
        if self.chunk < self.acq.required_samples + 4 * self.acq.n_code:
            raise L.SydrError("chunk_seconds shorter than the acquisition dwell")
        self.max_epochs = int(math.ceil(self.chunk / self.fs * 1000.0)) + 8
        pad = IQ_PAD_BYTES // self.np_dtype.itemsize
        self._win = [torch.zeros(2 * (self.tail + self.chunk) + pad, dtype=self._tdt, device=self.device)
                     for _ in range(2)]
        self._copy = torch.cuda.Stream(device=self.device)
        self._host_bufs = FileChunkReader.alloc_buffers(self._tdt, self.chunk, 3)
        self._pool = ThreadPoolExecutor(max(1, self.reader_threads))
        self._rec_host = [None, None]
        self._nep_host = [torch.zeros(self.n_channels, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        self._trk = None
        self._nav = None
        self.stats = {}

    def close(self):
        self.acq.close()
        self._pool.shutdown(wait=False)

    # ------------------------------------------------------------------------------------
    def _start_tracking(self, peaks):
        order = np.argsort(-peaks["ratio"], kind="stable")
        sel = [i for i in order if peaks["ratio"][i] > self.threshold][:self.n_channels]
        sel = sorted(sel, key=lambda i: int(peaks["prn"][i]))
        chans = []
        for i in sel:
            carrier, _, cur = self.acq.handoff(peaks[i])
            chans.append(dict(prn=int(peaks["prn"][i]), carrier_freq=carrier, start_sample=cur, iq_len=0))
        if chans:
            n_ch = len(chans)
            if self.loop == "kaplan":
                states, kstates = make_kaplan_states(self.fs, chans, self.channel_cfg)
                self._trk = KaplanTrackingEngine(self.fs, states, kstates, self.max_epochs, device=self.device,
                                                 **self.trk_cfg)
                self._rec_host = [torch.empty(n_ch * self.max_epochs * 128, dtype=torch.uint8, pin_memory=True)
                                  for _ in range(2)]
                self._krec_host = [torch.empty(n_ch * self.max_epochs * 40, dtype=torch.uint8, pin_memory=True)
                                   for _ in range(2)]
                if self.want_bits:
                    self._nav = NavBitEngine(n_ch, max_bits=self.max_epochs // 20 + 2, device=self.device)
                return chans
            states = make_trk_states(self.fs, chans, self.channel_cfg)
            if self._trk is not None and self._trk.n_ch == n_ch:        # a receiver that is run again keeps its buffers
                self._trk.reset(states)
                if self._nav is not None:
                    self._nav.reset()
                return chans
            self._trk = TrackingEngine(self.fs, states, self.max_epochs, device=self.device, **self.trk_cfg)
            if self.want_bits:
                self._nav = NavBitEngine(n_ch, max_bits=self.max_epochs // 20 + 2, device=self.device)
            self._rec_host = [torch.empty(n_ch * self.max_epochs * 128, dtype=torch.uint8, pin_memory=True)
                              for _ in range(2)]
        return chans

    def _collect(self, k, done_ev, chans):
        """Host side of chunk k: wait for its results and unpack them."""
        done_ev.synchronize()
        out = dict(chunk=k)
        if not chans:
            return out
        n_ch = len(chans)
        nep = self._nep_host[k & 1][:n_ch].numpy().copy()
        out["nepochs"] = nep
        if self.want_records:
            rec = self._rec_host[k & 1].numpy().view(L.TRK_EPOCH_DTYPE).reshape(n_ch, self.max_epochs)
            out["epochs"] = [rec[c, :nep[c]].copy() for c in range(n_ch)]
            if self.loop == "kaplan":
                kr = self._krec_host[k & 1].numpy().view(L.KAPLAN_EPOCH_DTYPE).reshape(n_ch, self.max_epochs)
                out["kaplan"] = [kr[c, :nep[c]].copy() for c in range(n_ch)]
        if self.want_bits:
            out["bits"] = self._unpack_bits(self._bits_pending.pop(k), n_ch)
        return out

    def run(self, skip_samples: int = 0, max_samples: int | None = None):
        """Generator over chunks: {'chunk', 'peaks' and 'channels' (first chunk), 'epochs', 'bits', 'nepochs'}."""
        reader = FileChunkReader(self.rf.filepath, self.np_dtype, self.chunk, skip_samples, max_samples,
                                 threads=self.reader_threads, buffers=self._host_bufs, pool=self._pool)
        comp = torch.cuda.current_stream()
        T, CH = self.tail, self.chunk
        chans, peaks = None, None
        win_free = [None, None]               # event: compute no longer reads window i
        pending = None                        # (k, done event) of the chunk whose results are still on the device
        self._bits_pending = {}
        total = 0
        try:
            for k, slot, host, n in reader:
                w = self._win[k & 1]
                # ---- H2D on the copy stream, behind the last kernels that read this window
                if win_free[k & 1] is not None:
                    self._copy.wait_event(win_free[k & 1])
                with torch.cuda.stream(self._copy):
                    w[2 * T:2 * (T + n)].copy_(host, non_blocking=True)
                    h2d = torch.cuda.Event()
                    h2d.record(self._copy)
                comp.wait_event(h2d)
                if k > 0:                     # open epochs of the previous window: its last `tail` samples
                    prev = self._win[(k - 1) & 1]
                    w[:2 * T].copy_(prev[2 * CH:2 * (CH + T)], non_blocking=True)
                first = None
                if k == 0:
                    self.acq.launch(w[2 * T:2 * (T + n)])
                    peaks = self.acq.fetch()["peaks"]            # 24 B per PRN; the only sync of the hot loop
                    chans = self._start_tracking(peaks)
                    first = dict(peaks=peaks, channels=chans)
                done = torch.cuda.Event()
                if chans:
                    self._trk.launch(w[:2 * (T + CH)], iq_len=k * CH + n, iq_base=T - k * CH)
                    n_ch = len(chans)
                    self._nep_host[k & 1][:n_ch].copy_(self._trk._nep, non_blocking=True)
                    if self.want_bits:
                        # K-NAV on the records just written; its small outputs are snapshotted on the
                        # device because the next chunk's launch reuses them
                        self._nav.launch(self._trk)
                        self._bits_pending[k] = (self._nav._nbits.clone(), self._nav._bits.clone())
                    if self.want_records:
                        self._rec_host[k & 1].copy_(self._trk._out, non_blocking=True)
                        if self.loop == "kaplan":
                            self._krec_host[k & 1].copy_(self._trk._kout, non_blocking=True)
                done.record(comp)
                win_free[k & 1] = done
                # ---- results of the previous chunk while this one runs
                if pending is not None:
                    pk, pev, pfirst = pending
                    res = self._collect(pk, pev, chans)
                    if pfirst:
                        res.update(pfirst)
                    yield res
                pending = (k, done, first)
                # the pinned buffer may be refilled once its H2D has completed
                h2d.synchronize()
                reader.release(slot)
                total += n
            if pending is not None:
                pk, pev, pfirst = pending
                res = self._collect(pk, pev, chans)
                if pfirst:
                    res.update(pfirst)
                yield res
        finally:
            reader.close()
        self.stats = dict(samples=total, chunks=reader.n_chunks)

    def _unpack_bits(self, item, n_ch):
        nb, bits = item
        nb = nb.cpu().numpy()
        b = bits.view(n_ch, -1).cpu().numpy()
        return [b[c, :nb[c]].copy() for c in range(n_ch)]

    def run_all(self, skip_samples: int = 0, max_samples: int | None = None) -> dict:
        """Whole file: peaks, channels, per-channel concatenated epoch records and navigation bits."""
        out = dict(peaks=None, channels=[], epochs=None, bits=None)
        ep, bits = None, None
        for res in self.run(skip_samples, max_samples):
            if "peaks" in res:
                out["peaks"], out["channels"] = res["peaks"], res["channels"]
                n_ch = len(res["channels"])
                ep = [[] for _ in range(n_ch)]
                bits = [[] for _ in range(n_ch)]
            if "epochs" in res:
                for c, e in enumerate(res["epochs"]):
                    ep[c].append(e)
            if "bits" in res:
                for c, b in enumerate(res["bits"]):
                    bits[c].append(b)
            if "kaplan" in res:
                kap = out.setdefault("_kap", [[] for _ in res["kaplan"]])
                for c, kx in enumerate(res["kaplan"]):
                    kap[c].append(kx)
        if ep is not None and self.want_records:
            out["epochs"] = [np.concatenate(e) if e else np.zeros(0, dtype=L.TRK_EPOCH_DTYPE) for e in ep]
        if bits is not None and self.want_bits:
            out["bits"] = [np.concatenate(b) if b else np.zeros(0, dtype=np.int8) for b in bits]
        if "_kap" in out:
            out["kaplan"] = [np.concatenate(kx) for kx in out.pop("_kap")]
        if self._trk is not None:
            out["states"] = self._trk.states()
        return out

    # ------------------------------------------------------------------------------------
    def run_to_database(self, database, skip_samples: int = 0, max_samples: int | None = None, wall_time=None,
                        channel_ids: dict | None = None):
        """Whole file into a `DatabaseHandler` (sydr_b200.io.database, the reference's SQLite format):
        one `channel` row and one `acquisition` row per tracked satellite, one `tracking` row per
        channel-epoch, inserted column-wise per chunk while the next chunk is on the GPU, and (want_bits) one
        `decoding` row per decoded LNAV subframe -- the DECODING_UPDATE packets of the reference's run()
        (channel_l1ca_borre.py:455-573: the K-NAV bits go through the same preamble search / subframe
        synchronisation, `lnav_frame.advance_frame`, with `subframe_id`, `tow` and `bits` as the packet carries them).
        `time_sample` is the receiver's sample counter at the millisecond tick on which the reference
        would have emitted the packet (sydr/receiver/receiver.py:120-139, 357-360); `time` is the
        wall clock (`wall_time()` if given, for reproducible files).  `channel_ids` maps PRN -> channel id
        when the caller has already registered its channels (no `channel` rows are written then).
        Returns run_all()-style totals."""
        import time as _time
        from .io.database import cn0_column
        if not self.want_records:
            raise L.SydrError("run_to_database needs want_records=True")
        from .channel.lnav_frame import advance_frame
        from .utils.constants import LNAV_SUBFRAME_SIZE, LNAV_WORD_SIZE
        from .utils.enumerations import ChannelMessage, TrackingFlags

        class _Frame:                                # what advance_frame works on (the channel classes' members)
            def __init__(self):
                self.navBitBufferSize = LNAV_SUBFRAME_SIZE + 2 * LNAV_WORD_SIZE + 2
                self.navBitsBuffer = np.zeros(self.navBitBufferSize, dtype=int)
                self.navBitsCounter = 0
                self.preambuleFound = False
                self.trackFlags = TrackingFlags.BIT_SYNC

        now = wall_time or _time.time
        spm = int(self.fs * 1e-3)
        chans, done = [], None
        frames, bits_done, last_tick = [], [], []
        rows = decoded_rows = 0
        for res in self.run(skip_samples, max_samples):
            if "peaks" in res:
                chans = res["channels"]
                done = [0] * len(chans)
                frames = [_Frame() for _ in chans]
                bits_done = [0] * len(chans)
                last_tick = [-(1 << 62)] * len(chans)
                dwell = self.acq.required_samples
                cids = [k if channel_ids is None else int(channel_ids[ch["prn"]]) for k, ch in enumerate(chans)]
                for cid, ch in zip(cids, chans):
                    pk = res["peaks"][[int(p) for p in res["peaks"]["prn"]].index(ch["prn"])]
                    if channel_ids is None:
                        database.addData("channel", {"id": cid, "physical_id": cid, "system": "GPS",
                                                     "satellite_id": int(ch["prn"]), "signal": "GPS_L1_CA",
                                                     "start_time": float(now()), "start_sample": 0})
                    database.addData("acquisition", {
                        "cid": cid, "carrierFrequency": float(ch["carrier_freq"]), "codeOffset": int(pk["code_idx"]),
                        "frequency_idx": int(pk["freq_idx"]), "code_idx": int(pk["code_idx"]),
                        "peak_ratio": float(pk["ratio"]), "channel_id": cid, "time": float(now()),
                        "time_sample": int(-(-dwell // spm) * spm)})
                database.commit()
            if "epochs" not in res:
                continue
            sync = self._nav.states()["sync_epoch"] if self._nav is not None else [-1] * len(chans)
            for k, rec in enumerate(res["epochs"]):
                cid = cids[k]
                if not len(rec):
                    continue
                end = (rec["start"] + rec["n"]).astype(np.int64)
                # the receiver ticks once per millisecond and a channel emits at most one epoch per tick
                # (channel.py:121-160): tick_k = max(ceil(end_k / spm) spm, tick_{k-1} + spm), a running maximum
                tick = -(-end // spm) * spm
                k_idx = spm * (done[k] + np.arange(len(rec), dtype=np.int64))
                run = np.maximum.accumulate(np.concatenate(([last_tick[k]], tick - k_idx)))[1:]
                last_tick[k] = int(run[-1])
                tick = run + k_idx
                s = int(sync[k])
                # a synchronisation found later than this chunk does not reach back into it
                cn0 = cn0_column(done[k], len(rec), s if 0 <= s < done[k] + len(rec) else -1)
                database.addTrackingRecords(cid, rec, time=float(now()), time_sample=tick, cn0=cn0,
                                            kaplan=res["kaplan"][k] if "kaplan" in res else None)
                # navigation bits of this chunk -> preamble search / subframe sync -> DECODING_UPDATE rows.  Bit j of a
                # channel ends with epoch sync + 20 (j + 1) - 1 (channel_l1ca_borre.py:455-470), an epoch of this chunk.
                if "bits" in res and s >= 0:
                    fr = frames[k]
                    for b in res["bits"][k]:
                        fr.navBitsBuffer[fr.navBitsCounter] = int(b)
                        fr.navBitsCounter += 1
                        j = bits_done[k]
                        bits_done[k] += 1
                        got = advance_frame(fr)
                        if got is None:
                            continue
                        tow, subframe_id, subframe_bits = got
                        e = min(max(s + 20 * (j + 1) - 1 - done[k], 0), len(rec) - 1)
                        database.addData("decoding", {"cid": cid, "type": ChannelMessage.DECODING_UPDATE, "subframe_id": int(subframe_id),
                                                      "tow": tow, "bits": subframe_bits, "channel_id": cid, "time": float(now()),
                                                      "time_sample": int(tick[e])})
                        decoded_rows += 1
                done[k] += len(rec)
                rows += len(rec)
            database.commit()
        return dict(channels=chans, tracking_rows=rows, decoding_rows=decoded_rows)
