"""Batched GPU dispatcher with the interface of sydr/channel/channelManager.py (ChannelManager).

The reference runs one OS process per channel and lock-steps them every millisecond with two
Events and one pickled Queue message each (channelManager.py:149-188, channel.py:121-160).
Here all channels live in this process and share one copy of the samples in HBM:

  * `addNewRFData` uploads the new samples to a linear device buffer -- the interleaved
    int8/int16 integers of the file when the data comes from `sydr_b200.signal.rfsignal.RFSignal`
    (2 or 4 B/sample instead of the reference's 16 B complex128 ring).  When the block is a slice
    of the reader's current 120 ms file chunk, the rest of that chunk is uploaded as well
    (look-ahead), so the GPU can work on every epoch the chunk contains in one launch;
  * `run` is one millisecond tick.  Acquisition of every channel whose dwell is complete is one
    batched `sydr_acq_run`; tracking of all channels is one `sydr_trk_run` that advances each
    channel over all samples present on the device and appends its epoch records; the records
    are then replayed on the host against the reference's per-tick rule ("an epoch completes at
    the first tick where the unread samples reach track_requiredSamples, at most one per tick",
    channel_l1ca_borre.py:347-349) so that the packets, their tick, `unprocessed_samples` and
    `time_since_tow` are what the reference's channel processes would have reported.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import _lib as L
from ..engine import (AcquisitionEngine, KaplanTrackingEngine, TrackingEngine, make_kaplan_states, make_trk_states)
from ..signal.rfsignal import IQBlock, RFSignal
from ..utils.circularbuffer import CircularBuffer
from ..utils.enumerations import ChannelState
from .channel import Channel


class ChannelManager:
    TIMEOUT = 100
    DEVICE_BUFFER_MS = 8000        # capacity of the linear device buffer, in milliseconds of signal
    MAX_QUEUED_EPOCHS = 512        # epoch records one tracking launch may produce per channel

    def __init__(self, rfSignal: RFSignal, keepCorrelationMaps: bool = True, hostCopy: bool = False, device=None):
        L.require_device()
        self.rfSignal = rfSignal
        self.channels = {}
        self.nbChannels = 0
        self.keepCorrelationMaps = keepCorrelationMaps
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        # index arithmetic of the reference's 100 ms ring (channelManager.py:57-61); samples stay on the GPU
        buffersize = int(self.rfSignal.samplingFrequency * 1e-3 * 100)
        self.sharedBuffer = CircularBuffer(buffersize, self.rfSignal.dtype, store=hostCopy)
        self.resultQueue = None
        self._d_iq = None            # device samples, linear; index 0 = absolute sample self._base
        self._iq_dtype = None
        self._capacity = 0           # complex samples
        self._base = 0
        self._written = 0            # absolute samples handed over through addNewRFData
        self._uploaded = 0           # absolute samples present on the device (>= written with look-ahead)
        self._acq_engines = {}
        self._trk = None             # TrackingEngine over the slots of all channels
        self._trk_slots = {}         # channel id -> slot
        self._queues = {}            # channel id -> list of pending epoch records
        self._abs_cur = {}           # channel id -> absolute sample of the next epoch start
        self._streams = None

    # ---- channel set ---------------------------------------------------------------------------
    def addChannel(self, ChannelObject, configuration: dict, nbChannels=1):
        for _ in range(nbChannels):
            cid = self.nbChannels
            self.channels[cid] = ChannelObject(cid, self.sharedBuffer, self.resultQueue, self.rfSignal, configuration)
            self.nbChannels += 1

    def requestTracking(self, satelliteID: int):
        for channel in self.channels.values():
            if channel.channelState is ChannelState.IDLE:
                channel.setSatellite(satelliteID)
                channel.start()
                start = self._written + channel.currentSample - self.sharedBuffer.idxWrite \
                    if self.sharedBuffer.full else channel.currentSample
                if start < self._base:
                    # the device buffer has been compacted past the ring position the channel wants to start from (the
                    # reference's ring keeps 100 ms; _compact keeps that much behind the write position, so this means a
                    # start further back than the ring itself holds)
                    raise L.SydrError(f"channel {channel.channelID}: start sample {start} lies before the oldest sample "
                                      f"kept on the device ({self._base})")
                self._abs_cur[channel.channelID] = start
                logging.getLogger(__name__).debug(f"CID {channel.channelID} initialised to satellite [G{satelliteID}].")
                return channel
        raise Warning(f"Could not find an IDLE channel for tracking satellite [G{satelliteID}].")

    def getChannel(self, channelID):
        if channelID not in self.channels:
            raise ValueError("Channel ID does not exist.")
        return self.channels[channelID]

    def close(self):
        for eng in self._acq_engines.values():
            eng.close()
        self._acq_engines.clear()
        self._trk = None
        self._d_iq = None

    # ---- samples -------------------------------------------------------------------------------
    def _ensure_buffer(self, sample):
        """Allocate the device buffer for the first block's representation."""
        if self._d_iq is not None:
            return
        if sample.dtype in (np.int8, np.int16) and not getattr(self.rfSignal, "isComplex", True):
            # real-valued integer samples are not interleaved I,Q pairs (rfsignal.py:107-126): reading them as pairs would
            # silently halve the sample count
            raise L.SydrError("ChannelManager: real-valued integer samples are not supported (is_complex = false); "
                              "pass complex samples")
        if sample.dtype == np.int8:
            self._iq_dtype, tdt, per = L.IQ_I8, torch.int8, 2
        elif sample.dtype == np.int16:
            self._iq_dtype, tdt, per = L.IQ_I16, torch.int16, 2
        else:
            self._iq_dtype, tdt, per = L.IQ_F32, torch.complex64, 1
        self._per = per
        self._capacity = int(self.rfSignal.samplingFrequency * 1e-3 * self.DEVICE_BUFFER_MS)
        pad = 4096
        self._d_iq = torch.zeros(self._capacity * per + pad, dtype=tdt, device=self.device)

    def _upload(self, block, abs_start):
        """Copy host samples (raw interleaved integers or complex) to the device at `abs_start`."""
        n = len(block) // self._per
        if abs_start + n - self._base > self._capacity:
            self._compact()
        lo = (abs_start - self._base) * self._per
        src = torch.from_numpy(np.ascontiguousarray(block))
        self._d_iq[lo:lo + len(block)].copy_(src, non_blocking=False)
        self._uploaded = max(self._uploaded, abs_start + n)

    def _compact(self):
        """Drop the samples every channel has consumed: move the tail of the device buffer to its
        start and rebase the device-side channel states."""
        # what a channel started later may still ask for: one ring length behind the write position (requestTracking)
        ring = int(getattr(self.sharedBuffer, "maxSize", 0))
        keep_from = min([max(self._written - ring, self._base)] + [c for cid, c in self._abs_cur.items()
                                                                    if self.channels[cid].channelState is not ChannelState.IDLE])
        keep_from -= keep_from % 16
        shift = keep_from - self._base
        if shift <= 0:
            raise L.SydrError("device sample buffer exhausted: raise ChannelManager.DEVICE_BUFFER_MS")
        n_keep = (self._uploaded - keep_from) * self._per
        tail = self._d_iq[shift * self._per:shift * self._per + n_keep].clone()
        self._d_iq[:n_keep].copy_(tail)
        self._base = keep_from
        if self._trk is not None:
            st = self._trk.states()
            st["cur"] -= shift
            self._trk.reset(st)

    def addNewRFData(self, data):
        """channelManager.py:131-145.  `data`: the next block of samples (normally 1 ms)."""
        raw = getattr(data, "raw", None)
        block = raw if raw is not None else np.asarray(data)
        if raw is None and np.iscomplexobj(block):
            block = block.astype(np.complex64)
        self._ensure_buffer(block)
        n = len(block) // self._per
        if self._written + n > self._uploaded:
            ahead = self._lookahead(raw) if raw is not None else None
            if ahead is not None:
                self._upload(ahead, self._written)          # the rest of the reader's chunk, this block first
            else:
                self._upload(block, self._written)
        if self.sharedBuffer.buffer is not None:
            self.sharedBuffer.shift(np.asarray(data))
        else:
            if self.sharedBuffer.maxSize % n != 0:
                raise ValueError("Data shift need to be a multiple from the max buffer size.")
            self.sharedBuffer.shiftIdxWrite(n)
        self._written += n

    def _lookahead(self, raw):
        """If `raw` is a slice of the reader's current file chunk, return that chunk from the
        slice's start to its end (the samples of the coming ticks)."""
        chunk = getattr(self.rfSignal, "chunck", None)
        craw = getattr(chunk, "raw", None)
        if craw is None or raw.dtype != craw.dtype:
            return None
        a = raw.__array_interface__['data'][0]
        c0 = craw.__array_interface__['data'][0]
        if not (c0 <= a < c0 + craw.nbytes):
            return None
        off = (a - c0) // craw.itemsize
        return craw[off:]

    def prefetch(self, data):
        """Upload samples of coming ticks ahead of time (they become visible to the channels only
        when `addNewRFData` hands them over); lets the GPU work on many epochs per launch."""
        raw = getattr(data, "raw", None)
        block = raw if raw is not None else np.asarray(data)
        if raw is None and np.iscomplexobj(block):
            block = block.astype(np.complex64)
        self._ensure_buffer(block)
        self._upload(block, max(self._uploaded, self._written))

    # ---- one tick ------------------------------------------------------------------------------
    def run(self):
        """channelManager.py:149-188: process the current buffer content, return the flattened
        packets of all channels for this millisecond."""
        active = [c for c in self.channels.values() if c.is_alive()]
        # Channel classes without a device loop closure in this dispatcher (the Kaplan variant) run their own
        # _processHandler() per tick, like the reference's channel process: GPU correlators / acquisition through
        # the drop-in functions, loops on the host.  They read the host copy of the ring.
        alone = [c for c in active if not hasattr(c, "_trackingConfiguration")]
        if alone and self.sharedBuffer.buffer is None:
            raise L.SydrError("channels ticked stand-alone need ChannelManager(hostCopy=True)")
        batched = [c for c in active if hasattr(c, "_trackingConfiguration")]
        self._acquire([c for c in batched if c.channelState is ChannelState.ACQUIRING])
        tracking = [c for c in batched if c.channelState is ChannelState.TRACKING]
        self._track_ahead(tracking)
        results = []
        for chan in active:
            if chan in alone:
                results.extend(chan._processHandler())
                results.append(chan.prepareChannelUpdate())
                continue
            packets = []
            acq = getattr(chan, "_pendingAcquisition", None)
            if acq is not None:
                packets.append(acq)
                chan._pendingAcquisition = None
            elif chan.channelState is ChannelState.TRACKING:
                rec = self._due_epoch(chan)
                if rec is not None:
                    packets.append(self._ingest(chan, rec))
                dec = chan.runDecoding()
                if dec is not None:
                    packets.append(dec)
            chan._afterTick()
            packets.append(chan.prepareChannelUpdate())
            results.extend(packets)
        return results

    def runBlock(self, data, nbMilliseconds: int):
        """Convenience for whole-block processing: upload `data` (nbMilliseconds of samples) once,
        then tick through it; returns one packet list per millisecond."""
        self.prefetch(data)
        out = []
        spm = self.rfSignal.samplesPerMs
        for k in range(nbMilliseconds):
            if isinstance(data, IQBlock):
                blk = data.block(k * spm, (k + 1) * spm)
            else:
                per = 1 if np.iscomplexobj(data) else 2          # complex samples or interleaved I,Q integers
                blk = data[k * spm * per:(k + 1) * spm * per]
            self.addNewRFData(blk)
            out.append(self.run())
        return out

    # ---- acquisition ---------------------------------------------------------------------------
    def _acquire(self, chans):
        ready = [c for c in chans if self.sharedBuffer.getNbUnreadSamples(c.currentSample) >= c.acq_requiredSamples]
        groups = {}
        for c in ready:
            key = (c.acq_dopplerRange, c.acq_dopplerSteps, c.acq_coherentIntegration, c.acq_nonCoherentIntegration,
                   self._abs_cur[c.channelID])
            groups.setdefault(key, []).append(c)
        fs = self.rfSignal.samplingFrequency
        for (dr, ds, coh, noncoh, start), group in groups.items():
            prns = tuple(int(c.satelliteID) for c in group)
            ekey = (dr, ds, coh, noncoh, prns)
            eng = self._acq_engines.get(ekey)
            if eng is None:
                eng = AcquisitionEngine(fs, self.rfSignal.interFrequency, dr, ds, coh, noncoh, list(prns),
                                        device=self.device)
                self._acq_engines[ekey] = eng
            lo = (start - self._base) * self._per
            view = self._d_iq[lo:lo + eng.required_samples * self._per]
            res = eng.run(view, want_maps=self.keepCorrelationMaps)
            for slot, c in enumerate(group):
                pk = res["peaks"][slot]
                cmap = res["maps"][slot].astype(np.float64) if self.keepCorrelationMaps else None
                before = c.currentSample
                c._pendingAcquisition = c._ingestAcquisition([int(pk["freq_idx"]), int(pk["code_idx"])],
                                                              float(pk["ratio"]), cmap)
                self._abs_cur[c.channelID] += c.currentSample - before
                self._start_tracking_slot(c)

    # ---- tracking ------------------------------------------------------------------------------
    def _start_tracking_slot(self, chan):
        """Give the channel a slot in the device-side state array (all slots exist from the first
        use; idle ones carry status 1 and are skipped by the kernel)."""
        fs = self.rfSignal.samplingFrequency
        kaplan = getattr(chan, "_deviceLoop", "borre") == "kaplan"
        one = [dict(prn=int(chan.satelliteID), carrier_freq=chan.carrierFrequency,
                    start_sample=self._abs_cur[chan.channelID] - self._base)]
        if self._trk is None:
            dummy = [dict(prn=1, carrier_freq=0.0, start_sample=0)] * self.nbChannels
            if kaplan:           # every channel of a manager runs the same loop closure (one class per receiver)
                idle, kidle = make_kaplan_states(fs, dummy, chan._trackingConfiguration)
                idle["status"] = 1
                self._trk = KaplanTrackingEngine(fs, idle, kidle, self.MAX_QUEUED_EPOCHS, device=self.device)
            else:
                idle = make_trk_states(fs, dummy)
                idle["status"] = 1
                self._trk = TrackingEngine(fs, idle, self.MAX_QUEUED_EPOCHS, device=self.device)
            self._trk_slots = {cid: k for k, cid in enumerate(self.channels)}
        if kaplan != isinstance(self._trk, KaplanTrackingEngine):
            raise L.SydrError("a ChannelManager batches one loop closure: do not mix Borre and Kaplan channels")
        st = self._trk.states()
        slot = self._trk_slots[chan.channelID]
        if kaplan:
            new, knew = make_kaplan_states(fs, one, chan._trackingConfiguration)
            ks = self._trk.kaplan_states()
            ks[slot] = knew[0]
            self._trk._kstates.copy_(torch.from_numpy(ks.view(np.uint8).reshape(-1)))
        else:
            new = make_trk_states(fs, one, chan._trackingConfiguration)
        new["epochs_done"] = 0
        st[slot] = new[0]
        self._trk.reset(st)
        self._queues[chan.channelID] = []

    def _track_ahead(self, chans):
        """Launch the tracking kernel when some channel has no pending record although the device
        holds enough samples for its next epoch."""
        need = False
        for c in chans:
            if not self._queues[c.channelID] and \
                    self._uploaded - self._abs_cur[c.channelID] >= c.track_requiredSamples:
                need = True
        if not need:
            return
        # every record of the previous launch has been queued: this launch writes its records from 0
        self._trk.launch(self._d_iq, iq_len=self._uploaded - self._base, append=False)
        recs = self._trk.fetch()
        st = self._trk.states()
        bad = [cid for cid, k in self._trk_slots.items() if st["status"][k] < 0]
        if bad:
            raise L.SydrError(f"tracking aborted on channels {bad} (NCO state left the supported range)")
        kex = self._trk.fetch_kaplan() if isinstance(self._trk, KaplanTrackingEngine) else None
        for c in chans:
            k = self._trk_slots[c.channelID]
            self._queues[c.channelID].extend(recs[k] if kex is None else list(zip(recs[k], kex[k])))

    def _due_epoch(self, chan):
        """The reference's per-tick rule (channel_l1ca_borre.py:347-349)."""
        q = self._queues[chan.channelID]
        if not q:
            return None
        if self.sharedBuffer.getNbUnreadSamples(chan.currentSample) < chan.track_requiredSamples:
            return None
        return q.pop(0)

    def _ingest(self, chan, rec):
        if isinstance(rec, tuple):                      # Kaplan: (record, extras)
            rec, kex = rec
            n = int(rec["n"])
            if n != chan.track_requiredSamples:
                raise L.SydrError(f"CID {chan.channelID}: device epoch length {n} != host {chan.track_requiredSamples}")
            self._abs_cur[chan.channelID] += n
            return chan._ingestEpoch(rec, kex)
        n = int(rec["n"])
        if n != chan.track_requiredSamples:
            raise L.SydrError(f"CID {chan.channelID}: device epoch length {n} != host {chan.track_requiredSamples}")
        self._abs_cur[chan.channelID] += n
        corr = [float(v) for v in rec["corr"]]
        return chan._ingestEpoch(corr, float(rec["dll"]), float(rec["pll"]), float(rec["carrier_freq"]),
                                 float(rec["code_freq"]), float(rec["code_err"]), float(rec["carrier_err"]),
                                 float(rec["rem_code"]), float(rec["rem_carrier"]))
