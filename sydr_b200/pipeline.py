"""Cold-start pipeline on one GPU: 32-PRN PCPS acquisition -> hand-off -> batched closed-loop
tracking of the acquired channels over a recording (the BASELINE.json headline workload).

This is the public call a user makes for whole-recording processing; it replaces the
receiver's per-millisecond loop over per-channel processes
(sydr/receiver/receiver.py:120-139, sydr/channel/channelManager.py:149-188) by two batched
GPU dispatches with the reference's scalar hand-off (channel_l1ca_borre.py:301-316) between
them.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib as L
from .engine import (CODE_CHIPS, CODE_FREQ, IQ_PAD_BYTES, AcquisitionEngine, KaplanTrackingEngine, TrackingEngine,
                     make_kaplan_states, make_trk_states, n_complex_samples)


def _mark():
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


class ColdStartPipeline:
    def __init__(self, fs, nbits, search_prns, n_channels, doppler_range=5000.0, doppler_step=250.0, coh=1,
                 noncoh=10, max_seconds=2.0, inter_freq=0.0, threshold=1.5, channel_cfg=None, device=None,
                 cluster=0, threads=0, use_tma=True, loop="borre", dense=False):
        L.require_device()
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.fs, self.nbits, self.inter_freq = float(fs), int(nbits), float(inter_freq)
        self.n_channels, self.threshold = int(n_channels), float(threshold)
        self.channel_cfg = channel_cfg
        self.trk_cfg = dict(cluster=cluster, threads=threads, use_tma=use_tma)
        self._dense = bool(dense)
        self.acq = AcquisitionEngine(fs, inter_freq, doppler_range, doppler_step, coh, noncoh, list(search_prns),
                                     device=self.device)
        self.max_samples = int(round(max_seconds * fs))
        self.max_epochs = int(math.ceil(max_seconds * 1000.0)) + 8
        self._tdt = torch.int8 if nbits == 8 else torch.int16
        pad = IQ_PAD_BYTES // (1 if nbits == 8 else 2)
        self._iq_elems = 2 * self.max_samples + pad
        self._d_iq_buf = None                                     # upload buffer, allocated on first use (device-resident callers never need it)
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._side_stream = torch.cuda.Stream(device=self.device)
        # Hand-off on the device (K-HAND): a template state with the loop coefficients, the tracking
        # engine for n_channels slots (unused ones idle) and a pinned landing zone for the peak table,
        # so that acquisition, hand-off and tracking are enqueued back to back with no host round trip.
        # loop = "borre" (channel_l1ca_borre.py, DLL + Costas PLL) or "kaplan" (channel_l1ca_kaplan.py, FLL-assisted
        # PLL with lock detectors and the PULL_IN / WIDE / NARROW machine): which loop closure the kernel runs
        self.loop = str(loop)
        dummy = [dict(prn=1, carrier_freq=0.0, start_sample=0)]
        if self.loop == "kaplan":
            tmpl, ktmpl = make_kaplan_states(self.fs, dummy, channel_cfg)
        elif self.loop == "borre":
            tmpl, ktmpl = make_trk_states(self.fs, dummy, channel_cfg), None
        else:
            raise L.SydrError(f"unknown loop closure '{loop}'")
        self._tmpl = torch.from_numpy(tmpl.view(np.uint8).reshape(-1).copy()).to(self.device)
        idle = np.repeat(tmpl, self.n_channels)
        idle["status"] = 1
        if ktmpl is None:
            self._trk = TrackingEngine(self.fs, idle, self.max_epochs, device=self.device, dense=self._dense,
                                       **self.trk_cfg)
            self._ktmpl = None
        else:
            ks = np.repeat(ktmpl, self.n_channels)
            self._trk = KaplanTrackingEngine(self.fs, idle, ks, self.max_epochs, device=self.device, **self.trk_cfg)
            self._ktmpl = self._trk._kstates.clone()          # fresh Kaplan states, copied in at every hand-off
        self._n_sel = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._peaks_host = torch.empty(len(self.acq.prns) * 24, dtype=torch.uint8, pin_memory=True)
        self._track_required = int(math.ceil(CODE_CHIPS / (CODE_FREQ / self.fs)))

    def close(self):
        self.acq.close()

    @property
    def _d_iq(self) -> torch.Tensor:
        if self._d_iq_buf is None:
            self._d_iq_buf = torch.zeros(self._iq_elems, dtype=self._tdt, device=self.device)
        return self._d_iq_buf

    # ---- stages --------------------------------------------------------------------------
    def device_buffer(self, n_samples: int) -> torch.Tensor:
        return self._d_iq[:2 * n_samples]

    def upload(self, host_iq: torch.Tensor) -> torch.Tensor:
        """Pinned host interleaved IQ -> device (async on the current stream)."""
        n = host_iq.numel()
        if n > 2 * self.max_samples:
            raise L.SydrError("recording chunk longer than max_seconds")
        d = self._d_iq[:n]
        d.copy_(host_iq, non_blocking=True)
        return d

    def select_channels(self, peaks: np.ndarray):
        """PRNs whose metric exceeds the threshold, best first, at most n_channels."""
        order = np.argsort(-peaks["ratio"], kind="stable")
        sel = [i for i in order if peaks["ratio"][i] > self.threshold][:self.n_channels]
        return sorted(sel, key=lambda i: int(peaks["prn"][i]))

    def _handoff(self, n_samples: int, stream=None):
        """Enqueue K-HAND: peak table -> channel states on the device (channel_l1ca_borre.py:301-311)."""
        a = self.acq
        if self._ktmpl is not None:
            self._trk._kstates.copy_(self._ktmpl, non_blocking=True)
        L.check(L.load().sydr_acq_handoff(a.peaks_device().data_ptr(), len(a.prns), a.inter_freq, a.doppler_range,
                                          a.doppler_step, a.required_samples, self._track_required, 0,
                                          self.threshold, self._tmpl.data_ptr(), int(n_samples),
                                          self._trk._states.data_ptr(), self.n_channels, self._n_sel.data_ptr(),
                                          (stream or torch.cuda.current_stream()).cuda_stream), "sydr_acq_handoff")

    def _peaks_to_host_async(self):
        """Copy the peak table to pinned memory on a side stream, behind the acquisition only."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._side_stream.wait_event(ev)
        with torch.cuda.stream(self._side_stream):
            self._peaks_host.copy_(self.acq.peaks_device(), non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._side_stream)
        return done

    def _channels_of(self, peaks: np.ndarray, n_samples: int):
        """The channel list the device hand-off produced, restated on the host for the caller."""
        chans = []
        for i in self.select_channels(peaks):
            carrier, _, cur = self.acq.handoff(peaks[i])
            chans.append(dict(prn=int(peaks["prn"][i]), carrier_freq=carrier, start_sample=cur, iq_len=n_samples))
        return chans

    def enqueue_device(self, d_iq: torch.Tensor, marks: list | None = None, peaks_out: torch.Tensor | None = None) -> dict:
        """Enqueue acquisition, hand-off and tracking of IQ resident in HBM on the current stream (three
        launches back to back; the peak table travels to pinned memory on a side stream).  Returns the
        context finish() needs; nothing here waits for the GPU.  `marks` (optional) receives four timing
        events: before / after the acquisition, before / after tracking."""
        n = n_complex_samples(d_iq)
        mark = (lambda: marks.append(_mark())) if marks is not None else (lambda: None)
        mark()
        self.acq.launch(d_iq)
        mark()
        got = self._peaks_to_host_async()
        if peaks_out is not None:
            # a device copy of this step's peak table for the caller (multi-GPU: what the all-gather sends), in stream
            # order between acquisition and tracking: on a side stream it would wait for SM room behind the tracking
            # launches in flight
            peaks_out.copy_(self.acq.peaks_device(), non_blocking=True)
        self._handoff(n)
        mark()
        self._trk.launch(d_iq)
        mark()
        return dict(got=got, n=n)

    def finish(self, ctx: dict, records: bool = False, copy: bool = False) -> dict:
        """Host side of an enqueued step: the peak table, the channel list and (records=True) the
        per-epoch tracking records."""
        ctx["got"].synchronize()
        peaks = self._peaks_host.numpy().view(L.ACQ_PEAK_DTYPE).copy()
        chans = self._channels_of(peaks, ctx["n"])
        self._n_active = len(chans)
        out = dict(peaks=peaks, channels=chans)
        if records:
            out["epochs"] = self.collect(copy=copy)
            if self._ktmpl is not None:
                out["kaplan"] = self._trk.fetch_kaplan()[:self._n_active]
        return out

    def process_device(self, d_iq: torch.Tensor, marks: list | None = None) -> dict:
        """Acquisition + hand-off + tracking on IQ already resident in HBM (records stay on the device
        until collect())."""
        return self.finish(self.enqueue_device(d_iq, marks))

    def collect(self, copy: bool = True) -> list:
        """D2H of the per-epoch tracking records of the last process_*().  copy=False: views of pinned
        staging memory, valid until three further collect() / process_host() calls."""
        return self._trk.fetch(copy=copy)[:self._n_active]

    def enqueue_host(self, host_iq: torch.Tensor, pieces: int = 4, peaks_out: torch.Tensor | None = None) -> dict:
        """Enqueue one end-to-end step from pinned host IQ: the upload is cut into `pieces` segments on
        a copy stream; acquisition starts as soon as the dwell has landed and tracking follows the
        upload piece by piece (state carried on the device, records appended), so H2D and compute
        overlap.  Nothing here waits for the GPU."""
        n_el = host_iq.numel()
        n = n_el // 2
        if n > self.max_samples:
            raise L.SydrError("recording chunk longer than max_seconds")
        comp = torch.cuda.current_stream()
        first = min(n, self.acq.required_samples + 4 * self.acq.n_code)
        step = max(1, -(-(n - first) // max(1, pieces)))
        bounds = [first]
        while bounds[-1] < n:
            bounds.append(min(n, bounds[-1] + step))
        d = self._d_iq[:n_el]
        events = []
        self._copy_stream.wait_stream(comp)                      # earlier kernels are done with the buffer
        with torch.cuda.stream(self._copy_stream):
            lo = 0
            for hi in bounds:
                d[2 * lo:2 * hi].copy_(host_iq[2 * lo:2 * hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                events.append(ev)
                lo = hi
        comp.wait_event(events[0])
        self.acq.launch(d[:2 * first])
        got = self._peaks_to_host_async()
        if peaks_out is not None:
            peaks_out.copy_(self.acq.peaks_device(), non_blocking=True)
        self._handoff(n)
        for hi, ev in zip(bounds, events):
            comp.wait_event(ev)
            self._trk.launch(d, iq_len=hi, append=True)
        return dict(got=got, n=n)

    def process_host(self, host_iq: torch.Tensor, pieces: int = 4, copy: bool = False) -> dict:
        """End to end: pinned host IQ in, acquisition table + per-epoch tracking records out.  The
        record arrays are views of pinned staging memory that stay valid for the next three calls
        (copy=True for private copies: first-touch page faults of 3 MB cost ~1 ms per call)."""
        return self.finish(self.enqueue_host(host_iq, pieces), records=True, copy=copy)


class ColdStartPool:
    """Several steps in flight on one GPU.

    The 12-channel tracking kernel is a latency chain that occupies 96 of the 148 SMs; the
    acquisition of the *next* chunk (or recording) fits beside it.  The pool owns `lanes`
    independent ColdStartPipeline instances, each on its own stream; submit_*() enqueues a whole
    step on the next lane and returns a ticket at once, result() hands back what process_*() would.
    With two lanes the step period drops from acquisition + tracking to about the SM-time bound
    (5.7 -> 4.4 ms for the headline chunk).  This is the shape for recordings that ARRIVE one by one (submit_host: the
    upload of one overlaps the tracking of another; PCIe-bound).  Recordings already resident in HBM are tracked
    2.1 x faster side by side in one launch (ColdStartBatch); with more than 2 lanes x 3 streams set
    CUDA_DEVICE_MAX_CONNECTIONS=32 before CUDA starts, or streams share hardware queues and serialise
    (profiles/r2/pack_shapes.txt)."""

    def __init__(self, lanes: int = 2, **pipeline_kwargs):
        pipeline_kwargs.setdefault("dense", int(lanes) > 1)       # launches share the GPU: pack the tracking CTAs 3 per SM
        self.lanes = [ColdStartPipeline(**pipeline_kwargs) for _ in range(int(lanes))]
        dev = self.lanes[0].device
        self._streams = [torch.cuda.Stream(device=dev) for _ in self.lanes]
        self._pending = {}
        self._next = 0
        self._ticket = 0

    def close(self):
        for p in self.lanes:
            p.close()

    def _lane(self):
        i = self._next
        if any(l == i for l, _ in self._pending.values()):
            raise L.SydrError("every lane has a step in flight: collect a result() first")
        self._next = (i + 1) % len(self.lanes)
        return i

    def submit_device(self, d_iq: torch.Tensor, marks: list | None = None, peaks_out: torch.Tensor | None = None) -> int:
        i = self._lane()
        s = self._streams[i]
        s.wait_stream(torch.cuda.current_stream())               # the caller's work on d_iq is ordered first
        with torch.cuda.stream(s):
            ctx = self.lanes[i].enqueue_device(d_iq, marks, peaks_out)
        self._ticket += 1
        self._pending[self._ticket] = (i, ctx)
        return self._ticket

    def submit_host(self, host_iq: torch.Tensor, pieces: int = 4, peaks_out: torch.Tensor | None = None) -> int:
        i = self._lane()
        with torch.cuda.stream(self._streams[i]):
            ctx = self.lanes[i].enqueue_host(host_iq, pieces, peaks_out)
        self._ticket += 1
        self._pending[self._ticket] = (i, ctx)
        return self._ticket

    def result(self, ticket: int, records: bool = True, copy: bool = False) -> dict:
        i, ctx = self._pending.pop(ticket)
        with torch.cuda.stream(self._streams[i]):
            out = self.lanes[i].finish(ctx, records=records, copy=copy)
        out["lane"] = i
        return out

    def lane_index(self, ticket: int) -> int:
        return self._pending[ticket][0]

    def stream(self, lane: int) -> torch.cuda.Stream:
        return self._streams[lane]

    def peak_stream(self, lane: int) -> torch.cuda.Stream:
        """The lane's side stream: ordered behind its acquisition (peak table complete), not behind its
        tracking.  Where a multi-GPU caller enqueues the all-gather of the 768-byte peak table."""
        return self.lanes[lane]._side_stream


class ColdStartBatch:
    """B recordings per step, tracked by ONE launch (throughput shape; BASELINE.json configs[4] with the cold start of
    configs[2] in front): the recordings lie back to back in one device buffer (`stride` int8/int16 elements apart, see
    slot()); a step enqueues, on one stream, acquisition + device hand-off (K-HAND) of every recording -- each hand-off
    writes the recording's n_channels states with iq_base at its slot -- and then a single tracking launch over all
    B x n_channels channels with the PACK instantiation of K-TRK (cfg.dense = 2: one CTA and one staged window per channel,
    two channels per SM, so that one channel's serial loop closure runs under the other's correlation).  148 SMs hold
    296 channels: 24 recordings of 12 channels per launch.  Per-recording results equal ColdStartPipeline's to the
    summation-order tolerance of the correlator sums.  Replaces the per-channel processes of
    sydr/receiver/receiver.py:120-139 / sydr/channel/channelManager.py:149-188 for many recordings at once."""

    def __init__(self, recordings, fs, nbits, search_prns, n_channels, doppler_range=5000.0, doppler_step=250.0, coh=1,
                 noncoh=10, max_seconds=2.0, inter_freq=0.0, threshold=1.5, channel_cfg=None, device=None,
                 cluster=1, threads=0, use_tma=True, dense=2):
        L.require_device()
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.B, self.n_channels = int(recordings), int(n_channels)
        self.fs, self.nbits, self.threshold = float(fs), int(nbits), float(threshold)
        self.acq = AcquisitionEngine(fs, inter_freq, doppler_range, doppler_step, coh, noncoh, list(search_prns), device=self.device)
        self.max_samples = int(round(max_seconds * fs))
        self.max_epochs = int(math.ceil(max_seconds * 1000.0)) + 8
        self._tdt = torch.int8 if nbits == 8 else torch.int16
        pad = IQ_PAD_BYTES // (1 if nbits == 8 else 2)
        self.stride = (2 * self.max_samples + pad + 15) // 16 * 16       # elements between recordings (iq_base: a multiple of 8 samples)
        tmpl = make_trk_states(self.fs, [dict(prn=1, carrier_freq=0.0, start_sample=0)], channel_cfg)
        tmpls = np.repeat(tmpl, self.B)
        tmpls["iq_base"] = np.arange(self.B, dtype=np.int64) * (self.stride // 2)
        self._tmpls = torch.from_numpy(tmpls.view(np.uint8).reshape(self.B, -1).copy()).to(self.device)
        idle = np.repeat(tmpls, self.n_channels)
        idle["status"] = 1
        self._trk = TrackingEngine(self.fs, idle, self.max_epochs, device=self.device, cluster=cluster, threads=threads,
                                   use_tma=use_tma, dense=dense)
        self._state_bytes = tmpl.dtype.itemsize
        self._n_sel = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        self._peak_bytes = len(self.acq.prns) * 24
        self._peaks_dev = torch.zeros(self.B * self._peak_bytes, dtype=torch.uint8, device=self.device)
        self._peaks_host = torch.empty(self.B * self._peak_bytes, dtype=torch.uint8, pin_memory=True)
        self._track_required = int(math.ceil(CODE_CHIPS / (CODE_FREQ / self.fs)))
        self._side_stream = torch.cuda.Stream(device=self.device)
        self._buf = None

    def close(self):
        self.acq.close()

    def buffer(self) -> torch.Tensor:
        """The device buffer of the B recordings (allocated on first use)."""
        if self._buf is None:
            # (the storage extends IQ_PAD_BYTES past the last recording: the tracking launch takes the view as it is)
            pad = IQ_PAD_BYTES // (1 if self.nbits == 8 else 2)
            self._store = torch.zeros(self.B * self.stride + pad, dtype=self._tdt, device=self.device)
            self._buf = self._store[:self.B * self.stride]
        return self._buf

    def release(self):
        """Give the device buffer of the recordings back (it is allocated again on the next buffer() / slot())."""
        self._buf = self._store = None

    def slot(self, r: int, n_samples: int | None = None) -> torch.Tensor:
        """Where recording r lives: interleaved I/Q elements [r * stride, r * stride + 2 n)."""
        n = self.max_samples if n_samples is None else int(n_samples)
        if n > self.max_samples:
            raise L.SydrError("recording longer than max_seconds")
        return self.buffer()[r * self.stride:r * self.stride + 2 * n]

    def enqueue(self, n_samples: int | None = None, marks: list | None = None, peaks_out: torch.Tensor | None = None) -> dict:
        """Enqueue one step over the recordings resident in buffer() on the current stream: B x (acquisition, hand-off),
        then one tracking launch.  `marks` receives four timing events: before / after the acquisitions + hand-offs,
        before / after the tracking launch.  Nothing here waits for the GPU."""
        n = self.max_samples if n_samples is None else int(n_samples)
        a, lib = self.acq, L.load()
        cur = torch.cuda.current_stream()
        mark = (lambda: marks.append(_mark())) if marks is not None else (lambda: None)
        buf = self.buffer()
        mark()
        for r in range(self.B):
            a.launch(buf[r * self.stride:r * self.stride + 2 * a.required_samples])
            self._peaks_dev[r * self._peak_bytes:(r + 1) * self._peak_bytes].copy_(a.peaks_device(), non_blocking=True)
            L.check(lib.sydr_acq_handoff(a.peaks_device().data_ptr(), len(a.prns), a.inter_freq, a.doppler_range,
                                         a.doppler_step, a.required_samples, self._track_required, 0, self.threshold,
                                         self._tmpls[r].data_ptr(), n,
                                         self._trk._states.data_ptr() + r * self.n_channels * self._state_bytes,
                                         self.n_channels, self._n_sel[r:].data_ptr(), cur.cuda_stream), "sydr_acq_handoff")
        if peaks_out is not None:          # a device copy of the step's B peak tables for the caller (multi-GPU: what the all-gather sends)
            peaks_out.copy_(self._peaks_dev, non_blocking=True)
        mark()
        ev = torch.cuda.Event()
        ev.record(cur)
        self._side_stream.wait_event(ev)
        with torch.cuda.stream(self._side_stream):                 # the peak tables travel behind the acquisitions only
            self._peaks_host.copy_(self._peaks_dev, non_blocking=True)
            got = torch.cuda.Event()
            got.record(self._side_stream)
        mark()
        self._trk.launch(buf)
        mark()
        return dict(got=got, n=n)

    def finish(self, ctx: dict, records: bool = False, copy: bool = False) -> list:
        """Host side of a step: per recording the peak table, the channel list and (records=True) the per-epoch
        tracking records of its channels."""
        ctx["got"].synchronize()
        peaks = self._peaks_host.numpy().view(L.ACQ_PEAK_DTYPE).reshape(self.B, -1).copy()
        recs = self._trk.fetch(copy=copy) if records else None
        out = []
        for r in range(self.B):
            order = np.argsort(-peaks[r]["ratio"], kind="stable")
            sel = sorted([i for i in order if peaks[r]["ratio"][i] > self.threshold][:self.n_channels], key=lambda i: int(peaks[r]["prn"][i]))
            chans = []
            for i in sel:
                carrier, _, start = self.acq.handoff(peaks[r][i])
                chans.append(dict(prn=int(peaks[r]["prn"][i]), carrier_freq=carrier, start_sample=start, iq_len=ctx["n"], rec=r))
            o = dict(peaks=peaks[r], channels=chans)
            if records:
                o["epochs"] = recs[r * self.n_channels:r * self.n_channels + len(chans)]
            out.append(o)
        return out

    def device_summary(self, last: int = 200):
        """Without moving the records to the host: per channel slot the epochs tracked, the status and the mean carrier
        frequency over its last `last` epochs (what a lock / truth check needs), reduced on the device."""
        nep = self._trk._nep.to(torch.int64)
        rows = self._trk._out.view(torch.float64).view(self._trk.n_ch, self.max_epochs, 16)[:, :, 8]      # carrier_freq
        k = torch.arange(self.max_epochs, device=self.device)[None, :]
        m = (k < nep[:, None]) & (k >= (nep[:, None] - last))
        mean = torch.where(m, rows, torch.zeros((), dtype=rows.dtype, device=self.device)).sum(dim=1) / m.sum(dim=1).clamp(min=1)
        st = self._trk.states()
        return nep.cpu().numpy(), st["status"].copy(), st["prn"].copy(), mean.cpu().numpy()

    def process(self, n_samples: int | None = None, records: bool = True, copy: bool = True) -> list:
        return self.finish(self.enqueue(n_samples), records=records, copy=copy)
