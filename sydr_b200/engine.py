"""Batched GPU dispatch of the two SyDR hot paths (replaces the per-channel multiprocessing
loop of sydr/channel/channelManager.py:149-188 and sydr/receiver/receiver.py:120-139).

PyTorch is used only for device buffers and streams; all arithmetic happens in
libsydr_b200.so (hand-written sm_100a CUDA) through the C ABI of include/sydr_b200.h.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

CODE_FREQ = 1.023e6
CODE_CHIPS = 1023
IQ_PAD_BYTES = 4096          # readable slack after the last sample (vector loads / TMA windows)

_NP2IQ = {np.dtype(np.int8): L.IQ_I8, np.dtype(np.int16): L.IQ_I16,
          np.dtype(np.complex64): L.IQ_F32, np.dtype(np.complex128): L.IQ_F64}
_TORCH2IQ = {torch.int8: L.IQ_I8, torch.int16: L.IQ_I16, torch.complex64: L.IQ_F32,
             torch.complex128: L.IQ_F64}
_IQ_BYTES = {L.IQ_I8: 2, L.IQ_I16: 4, L.IQ_F32: 8, L.IQ_F64: 16}


def _stream_ptr(stream) -> int:
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    return stream.cuda_stream


def iq_code(t: torch.Tensor) -> int:
    try:
        return _TORCH2IQ[t.dtype]
    except KeyError:
        raise L.SydrError(f"unsupported IQ tensor dtype {t.dtype}") from None


def n_complex_samples(t: torch.Tensor) -> int:
    return t.numel() // 2 if t.dtype in (torch.int8, torch.int16) else t.numel()


def to_device_iq(iq: np.ndarray, device="cuda", pad=True, pinned_src: torch.Tensor | None = None) -> torch.Tensor:
    """Host IQ (interleaved int8/int16, or complex64/128) -> device tensor with readable padding.
    The returned tensor is a view of the valid samples; its storage is IQ_PAD_BYTES longer."""
    L.require_device()
    a = np.ascontiguousarray(iq)
    if a.dtype not in _NP2IQ:
        raise L.SydrError(f"unsupported IQ dtype {a.dtype}")
    src = torch.from_numpy(a.reshape(-1)) if pinned_src is None else pinned_src
    n = src.numel()
    pad_el = (IQ_PAD_BYTES // src.element_size()) if pad else 0
    buf = torch.empty(n + pad_el, dtype=src.dtype, device=device)
    buf[:n].copy_(src, non_blocking=True)
    if pad_el:
        buf[n:].zero_()
    return buf[:n]


def loop_coefficients(bw: float, damping: float, gain: float):
    """LoopFiltersCoefficients, sydr/dsp/tracking.py:56-61 (host scalar, FP64)."""
    wn = bw * 8.0 * damping / (4.0 * damping ** 2 + 1)
    return gain / wn ** 2, 2.0 * damping / wn


# ======================================================================================
class AcquisitionEngine:
    """PCPS + TwoCorrelationPeakComparison for a list of PRNs in one batched dispatch.

    Mirrors the arguments of PCPS (sydr/dsp/acquisition.py:9-10) and the derived sizes of
    ChannelL1CA.runAcquisition (sydr/channel/channel_l1ca_borre.py:281-299).
    """

    def __init__(self, fs, inter_freq, doppler_range, doppler_step, coh, noncoh, prns,
                 bin_lo=0, bin_hi=-1, device=None):
        L.require_device()
        lib = L.load()
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        L.check(lib.sydr_set_device(self.device.index), "sydr_set_device")
        self.prns = np.ascontiguousarray(np.asarray(prns, dtype=np.int32))
        self._plan = C.c_void_p()
        L.check(lib.sydr_acq_plan_create(float(fs), float(inter_freq), float(doppler_range), float(doppler_step),
                                         int(coh), int(noncoh), self.prns.ctypes.data, len(self.prns),
                                         int(bin_lo), int(bin_hi), C.byref(self._plan)), "sydr_acq_plan_create")
        n_code, n_bins, n_rows, chip = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        need = C.c_longlong()
        L.check(lib.sydr_acq_plan_info(self._plan, C.byref(n_code), C.byref(n_bins), C.byref(n_rows), C.byref(chip),
                                       C.byref(need)), "sydr_acq_plan_info")
        self.fs, self.inter_freq = float(fs), float(inter_freq)
        self.doppler_range, self.doppler_step = float(doppler_range), float(doppler_step)
        self.coh, self.noncoh = int(coh), int(noncoh)
        self.n_code, self.n_bins, self.n_rows, self.chip = n_code.value, n_bins.value, n_rows.value, chip.value
        self.bin_lo = int(bin_lo)
        self.required_samples = need.value
        n_prn = len(self.prns)
        self._peaks = torch.empty(n_prn * 24, dtype=torch.uint8, device=self.device)
        self._rows = torch.empty(n_prn * self.n_rows * 16, dtype=torch.uint8, device=self.device)
        self._maps = None

    def close(self):
        if getattr(self, "_plan", None) and self._plan.value:
            L.load().sydr_acq_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_spectrum(self, slot: int, code_fft: np.ndarray):
        """Use the caller's `codeFFT` (PCPS argument) for PRN slot `slot`."""
        spec = np.ascontiguousarray(code_fft, dtype=np.complex128)
        if spec.shape != (self.n_code,):
            raise L.SydrError(f"codeFFT must have {self.n_code} entries")
        L.check(L.load().sydr_acq_plan_set_spectrum(self._plan, int(slot), spec.ctypes.data), "set_spectrum")

    def launch(self, iq_dev: torch.Tensor, want_maps=False, stream=None):
        """Enqueue one acquisition on `stream`; results stay on the device."""
        n = n_complex_samples(iq_dev)
        maps_ptr = 0
        if want_maps:
            need = len(self.prns) * self.n_rows * self.n_code
            if self._maps is None or self._maps.numel() != need:
                self._maps = torch.empty(need, dtype=torch.float32, device=self.device)
            maps_ptr = self._maps.data_ptr()
        L.check(L.load().sydr_acq_run(self._plan, iq_dev.data_ptr(), iq_code(iq_dev), n, self._peaks.data_ptr(),
                                      self._rows.data_ptr(), maps_ptr, _stream_ptr(stream)), "sydr_acq_run")

    def peaks_device(self) -> torch.Tensor:
        return self._peaks

    def fetch(self, want_maps=False, want_rows=False) -> dict:
        out = {"peaks": self._peaks.cpu().numpy().view(L.ACQ_PEAK_DTYPE).copy()}
        if want_rows:
            out["rows"] = self._rows.cpu().numpy().view(L.ACQ_ROW_DTYPE).reshape(len(self.prns), self.n_rows).copy()
        if want_maps and self._maps is not None:
            out["maps"] = self._maps.cpu().numpy().reshape(len(self.prns), self.n_rows, self.n_code)
        return out

    def run(self, iq_dev: torch.Tensor, want_maps=False, want_rows=False, stream=None) -> dict:
        self.launch(iq_dev, want_maps=want_maps, stream=stream)
        return self.fetch(want_maps=want_maps, want_rows=want_rows)

    def handoff(self, peak, current_sample=0, track_required=None):
        """Acquisition -> tracking scalars, channel_l1ca_borre.py:301-311."""
        if track_required is None:
            track_required = int(math.ceil(CODE_CHIPS / (CODE_FREQ / self.fs)))
        doppler = -((-self.doppler_range) + self.doppler_step * int(peak["freq_idx"]))
        code_offset = int(peak["code_idx"])
        carrier = self.inter_freq + doppler
        cur = current_sample + self.required_samples - track_required + code_offset + 1
        return carrier, code_offset, cur


# ======================================================================================
BORRE_DEFAULTS = dict(  # config/channels/channel_GPS_L1CA_borre.ini
    correlator_early=-0.5, correlator_prompt=0.0, correlator_late=0.5,
    dll_damping_ratio=0.7, dll_noise_bandwidth=1.0, dll_loop_gain=1.0, dll_pdi=0.001,
    pll_damping_ratio=0.7, pll_noise_bandwidth=8.0, pll_loop_gain=0.25, pll_pdi=0.001,
)


def make_trk_states(fs, channels, cfg=None) -> np.ndarray:
    """Initial per-channel state as ChannelL1CA holds it when tracking starts
    (channel_l1ca_borre.py:110-120, 231-251, 301-311).

    channels: iterable of dicts with prn, carrier_freq, start_sample and optionally
    iq_base (sample offset of the channel's recording in the device buffer) and iq_len.
    """
    c = dict(BORRE_DEFAULTS)
    if cfg:
        c.update({k: float(v) for k, v in cfg.items() if k in c})
    dll_t1, dll_t2 = loop_coefficients(c["dll_noise_bandwidth"], c["dll_damping_ratio"], c["dll_loop_gain"])
    pll_t1, pll_t2 = loop_coefficients(c["pll_noise_bandwidth"], c["pll_damping_ratio"], c["pll_loop_gain"])
    channels = list(channels)
    st = np.zeros(len(channels), dtype=L.TRK_STATE_DTYPE)
    code_step = CODE_FREQ / fs
    st["prn"] = [ch["prn"] for ch in channels]
    st["iq_base"] = [ch.get("iq_base", 0) for ch in channels]
    st["iq_len"] = [ch.get("iq_len", 0) for ch in channels]
    st["cur"] = [ch["start_sample"] for ch in channels]
    st["carrier_freq"] = [ch["carrier_freq"] for ch in channels]
    st["code_freq"] = CODE_FREQ
    st["code_step"] = code_step
    st["n_req"] = int(np.ceil((CODE_CHIPS - 0.0) / code_step))
    st["dll_tau1"], st["dll_tau2"], st["dll_pdi"] = dll_t1, dll_t2, c["dll_pdi"]
    st["pll_tau1"], st["pll_tau2"], st["pll_pdi"] = pll_t1, pll_t2, c["pll_pdi"]
    st["spacing"] = (c["correlator_early"], c["correlator_prompt"], c["correlator_late"])
    return st


def min_tap_gap(spacing: np.ndarray) -> float:
    """Smallest circular distance (chips) between the chip-boundary positions of two correlators:
    tap s changes chip where the code phase is congruent to -spacing_s modulo one chip."""
    gap = 1.0
    for row in np.atleast_2d(spacing):
        pos = np.unique(np.round(np.mod(-np.asarray(row, dtype=np.float64), 1.0), 12))
        if len(pos) > 1:
            d = np.diff(np.r_[pos, pos[0] + 1.0])
            gap = min(gap, float(d.min()))
    return gap


def _rec_channels(iq_base: np.ndarray) -> int:
    """Channel slots per recording when the channels come recording by recording in runs of equal length
    (the prefix-moment kernel then keeps every CTA on one recording); 0 otherwise."""
    b = np.asarray(iq_base)
    if len(b) == 0:
        return 0
    cuts = np.flatnonzero(b[1:] != b[:-1]) + 1
    runs = np.diff(np.concatenate(([0], cuts, [len(b)])))
    return int(runs[0]) if len(runs) > 1 and (runs == runs[0]).all() else 0


class TrackingEngine:
    """Closed-loop E/P/L tracking of many channels in one launch (K-TRK)."""

    def __init__(self, fs, states: np.ndarray, max_epochs: int, cluster=0, threads=0, use_tma=True, device=None,
                 dense=False, kernel=0, group=0):
        L.require_device()
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        L.check(L.load().sydr_set_device(self.device.index), "sydr_set_device")
        self.fs = float(fs)
        self.n_ch = len(states)
        self.max_epochs = int(max_epochs)
        st = np.ascontiguousarray(states)
        assert st.dtype == L.TRK_STATE_DTYPE
        # dense: 0 / False = latency instantiation, 1 / True = DENSE (several launches share the GPU), 2 = PACK (throughput
        # shape of the staged kernel: with cluster = 1, one CTA and one staged window per channel, two channels per SM)
        # kernel: 0 = automatic, 1 = prefix-moment kernel (throughput shape, `group` channels of a recording per CTA),
        # 2 = per-channel kernels only
        self.cfg = L.TrkConfig(int(cluster), int(threads), 1 if use_tma else 0, 0, 0, min_tap_gap(st["spacing"]), 0, 0,
                               int(dense), int(kernel), int(group), _rec_channels(st["iq_base"]), 0)
        self._states = torch.from_numpy(st.view(np.uint8).reshape(-1).copy()).to(self.device)
        self._out = torch.empty(self.n_ch * self.max_epochs * 128, dtype=torch.uint8, device=self.device)
        self._nep = torch.zeros(self.n_ch, dtype=torch.int32, device=self.device)
        # pinned staging for the results: one asynchronous D2H per fetch, no pageable bounce.  A ring of
        # buffers, so that fetch(copy=False) can hand out views that survive the next fetches.
        # (allocated on first use: a launch whose records stay on the device never pins host memory for them)
        self.RING = 4
        self._out_ring = [None] * self.RING
        self._ring_i = 0
        self._nep_host = torch.zeros(self.n_ch, dtype=torch.int32, pin_memory=True)

    def set_iq_len(self, iq_len):
        """Update the number of valid samples per channel (streaming: more data arrived)."""
        st = self.states()
        st["iq_len"] = iq_len
        self._states.copy_(torch.from_numpy(st.view(np.uint8).reshape(-1)))

    def reset(self, states: np.ndarray):
        """Load fresh channel states (same channel count) without reallocating."""
        st = np.ascontiguousarray(states)
        assert st.dtype == L.TRK_STATE_DTYPE and len(st) == self.n_ch
        self.cfg.min_tap_gap = min_tap_gap(st["spacing"])
        self.cfg.rec_channels = _rec_channels(st["iq_base"])
        self._states.copy_(torch.from_numpy(st.view(np.uint8).reshape(-1)))

    def launch(self, iq_dev: torch.Tensor, stream=None, iq_len: int = 0, append: bool = False, iq_base=None):
        """Enqueue one tracking launch.  iq_len > 0 limits every recording to its first iq_len
        samples (streaming upload); append=True continues the record arrays of earlier launches;
        iq_base (a multiple of 8, may be negative) places the recording inside `iq_dev` for this
        call: a sliding window holding samples [w0, w0 + len) passes iq_base = -w0, iq_len = w0 + len."""
        iq_dev = ensure_padded(iq_dev)
        self.cfg.iq_len = int(iq_len)
        self.cfg.append = 1 if append else 0
        self.cfg.use_iq_base = 0 if iq_base is None else 1
        self.cfg.iq_base = 0 if iq_base is None else int(iq_base)
        store_bytes = iq_dev.untyped_storage().nbytes() - iq_dev.storage_offset() * iq_dev.element_size()
        code = iq_code(iq_dev)
        alloc_samples = store_bytes // _IQ_BYTES[code]
        L.check(L.load().sydr_trk_run(iq_dev.data_ptr(), code, alloc_samples, self.fs, self._states.data_ptr(),
                                      self.n_ch, self._out.data_ptr(), self.max_epochs, self._nep.data_ptr(),
                                      C.byref(self.cfg), _stream_ptr(stream)), "sydr_trk_run")

    def states(self) -> np.ndarray:
        return self._states.cpu().numpy().view(L.TRK_STATE_DTYPE).copy()

    def fetch(self, copy: bool = True):
        """Per-channel record arrays of everything tracked since the last reset.  copy=False returns
        views of pinned staging memory (no host copy, no page faults: ~1 ms saved per 3 MB); they stay
        valid until RING - 1 further fetches of this engine."""
        self._ring_i = (self._ring_i + 1) % self.RING
        if self._out_ring[self._ring_i] is None:
            self._out_ring[self._ring_i] = torch.empty(self.n_ch * self.max_epochs * 128, dtype=torch.uint8, pin_memory=True)
        self._out_host = self._out_ring[self._ring_i]
        self._nep_host.copy_(self._nep, non_blocking=True)
        self._out_host.copy_(self._out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        nep = self._nep_host.numpy().copy()
        out = self._out_host.numpy().view(L.TRK_EPOCH_DTYPE).reshape(self.n_ch, self.max_epochs)
        if copy:
            return [out[c, :nep[c]].copy() for c in range(self.n_ch)]
        return [out[c, :nep[c]] for c in range(self.n_ch)]

    def run(self, iq_dev: torch.Tensor, stream=None):
        self.launch(iq_dev, stream=stream)
        res = self.fetch()
        st = self.states()
        if (st["status"] < 0).any():
            bad = np.nonzero(st["status"] < 0)[0].tolist()
            raise L.SydrError(f"tracking failed (code -4): aborted on channels {bad} (NCO state left the supported range, or a PRN without a code)")
        return res


def ensure_padded(iq_dev: torch.Tensor) -> torch.Tensor:
    """Return `iq_dev` or a copy whose storage extends IQ_PAD_BYTES past the last sample."""
    tail = iq_dev.untyped_storage().nbytes() - (iq_dev.storage_offset() + iq_dev.numel()) * iq_dev.element_size()
    if tail >= 256 and iq_dev.data_ptr() % 16 == 0 and iq_dev.is_contiguous():
        return iq_dev
    n = iq_dev.numel()
    buf = torch.zeros(n + IQ_PAD_BYTES // iq_dev.element_size(), dtype=iq_dev.dtype, device=iq_dev.device)
    buf[:n].copy_(iq_dev.reshape(-1))
    return buf[:n]


def epl_batch(iq_dev: torch.Tensor, fs: float, args: np.ndarray, stream=None) -> np.ndarray:
    """Independent open-loop EPL evaluations (sydr/dsp/tracking.py:92-116); returns (n, 6) float64."""
    L.require_device()
    a = np.ascontiguousarray(args)
    assert a.dtype == L.EPL_ARGS_DTYPE
    dev = iq_dev.device
    if iq_dev.dtype == torch.complex128:
        n = iq_dev.numel()
        conv = torch.empty(n + IQ_PAD_BYTES // 8, dtype=torch.complex64, device=dev)
        L.check(L.load().sydr_convert_to_f32(iq_dev.data_ptr(), L.IQ_F64, n, conv.data_ptr(), _stream_ptr(stream)),
                "sydr_convert_to_f32")
        iq_dev = conv[:n]
    iq_dev = ensure_padded(iq_dev)
    d_args = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
    d_out = torch.empty(len(a) * 6, dtype=torch.float64, device=dev)
    L.check(L.load().sydr_epl_batch(iq_dev.data_ptr(), iq_code(iq_dev), n_complex_samples(iq_dev), float(fs),
                                    d_args.data_ptr(), len(a), d_out.data_ptr(), _stream_ptr(stream)), "sydr_epl_batch")
    return d_out.cpu().numpy().reshape(len(a), 6)


# ======================================================================================
class NavBitEngine:
    """Bit synchronisation + 20-epoch prompt sums -> navigation bits on the device (K-NAV).

    Consumes the per-epoch records a TrackingEngine leaves in HBM and keeps, per channel, the
    state ChannelL1CA keeps between ticks (channel_l1ca_borre.py:398-413, 455-491), so only
    50 bit/s per channel have to cross PCIe.  Records may arrive in pieces of any length."""

    def __init__(self, n_channels: int, max_bits: int = 4096, device=None):
        L.require_device()
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_ch, self.max_bits = int(n_channels), int(max_bits)
        self._state = torch.empty(self.n_ch * L.NAV_STATE_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        self._bits = torch.zeros(self.n_ch * self.max_bits, dtype=torch.int8, device=self.device)
        self._sums = torch.zeros(self.n_ch * self.max_bits, dtype=torch.float64, device=self.device)
        self._nbits = torch.zeros(self.n_ch, dtype=torch.int32, device=self.device)
        self.reset()

    def reset(self):
        st = np.zeros(1, dtype=L.NAV_STATE_DTYPE)
        L.check(L.load().sydr_nav_state_init(st.ctypes.data), "sydr_nav_state_init")
        self._state.copy_(torch.from_numpy(np.repeat(st, self.n_ch).view(np.uint8).reshape(-1)))

    def launch_records(self, d_epochs: torch.Tensor, max_epochs: int, d_nepochs: torch.Tensor, first_epoch=0,
                       stream=None):
        """d_epochs: device bytes of [n_ch][max_epochs] sydr_trk_epoch; d_nepochs: int32 [n_ch]."""
        L.check(L.load().sydr_nav_bits(d_epochs.data_ptr(), int(max_epochs), d_nepochs.data_ptr(), int(first_epoch),
                                       self._state.data_ptr(), self.n_ch, self._bits.data_ptr(),
                                       self._sums.data_ptr(), self.max_bits, self._nbits.data_ptr(),
                                       _stream_ptr(stream)), "sydr_nav_bits")

    def launch(self, trk: "TrackingEngine", first_epoch=0, stream=None):
        """Consume records [first_epoch, n) of every channel of `trk` (enqueue after trk.launch).  A
        KaplanTrackingEngine's records are read with the Kaplan channel's bit-synchronisation rule."""
        assert trk.n_ch >= self.n_ch                      # the first n_ch slots of the tracking engine
        if hasattr(trk, "_kout"):
            L.check(L.load().sydr_nav_bits_kaplan(trk._out.data_ptr(), trk._kout.data_ptr(), trk.max_epochs,
                                                  trk._nep.data_ptr(), int(first_epoch), self._state.data_ptr(),
                                                  self.n_ch, self._bits.data_ptr(), self._sums.data_ptr(),
                                                  self.max_bits, self._nbits.data_ptr(), _stream_ptr(stream)),
                    "sydr_nav_bits_kaplan")
            return
        self.launch_records(trk._out, trk.max_epochs, trk._nep, first_epoch, stream)

    def states(self) -> np.ndarray:
        return self._state.cpu().numpy().view(L.NAV_STATE_DTYPE).copy()

    def fetch(self, want_sums=True):
        """Bits (and sums) produced by the last launch: list of (int8 array, float64 array) per channel."""
        nb = self._nbits.cpu().numpy()
        mx = int(nb.max()) if len(nb) else 0
        bits = self._bits.view(self.n_ch, self.max_bits)[:, :mx].cpu().numpy()
        sums = self._sums.view(self.n_ch, self.max_bits)[:, :mx].cpu().numpy() if want_sums else None
        return [(bits[c, :nb[c]].copy(), sums[c, :nb[c]].copy() if want_sums else None) for c in range(self.n_ch)]


# ======================================================================================
KAPLAN_DEFAULTS = dict(  # config/channels/channel_GPS_L1CA_kaplan.ini [TRACKING]
    correlator_epl_wide=0.5, correlator_epl_narrow=0.5, dll_threshold=10.0, dll_damping_ratio=0.7,
    dll_noise_bandwidth=2.0, dll_loop_gain=1.0, dll_pdi=0.001, pll_bandwidth_wide=25.0, pll_bandwidth_narrow=15.0,
    pll_threshold_wide=0.5, pll_threshold_narrow=0.8, fll_bandwidth_pullin=100.0, fll_bandwidth_wide=50.0,
    fll_bandwidth_narrow=15.0, fll_threshold_wide=0.5, fll_threshold_narrow=0.8,
)


def make_kaplan_states(fs, channels, cfg=None):
    """Initial states of ChannelL1CA_Kaplan channels when tracking starts (channel_l1ca_kaplan.py:262-340):
    (sydr_trk_state array, sydr_kaplan_state array)."""
    c = dict(KAPLAN_DEFAULTS)
    if cfg:
        c.update({k: float(v) for k, v in cfg.items() if k in c})
    if c["correlator_epl_wide"] != c["correlator_epl_narrow"]:
        raise L.SydrError("the device Kaplan loop needs equal wide and narrow correlator spacings "
                          "(use the host channel class otherwise)")
    channels = list(channels)
    w = c["correlator_epl_wide"]
    st = make_trk_states(fs, channels, dict(correlator_early=-w, correlator_prompt=0.0, correlator_late=w,
                                            dll_damping_ratio=c["dll_damping_ratio"],
                                            dll_noise_bandwidth=c["dll_noise_bandwidth"],
                                            dll_loop_gain=c["dll_loop_gain"], dll_pdi=c["dll_pdi"]))
    ks = np.zeros(len(channels), dtype=L.KAPLAN_STATE_DTYPE)
    ks["fll_bw_pullin"], ks["fll_bw_wide"], ks["fll_bw_narrow"] = (c["fll_bandwidth_pullin"], c["fll_bandwidth_wide"],
                                                                   c["fll_bandwidth_narrow"])
    ks["pll_bw_wide"], ks["pll_bw_narrow"] = c["pll_bandwidth_wide"], c["pll_bandwidth_narrow"]
    ks["fll_thr_wide"], ks["fll_thr_narrow"] = c["fll_threshold_wide"], c["fll_threshold_narrow"]
    ks["pll_thr_narrow"], ks["dll_threshold"] = c["pll_threshold_narrow"], c["dll_threshold"]
    ks["fll_bw"], ks["pll_bw"] = c["fll_bandwidth_pullin"], c["pll_bandwidth_wide"]      # L326-327
    ks["lock_state"] = 1                                                                 # PULL_IN, L333
    return st, ks


class KaplanTrackingEngine(TrackingEngine):
    """Closed-loop tracking with the Kaplan loop closure on the device (K-TRK, KAP instantiation):
    FLL-assisted PLL, lock indicators, C/N0, code lock / bit sync, PULL_IN -> WIDE -> NARROW."""

    def __init__(self, fs, states: np.ndarray, kstates: np.ndarray, max_epochs: int, cluster=0, threads=0,
                 use_tma=True, device=None):
        super().__init__(fs, states, max_epochs, cluster=cluster, threads=threads, use_tma=use_tma, device=device)
        ks = np.ascontiguousarray(kstates)
        assert ks.dtype == L.KAPLAN_STATE_DTYPE and len(ks) == self.n_ch
        self._kstates = torch.from_numpy(ks.view(np.uint8).reshape(-1).copy()).to(self.device)
        self._kout = torch.zeros(self.n_ch * self.max_epochs * 40, dtype=torch.uint8, device=self.device)

    def launch(self, iq_dev: torch.Tensor, stream=None, iq_len: int = 0, append: bool = False, iq_base=None):
        iq_dev = ensure_padded(iq_dev)
        self.cfg.iq_len = int(iq_len)
        self.cfg.append = 1 if append else 0
        self.cfg.use_iq_base = 0 if iq_base is None else 1
        self.cfg.iq_base = 0 if iq_base is None else int(iq_base)
        store_bytes = iq_dev.untyped_storage().nbytes() - iq_dev.storage_offset() * iq_dev.element_size()
        code = iq_code(iq_dev)
        L.check(L.load().sydr_trk_run_kaplan(iq_dev.data_ptr(), code, store_bytes // _IQ_BYTES[code], self.fs,
                                             self._states.data_ptr(), self._kstates.data_ptr(), self.n_ch,
                                             self._out.data_ptr(), self._kout.data_ptr(), self.max_epochs,
                                             self._nep.data_ptr(), C.byref(self.cfg), _stream_ptr(stream)),
                "sydr_trk_run_kaplan")

    def kaplan_states(self) -> np.ndarray:
        return self._kstates.cpu().numpy().view(L.KAPLAN_STATE_DTYPE).copy()

    def fetch_kaplan(self):
        """Per-channel arrays of the Kaplan extras (fll, cn0, lock indicators, lock state, flags), aligned
        with the records of fetch()."""
        nep = self._nep.cpu().numpy()
        out = self._kout.cpu().numpy().view(L.KAPLAN_EPOCH_DTYPE).reshape(self.n_ch, self.max_epochs)
        return [out[c, :nep[c]].copy() for c in range(self.n_ch)]
