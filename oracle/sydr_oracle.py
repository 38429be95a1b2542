"""CPU oracle for the SyDR DSP hot paths (PCPS acquisition + E/P/L tracking).

TEST INFRASTRUCTURE ONLY.  This module is a float64/complex128 NumPy restatement of
the reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
(``sydr_b200``) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference
from ``/root/reference`` (stubbing ``matplotlib`` and ``gps_time`` only) and stores its
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function
here against those fixtures, against the IS-GPS-200 first-10-chip table, against the
5-sample replica known-answer vector of ``sydr/c_functions/tracking.c:243-247`` and (when
built) against ``oracle/_ref/tracking.so`` compiled from the reference's own C sources.
The reference pins no results for ``numpy.fft`` (pocketfft) itself; parity is therefore
"equals what the reference functions return on the same input" (SURVEY.md §8c).

Each function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math

import numpy as np

# sydr/utils/constants.py:4-6 -- the GPS ICD value of pi, used by PLL_costa only.
GPS_PI = 3.1415926535898
GPS_TWO_PI = GPS_PI * 2.0
# sydr/utils/constants.py -- GPS L1 C/A code
CODE_FREQ = 1.023e6
CODE_CHIPS = 1023

# IS-GPS-200 Table 3-Ia, G2 code delay (chips) for PRN 1..37 (public ICD data; the
# reference carries the same table at sydr/signal/ca.py:13-68).
G2_DELAY = (
    5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258,
    469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862,
    863, 950, 947, 948, 950,
)
# IS-GPS-200 Table 3-Ia, first 10 chips (octal) of PRN 1..32 -- known-answer vector.
FIRST10_OCTAL = (
    0o1440, 0o1620, 0o1710, 0o1744, 0o1133, 0o1455, 0o1131, 0o1454,
    0o1626, 0o1504, 0o1642, 0o1750, 0o1764, 0o1772, 0o1775, 0o1776,
    0o1156, 0o1467, 0o1633, 0o1715, 0o1746, 0o1763, 0o1063, 0o1706,
    0o1743, 0o1761, 0o1770, 0o1774, 0o1127, 0o1453, 0o1625, 0o1712,
)


# --------------------------------------------------------------------------------------
# C/A code  (sydr/signal/ca.py:70-112, sydr/signal/gnsssignal.py:9-31)
# --------------------------------------------------------------------------------------
def _lfsr(taps):
    """10-stage Fibonacci LFSR, all-ones start, output = stage 10 (ca.py:70-91)."""
    reg = [1] * 10
    out = np.zeros(CODE_CHIPS, dtype=np.int8)
    for i in range(CODE_CHIPS):
        out[i] = reg[9]
        fb = 0
        for t in taps:
            fb ^= reg[t - 1]
        reg = [fb] + reg[:9]
    return out


_G1 = _lfsr((3, 10))
_G2 = _lfsr((2, 3, 6, 8, 9, 10))


def ca_chips(prn: int) -> np.ndarray:
    """0/1 chips of PRN `prn`: G1 xor (G2 delayed by G2_DELAY chips) (ca.py:93-104)."""
    d = G2_DELAY[prn - 1]
    g2 = np.concatenate([_G2[CODE_CHIPS - d:], _G2[:CODE_CHIPS - d]])
    return (_G1 ^ g2).astype(np.int8)


def ca_code(prn: int) -> np.ndarray:
    """+-1.0 float64 code, chip 1 -> +1.0 (ca.py:106-112 `2.0*x - 1.0`;
    gnsssignal.py:24 GenerateGPSGoldCode)."""
    return 2.0 * ca_chips(prn).astype(np.float64) - 1.0


def padded_code(prn: int) -> np.ndarray:
    """[c1022, c0..c1022, c0] (channel_l1ca_borre.py:170-173)."""
    c = ca_code(prn)
    return np.r_[c[-1], c, c[0]]


def first10_octal(prn: int) -> int:
    """ca.py:135-141 first_10_chips."""
    r = 0
    for b in ca_chips(prn)[:10]:
        r = 2 * r + int(b)
    return r


def samples_per_code(fs: float) -> int:
    """gnsssignal.py:62-70 (Python round = half-to-even)."""
    return round(fs / (CODE_FREQ / CODE_CHIPS))


def samples_per_chip(fs: float) -> int:
    """channel_l1ca_borre.py:284."""
    return round(fs / CODE_FREQ)


def upsample_code(code: np.ndarray, fs: float) -> np.ndarray:
    """gnsssignal.py:35-58: idx = trunc(ts * k / tc), ts = 1/fs, tc = 1/1.023e6."""
    ts = 1 / fs
    tc = 1 / CODE_FREQ
    n = samples_per_code(fs)
    idx = np.trunc(ts * np.array(range(n)) / tc).astype(int)
    return code[idx]


def code_spectrum(prn: int, fs: float) -> np.ndarray:
    """channel_l1ca_borre.py:281-282: conj(fft(UpsampleCode(code)))."""
    return np.conj(np.fft.fft(upsample_code(ca_code(prn), fs)))


# --------------------------------------------------------------------------------------
# Acquisition  (sydr/dsp/acquisition.py)
# --------------------------------------------------------------------------------------
def doppler_bins(doppler_range: float, doppler_step: float) -> np.ndarray:
    """acquisition.py:34."""
    return np.arange(-doppler_range, doppler_range + 1, doppler_step)


def pcps(rf, inter_freq, fs, code_fft, doppler_range, doppler_step, n_code,
         coh=1, noncoh=1) -> np.ndarray:
    """acquisition.py:31-74.  Returns float64 (bins, n_code)."""
    rf = np.squeeze(rf)
    phase_points = np.array(range(coh * n_code)) * 2 * np.pi / fs      # L33
    bins = doppler_bins(doppler_range, doppler_step)                   # L34
    cmap = np.zeros((len(bins), n_code))
    for b, f in enumerate(bins):
        f = inter_freq - f                                             # L42
        carrier = np.exp(-1j * f * phase_points)                       # L45
        nc = np.zeros(n_code)
        for i_nc in range(noncoh):
            seg = rf[i_nc * coh * n_code:(i_nc + 1) * coh * n_code]    # L51
            seg = carrier * seg                                        # L53
            cs = np.zeros(n_code, dtype=np.complex128)
            for i_c in range(coh):
                x = np.fft.fft(seg[i_c * n_code:(i_c + 1) * n_code])   # L59
                cs = cs + np.fft.ifft(x * code_fft)                    # L62,65
            nc = nc + np.abs(cs)                                       # L68
        cmap[b, :] = np.abs(nc)                                        # L70
    return cmap


def second_peak_range(code_idx: int, n_code: int, chip: int):
    """Index set searched for the second peak (acquisition.py:103-110), as a list of
    half-open (lo, hi) ranges.  Quirks kept: the last sample N-1 is never searched in
    the first and third branch; `exclude[0] < 1` (not `< 0`)."""
    e0 = int(code_idx - chip)
    e1 = int(code_idx + chip)
    if e0 < 1:
        return [(e1, n_code - 1)]
    if e1 >= n_code:
        return [(0, e0)]
    return [(0, e0), (e1, n_code - 1)]


def two_peak(cmap: np.ndarray, n_code: int, chip: int):
    """acquisition.py:97-115 TwoCorrelationPeakComparison."""
    fi, ci = np.unravel_index(cmap.argmax(), cmap.shape)               # L98
    fi, ci = int(fi), int(ci)
    p1 = cmap[fi, ci]
    idx = []
    for lo, hi in second_peak_range(ci, n_code, chip):
        idx += list(range(lo, hi))
    p2 = np.amax(cmap[fi, idx])                                        # L111
    return [fi, ci], p1 / p2                                           # L113


def acquisition_handoff(freq_idx, code_idx, inter_freq, doppler_range, doppler_step,
                        current_sample, acq_required, track_required):
    """channel_l1ca_borre.py:301-311: scalars handed from acquisition to tracking."""
    doppler = -((-doppler_range) + doppler_step * freq_idx)
    code_offset = int(np.round(code_idx))
    carrier = inter_freq + doppler
    cur = current_sample + acq_required
    cur -= track_required
    cur += code_offset + 1
    return carrier, code_offset, cur


# --------------------------------------------------------------------------------------
# Tracking  (sydr/dsp/tracking.py)
# --------------------------------------------------------------------------------------
def generate_replica(time, n, carrier_freq, rem_carrier):
    """tracking.py:8-17 (np.pi here; the C twin tracking.c:31-52 uses GPS pi)."""
    time = time[0:n + 1]
    temp = -(carrier_freq * 2.0 * np.pi * time) + rem_carrier
    rem = temp[n] % (2 * np.pi)
    return np.exp(1j * temp[:n]), rem


def code_indices(rem_code, spacing, code_step, n) -> np.ndarray:
    """tracking.py:110-112: ceil(linspace(shift, codeStep*n + shift, n, endpoint=False)).
    NumPy's linspace evaluates fl(fl(i*step') + start), step' = fl(fl(stop-start)/n)."""
    shift = rem_code + spacing
    return np.ceil(np.linspace(shift, code_step * n + shift, n, endpoint=False)).astype(int)


def code_indices_explicit(rem_code, spacing, code_step, n) -> np.ndarray:
    """Same as code_indices with the linspace arithmetic written out (what the CUDA
    kernel and tracking.c:81-89 evaluate)."""
    start = rem_code + spacing
    stop = code_step * n + start
    step = (stop - start) / n
    return np.ceil(np.arange(n, dtype=np.float64) * step + start).astype(int)


def epl(rf, code, fs, carrier_freq, rem_carrier, rem_code, code_step, spacings):
    """tracking.py:92-116.  `code` is the 1025-entry padded code.  Returns
    [IE, QE, IP, QP, IL, QL] float64."""
    rf = np.squeeze(rf)
    n = len(rf)
    t = np.arange(0.0, n) / fs                                          # L101
    replica = np.exp(1j * (-(carrier_freq * 2.0 * np.pi * t) + rem_carrier))  # L102
    sig = replica * rf                                                  # L105
    i_sig = np.real(sig)
    q_sig = np.imag(sig)
    out = [0.0] * (2 * len(spacings))
    for i, sp in enumerate(spacings):
        idx = code_indices(rem_code, sp, code_step, n)                  # L110-112
        out[2 * i] = np.sum(code[idx] * i_sig)                          # L113
        out[2 * i + 1] = np.sum(code[idx] * q_sig)                      # L114
    return out


def loop_coefficients(bw, damping, gain):
    """tracking.py:56-61."""
    wn = bw * 8.0 * damping / (4.0 * damping ** 2 + 1)
    return gain / wn ** 2, 2.0 * damping / wn


def dll_nneml(ie, qe, il, ql):
    """tracking.py:126-127."""
    return (np.sqrt(ie ** 2 + qe ** 2) - np.sqrt(il ** 2 + ql ** 2)) / \
           (np.sqrt(ie ** 2 + qe ** 2) + np.sqrt(il ** 2 + ql ** 2))


def pll_costa(ip, qp):
    """tracking.py:139-140 (GPS TWO_PI)."""
    return np.arctan(qp / ip) / GPS_TWO_PI


def borre_filter(x, mem, tau1, tau2, pdi):
    """tracking.py:183-184."""
    out = tau2 / tau1 * (x - mem)
    out += pdi / tau1 * x
    return out


# Default loop parameters, config/channels/channel_GPS_L1CA_borre.ini:15-29.
BORRE_INI = dict(
    spacings=(-0.5, 0.0, 0.5),
    dll=dict(bw=1.0, damping=0.7, gain=1.0, pdi=0.001),
    pll=dict(bw=8.0, damping=0.7, gain=0.25, pdi=0.001),
)


class BorreTrackOracle:
    """Scalar state machine of ChannelL1CA.runTracking (channel_l1ca_borre.py:333-451)
    run over a whole recording held in memory (no 100 ms ring, no per-ms ticks: those only
    decide *when* an epoch runs, not its result).  One instance = one channel."""

    def __init__(self, prn, fs, carrier_freq, start_sample, ini=BORRE_INI):
        self.fs = float(fs)
        self.code = padded_code(prn)
        self.spacings = list(ini["spacings"])
        d, p = ini["dll"], ini["pll"]
        self.dll_tau1, self.dll_tau2 = loop_coefficients(d["bw"], d["damping"], d["gain"])
        self.pll_tau1, self.pll_tau2 = loop_coefficients(p["bw"], p["damping"], p["gain"])
        self.dll_pdi, self.pll_pdi = d["pdi"], p["pdi"]
        # channel_l1ca_borre.py:110-120, 250-251
        self.code_freq = CODE_FREQ
        self.carrier_freq = float(carrier_freq)
        self.rem_code = 0.0
        self.rem_carrier = 0.0
        self.nco_code = 0.0
        self.nco_code_err = 0.0
        self.nco_carrier = 0.0
        self.nco_carrier_err = 0.0
        self.code_step = CODE_FREQ / self.fs
        self.n_req = int(np.ceil((CODE_CHIPS - self.rem_code) / self.code_step))
        self.cur = int(start_sample)

    def state(self):
        return dict(cur=self.cur, n=self.n_req, carrier_freq=self.carrier_freq,
                    rem_carrier=self.rem_carrier, rem_code=self.rem_code,
                    code_step=self.code_step, code_freq=self.code_freq,
                    nco_code_err=self.nco_code_err, nco_carrier_err=self.nco_carrier_err)

    def step(self, rf_all, corr_override=None):
        """One epoch (channel_l1ca_borre.py:354-429).  `corr_override` lets a test
        teacher-force the six sums."""
        n = self.n_req
        rf = rf_all[self.cur:self.cur + n]
        corr = epl(rf, self.code, self.fs, self.carrier_freq, self.rem_carrier,
                   self.rem_code, self.code_step, self.spacings)           # L354-361
        used = corr if corr_override is None else list(corr_override)
        self.rem_carrier -= self.carrier_freq * 2.0 * np.pi * n / self.fs    # L364
        self.rem_carrier %= (2 * np.pi)                                      # L365
        ie, qe, ip, qp, il, ql = used
        code_err = dll_nneml(ie, qe, il, ql)                                 # L383
        self.nco_code = borre_filter(code_err, self.nco_code_err, self.dll_tau1,
                                     self.dll_tau2, self.dll_pdi)            # L385-387
        self.nco_code_err = code_err
        ph_err = pll_costa(ip, qp)                                           # L391
        self.nco_carrier = borre_filter(ph_err, self.nco_carrier_err, self.pll_tau1,
                                        self.pll_tau2, self.pll_pdi)         # L393-395
        self.nco_carrier_err = ph_err
        self.code_freq -= self.nco_code                                      # L422
        self.carrier_freq += self.nco_carrier                                # L423
        self.rem_code += n * self.code_step - CODE_CHIPS                     # L424
        self.code_step = self.code_freq / self.fs                            # L425
        self.cur = self.cur + n                                              # L428 (no ring)
        self.n_req = int(np.ceil((CODE_CHIPS - self.rem_code) / self.code_step))  # L429
        return dict(corr=corr, dll=self.nco_code, pll=self.nco_carrier,
                    carrier_frequency=self.carrier_freq, code_frequency=self.code_freq,
                    code_err=code_err, carrier_err=ph_err, n=n)

    def run(self, rf_all, max_epochs=None):
        out = []
        while self.cur + self.n_req <= len(rf_all):
            if max_epochs is not None and len(out) >= max_epochs:
                break
            out.append(self.step(rf_all))
        return out


# ------------------------------------------------------------------------------------------
# Bit synchronisation + navigation-bit accumulation (the scalar step right after the
# correlators).  Plain Python loop, one iteration per tracking epoch.
# ------------------------------------------------------------------------------------------
LNAV_MS_PER_BIT = 20            # sydr/utils/constants.py
MIN_CONVERGENCE_TIME = 100      # sydr/channel/channel_l1ca_borre.py:30


class NavBitOracle:
    """What ChannelL1CA does with the prompt of every tracking epoch
    (sydr/channel/channel_l1ca_borre.py:367-373 prompt history, L398-413 bit-sync test, L414-419
    flags and counters, L455-491 runDecoding up to Prompt2Bit, L577-591 resetPrompt, L626-627
    history wrap; sydr/dsp/decoding.py:16-27 Prompt2Bit)."""

    def __init__(self):
        self.code_lock = False
        self.bit_sync = False
        self.code_counter = 0
        self.i_prompt = 0.0
        self.nb_prompt = 0
        self.history = np.zeros(LNAV_MS_PER_BIT)            # correlatorsBuffer[:, IDX_I_PROMPT]
        self.nav_sum = 0.0
        self.nav_count = 0
        self.bits = []
        self.sums = []
        self.sync_epoch = -1

    def step(self, ip: float):
        # runTracking
        self.history[self.nb_prompt] = ip                                       # L367
        self.nb_prompt += 1                                                     # L372
        if not self.bit_sync:                                                   # L400
            if self.code_lock and self.code_counter > MIN_CONVERGENCE_TIME \
                    and np.sign(self.i_prompt) != np.sign(ip):                  # L402-404
                self.bit_sync = True
                self.sync_epoch = self.code_counter
                self.nb_prompt = 0                                              # resetPrompt, L591
        self.code_lock = True                                                   # L416
        self.i_prompt = ip                                                      # L417
        self.code_counter += 1                                                  # L419
        # runDecoding
        if not self.bit_sync:                                                   # L472-476
            self.nav_sum = 0.0
            self.nav_count = 0
        else:
            self.nav_sum += self.history[self.nb_prompt - 1]                    # L479 (index -1 on the sync epoch)
            self.nav_count += 1
            if self.nav_count == LNAV_MS_PER_BIT:                               # L483
                self.bits.append(1 if self.nav_sum > 0 else 0)                  # Prompt2Bit
                self.sums.append(self.nav_sum)
                self.nav_sum = 0.0
                self.nav_count = 0
        # _processHandler tail
        if self.nb_prompt == LNAV_MS_PER_BIT:                                   # L626-627
            self.nb_prompt = 0


def nav_bits(i_prompts):
    """Bits, their 20-epoch sums and the synchronisation epoch for a prompt sequence."""
    o = NavBitOracle()
    for v in i_prompts:
        o.step(float(v))
    return np.array(o.bits, dtype=np.int8), np.array(o.sums, dtype=np.float64), o.sync_epoch, o


# ------------------------------------------------------------------------------------------
# Kaplan channel: FLL-assisted PLL, lock indicators, C/N0, PULL_IN / WIDE / NARROW machine.
# ------------------------------------------------------------------------------------------
GPS_HALF_PI = GPS_PI / 2.0
W0_BANDWIDTH_1 = 0.25           # sydr/utils/constants.py:80
W0_BANDWIDTH_2 = 0.53           # sydr/utils/constants.py:81
W0_SCALE_A2 = 1.414             # sydr/utils/constants.py:83
PULL_IN, WIDE_TRACK, NARROW_TRACK = 1, 2, 3        # sydr/utils/enumerations.py LoopLockState
FLAG_CODE_LOCK, FLAG_BIT_SYNC = 1, 2               # sydr/utils/enumerations.py TrackingFlags
KAPLAN_INI = {   # config/channels/channel_GPS_L1CA_kaplan.ini [TRACKING]
    "correlator_epl_wide": 0.5, "correlator_epl_narrow": 0.5, "dll_threshold": 10.0, "dll_damping_ratio": 0.7,
    "dll_noise_bandwidth": 2.0, "dll_loop_gain": 1.0, "dll_pdi": 0.001, "pll_bandwidth_wide": 25.0,
    "pll_bandwidth_narrow": 15.0, "pll_threshold_wide": 0.5, "pll_threshold_narrow": 0.8,
    "fll_bandwidth_pullin": 100.0, "fll_bandwidth_wide": 50.0, "fll_bandwidth_narrow": 15.0,
    "fll_threshold_wide": 0.5, "fll_threshold_narrow": 0.8}


def fll_atan(ip, qp, ip_prev, qp_prev, dt):
    """sydr/dsp/tracking.py:156-176 (FLL_ATAN + phase_unwrap)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.arctan(np.float64(qp) / np.float64(ip)) - np.arctan(np.float64(qp_prev) / np.float64(ip_prev))
    if np.isnan(e):
        e = 0.0
    if e >= GPS_HALF_PI:
        e = e - GPS_PI
    elif e <= -GPS_HALF_PI:
        e = e + GPS_PI
    return e / dt / GPS_TWO_PI


class KaplanTrackOracle:
    """Scalar state machine of ChannelL1CA_Kaplan.runTracking (sydr/channel/channel_l1ca_kaplan.py:342-619)
    over a recording held in memory; `corr_override` teacher-forces the six correlator sums."""

    def __init__(self, prn, fs, carrier_freq, start_sample, ini=KAPLAN_INI):
        c = {k: float(v) for k, v in ini.items()}
        self.c = c
        self.fs = float(fs)
        self.code = padded_code(prn)
        self.spacings = [-c["correlator_epl_wide"], 0.0, c["correlator_epl_wide"]]
        self.dll_tau1, self.dll_tau2 = loop_coefficients(c["dll_noise_bandwidth"], c["dll_damping_ratio"],
                                                         c["dll_loop_gain"])
        self.code_freq = CODE_FREQ
        self.carrier_freq = float(carrier_freq)
        self.rem_code = 0.0
        self.rem_carrier = 0.0
        self.code_step = CODE_FREQ / self.fs
        self.n_req = int(np.ceil((CODE_CHIPS - self.rem_code) / self.code_step))
        self.cur = int(start_sample)
        self.ip_prev = self.qp_prev = 0.0
        self.dll_discrim = 0.0
        self.fll_lock = self.pll_lock = self.dll_lock = 0.0
        self.cn0 = 0.0
        self.pdpn = 0.0
        self.accum_counter = 0
        self.vel_memory = 0.0
        self.fll_bw = c["fll_bandwidth_pullin"]
        self.pll_bw = c["pll_bandwidth_wide"]
        self.state = PULL_IN
        self.flags = 0
        self.code_counter = 0

    def step(self, rf_all, corr_override=None):
        c, n = self.c, self.n_req
        corr = None
        if corr_override is None or rf_all is not None:
            corr = epl(rf_all[self.cur:self.cur + n], self.code, self.fs, self.carrier_freq, self.rem_carrier,
                       self.rem_code, self.code_step, self.spacings)                               # L376-384
        ie, qe, ip, qp, il, ql = (corr if corr_override is None else [np.float64(v) for v in corr_override])
        if self.accum_counter == LNAV_MS_PER_BIT:                                                  # L387-389
            self.accum_counter = 0
        self.accum_counter += 1                                                                    # L393
        fll = pll = 0.0                                                                            # L409-432
        if self.state == PULL_IN:
            if self.code_counter > 1:
                fll = fll_atan(ip, qp, self.ip_prev, self.qp_prev, 1e-3)
        else:
            fll = fll_atan(ip, qp, self.ip_prev, self.qp_prev, 1e-3)
            pll = pll_costa(ip, qp)
        dll = dll_nneml(ie, qe, il, ql)
        w0f, w0p = self.fll_bw / W0_BANDWIDTH_1, self.pll_bw / W0_BANDWIDTH_2                       # L443-444
        update = (pll * w0p ** 2 + fll * w0f) * (1 * 1e-3)                                          # tracking.py:270
        carrier_err = update + self.vel_memory
        self.vel_memory = update
        carrier_err += pll * W0_SCALE_A2 * w0p                                                      # tracking.py:275
        code_err = borre_filter(dll, self.dll_discrim, self.dll_tau1, self.dll_tau2, c["dll_pdi"] * 1)   # L453
        if self.code_counter != 0:                                                                  # L463-506
            lock = ip * self.ip_prev - qp * self.qp_prev
            lock *= np.sign(ip * self.ip_prev + qp * self.qp_prev)
            lock /= (ip ** 2 + qp ** 2)
            lock = abs(lock)
            self.fll_lock = (1 - 0.005) * self.fll_lock + 0.005 * lock                              # lockindicator.py:6-24
            if self.state > PULL_IN:
                nbd, nbp = ip ** 2 - qp ** 2, ip ** 2 + qp ** 2
                self.pll_lock = (1 - 0.005) * self.pll_lock + 0.005 * (nbd / nbp)                   # lockindicator.py:28-44
            with np.errstate(divide="ignore"):
                self.pdpn += (ip ** 2 + qp ** 2) / (abs(ip) - abs(qp)) ** 2
            if self.accum_counter == LNAV_MS_PER_BIT:
                lam = 1 / (self.pdpn / self.accum_counter)                                          # lockindicator.py:76-99
                new = lam * (1 / (self.accum_counter * 1e-3))
                self.cn0 = (1 - 0.1) * self.cn0 + 0.1 * new
                self.pdpn = 0.0
            self.dll_lock = self.cn0
        self.code_counter += 1                                                                      # L515
        self.dll_discrim = dll
        self.rem_carrier -= self.carrier_freq * GPS_TWO_PI * n / self.fs                            # L529
        self.rem_carrier %= GPS_TWO_PI
        self.code_freq -= code_err
        self.carrier_freq += carrier_err
        self.rem_code += n * self.code_step - CODE_CHIPS
        self.code_step = self.code_freq / self.fs
        self.cur += n
        self.n_req = int(np.ceil((CODE_CHIPS - self.rem_code) / self.code_step))
        # trackingStateUpdate, L545-619
        if self.state != PULL_IN and self.dll_lock > c["dll_threshold"] and not (self.flags & FLAG_CODE_LOCK):
            self.flags |= FLAG_CODE_LOCK
        elif self.dll_lock < c["dll_threshold"] and (self.flags & FLAG_CODE_LOCK):
            self.flags ^= FLAG_CODE_LOCK
        if (self.flags & FLAG_CODE_LOCK) and not (self.flags & FLAG_BIT_SYNC):
            if np.sign(self.ip_prev) != np.sign(ip):
                self.flags |= FLAG_BIT_SYNC
                self.accum_counter = 1
                self.pdpn = 0.0
        self.ip_prev, self.qp_prev = ip, qp
        if self.state != NARROW_TRACK and self.fll_lock >= c["fll_threshold_narrow"] \
                and self.pll_lock >= c["pll_threshold_narrow"]:
            self.state, self.fll_bw, self.pll_bw = NARROW_TRACK, c["fll_bandwidth_narrow"], c["pll_bandwidth_narrow"]
        elif self.state != WIDE_TRACK and c["fll_threshold_wide"] <= self.fll_lock < c["fll_threshold_narrow"]:
            self.state, self.fll_bw, self.pll_bw = WIDE_TRACK, c["fll_bandwidth_wide"], c["pll_bandwidth_wide"]
        elif self.state != PULL_IN and self.fll_lock <= c["fll_threshold_wide"]:
            self.state, self.fll_bw, self.pll_bw = PULL_IN, c["fll_bandwidth_pullin"], 0.0
        return dict(corr=corr, dll=dll, pll=pll, fll=fll, carrier_frequency=self.carrier_freq,
                    code_frequency=self.code_freq, carrier_frequency_error=carrier_err, code_frequency_error=code_err,
                    cn0=self.cn0, pll_lock=self.pll_lock, fll_lock=self.fll_lock, lock_state=self.state,
                    flags=self.flags, rem_code=self.rem_code, rem_carrier=self.rem_carrier, n=n, n_req=self.n_req,
                    cur=self.cur)


def nav_bits_kaplan(i_prompts, bit_sync_flags):
    """The Kaplan channel's decodeBit (sydr/channel/channel_l1ca_kaplan.py:725-758) over a sequence of epochs:
    `bit_sync_flags[k]` = BIT_SYNC after epoch k's trackingStateUpdate.  Returns bits, their sums, the
    synchronisation epoch, and the pending (sum, count)."""
    bits, sums, sync = [], [], -1
    nav_sum, count = 0.0, 0
    for k, (ip, fl) in enumerate(zip(i_prompts, bit_sync_flags)):
        if not fl:                                                              # L733-737
            nav_sum, count = 0.0, 0
            continue
        if sync < 0:
            sync = k
        nav_sum += float(ip)                                                    # L740-741
        count += 1
        if count == LNAV_MS_PER_BIT:                                            # L744-752
            bits.append(1 if nav_sum > 0 else 0)
            sums.append(nav_sum)
            nav_sum, count = 0.0, 0
    return np.array(bits, dtype=np.int8), np.array(sums, dtype=np.float64), sync, (nav_sum, count)
