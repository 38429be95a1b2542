"""Synthetic GPS L1 C/A IQ recordings (host-side test-signal tool, not on the hot path).

The reference ships no generator and no recording (config/receiver.ini:17 points at a
private file), so every parity test and benchmark uses this model:

    x[n] = sum_p A_p c_p[floor(f_code,p t - tau_p) mod 1023] d_p[floor((f_code,p t - tau_p)/20460)]
                 exp(j (2 pi f_D,p t + phi_p))  +  w[n],        t = n / fs
    f_code,p = 1.023e6 (1 + f_D,p / 1575.42e6),   A = sigma sqrt(2 10^(CN0/10) / fs)

quantised with clip(round(.)) to int8 (sigma = 16 LSB) or int16 (sigma = 2048 LSB) and laid
out as interleaved I,Q -- byte for byte what `RFSignal.readFile` parses
(sydr/signal/rfsignal.py:107-130).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

L1_FREQ = 1575.42e6
CODE_FREQ = 1.023e6
CODE_CHIPS = 1023

_G2_DELAY = (5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258,
             469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862)


def _lfsr(taps):
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


def ca_code_pm1(prn: int) -> np.ndarray:
    """+-1 int8 C/A code of PRN 1..32 (chip '1' -> +1, as the reference)."""
    d = _G2_DELAY[prn - 1]
    g2 = np.roll(_G2, d)
    return (2 * (_G1 ^ g2) - 1).astype(np.int8)


@dataclass
class Sat:
    prn: int
    doppler: float          # Hz
    delay_chips: float      # code delay tau in chips, [0, 1023)
    phase: float = 0.0      # rad
    cn0: float = 45.0       # dB-Hz


@dataclass
class Scenario:
    fs: float
    nbits: int              # 8 or 16
    duration: float         # seconds
    sats: list
    seed: int
    inter_freq: float = 0.0
    sigma: float = field(default=0.0)

    def __post_init__(self):
        if not self.sigma:
            self.sigma = 16.0 if self.nbits == 8 else 2048.0

    @property
    def n_samples(self) -> int:
        return int(round(self.fs * self.duration))

    @property
    def dtype(self):
        return np.int8 if self.nbits == 8 else np.int16


def snapped_dopplers(rng, n, doppler_step, span=4500.0, jitter=30.0):
    """Dopplers uniform in +-span snapped to within +-jitter of an acquisition bin centre
    (so the 8 Hz PLL of the Borre ini pulls in)."""
    centres = np.round(rng.uniform(-span, span, n) / doppler_step) * doppler_step
    return centres + rng.uniform(-jitter, jitter, n)


def make_scenario(fs, nbits, duration, prns, seed, doppler_step=250.0, cn0=45.0) -> Scenario:
    rng = np.random.default_rng(seed)
    dop = snapped_dopplers(rng, len(prns), doppler_step)
    tau = rng.uniform(0.0, CODE_CHIPS, len(prns))
    ph = rng.uniform(0.0, 2 * np.pi, len(prns))
    sats = [Sat(int(p), float(d), float(t), float(h), cn0) for p, d, t, h in zip(prns, dop, tau, ph)]
    return Scenario(fs=fs, nbits=nbits, duration=duration, sats=sats, seed=seed)


PRNS_8 = (3, 7, 11, 14, 19, 22, 27, 31)
PRNS_12 = (1, 3, 7, 8, 11, 14, 17, 19, 22, 27, 30, 31)


def baseline_scenario(cfg: int, duration: float | None = None, recording: int = 0) -> Scenario:
    """The five BASELINE.json configurations (BASELINE.md 'Synthetic inputs')."""
    if cfg == 1:
        return make_scenario(4e6, 8, 10.0 if duration is None else duration, PRNS_8, 1001, 100.0)
    if cfg == 2:
        return make_scenario(10e6, 8, 0.010 if duration is None else duration, PRNS_8, 1002, 250.0)
    if cfg == 3:
        return make_scenario(25e6, 16, 60.0 if duration is None else duration, PRNS_12, 1003, 250.0)
    if cfg == 4:
        return make_scenario(50e6, 8, 0.020 if duration is None else duration, PRNS_8, 1004, 50.0)
    if cfg == 5:
        return make_scenario(25e6, 16, 10.0 if duration is None else duration, PRNS_12,
                             1005 + recording, 250.0)
    raise ValueError(f"unknown BASELINE config {cfg}")


def generate_iq(sc: Scenario, chunk: int = 1 << 20) -> np.ndarray:
    """Interleaved I,Q integer samples, shape (2*n_samples,), dtype int8/int16."""
    rng = np.random.default_rng(sc.seed + 7919)
    n = sc.n_samples
    out = np.empty(2 * n, dtype=sc.dtype)
    lim = 127 if sc.nbits == 8 else 32767
    nbits_nav = int(sc.duration * 50) + 3
    codes = {s.prn: ca_code_pm1(s.prn).astype(np.float64) for s in sc.sats}
    nav = {s.prn: rng.choice(np.array([-1.0, 1.0]), nbits_nav) for s in sc.sats}
    amp = {s.prn: sc.sigma * np.sqrt(2.0 * 10.0 ** (s.cn0 / 10.0) / sc.fs) for s in sc.sats}
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        t = np.arange(lo, hi, dtype=np.float64) / sc.fs
        x = sc.sigma * (rng.standard_normal(hi - lo) + 1j * rng.standard_normal(hi - lo))
        for s in sc.sats:
            fcode = CODE_FREQ * (1.0 + s.doppler / L1_FREQ)
            ph = fcode * t - s.delay_chips
            chip = np.floor(ph).astype(np.int64)
            c = codes[s.prn][np.mod(chip, CODE_CHIPS)]
            d = nav[s.prn][np.floor_divide(chip, 20 * CODE_CHIPS) + 1]
            car = np.exp(1j * (2 * np.pi * (sc.inter_freq + s.doppler) * t + s.phase))
            x += amp[s.prn] * c * d * car
        out[2 * lo:2 * hi:2] = np.clip(np.round(x.real), -lim, lim).astype(sc.dtype)
        out[2 * lo + 1:2 * hi:2] = np.clip(np.round(x.imag), -lim, lim).astype(sc.dtype)
    return out


def _iq_piece(a):
    sc, lo, hi, nav = a
    rng = np.random.default_rng([sc.seed + 7919, lo])
    lim = 127 if sc.nbits == 8 else 32767
    t = np.arange(lo, hi, dtype=np.float64) / sc.fs
    x = sc.sigma * (rng.standard_normal(hi - lo) + 1j * rng.standard_normal(hi - lo))
    for s in sc.sats:
        fcode = CODE_FREQ * (1.0 + s.doppler / L1_FREQ)
        chip = np.floor(fcode * t - s.delay_chips).astype(np.int64)
        c = ca_code_pm1(s.prn).astype(np.float64)[np.mod(chip, CODE_CHIPS)]
        d = nav[s.prn][np.floor_divide(chip, 20 * CODE_CHIPS) + 1]
        amp = sc.sigma * np.sqrt(2.0 * 10.0 ** (s.cn0 / 10.0) / sc.fs)
        x += amp * c * d * np.exp(1j * (2 * np.pi * (sc.inter_freq + s.doppler) * t + s.phase))
    out = np.empty(2 * (hi - lo), dtype=sc.dtype)
    out[0::2] = np.clip(np.round(x.real), -lim, lim).astype(sc.dtype)
    out[1::2] = np.clip(np.round(x.imag), -lim, lim).astype(sc.dtype)
    return out


def generate_iq_parallel(sc: Scenario, pool, chunk: int = 1 << 20) -> np.ndarray:
    """generate_iq's signal model with the 1 Mi-sample pieces spread over a multiprocessing pool (bench set-up on
    the host: seconds of 25 MS/s signal).  Same satellites, code phases, data bits and statistics as
    generate_iq; the noise comes from one generator per piece, so the bytes differ."""
    n = sc.n_samples
    nav = nav_bits_of(sc)
    pieces = pool.map(_iq_piece, [(sc, lo, min(n, lo + chunk), nav) for lo in range(0, n, chunk)], chunksize=1)
    return np.concatenate(pieces)


def nav_bits_of(sc: Scenario) -> dict:
    """The +-1 data bits generate_iq modulates on every satellite (same generator, same draws)."""
    rng = np.random.default_rng(sc.seed + 7919)
    nbits_nav = int(sc.duration * 50) + 3
    return {s.prn: rng.choice(np.array([-1.0, 1.0]), nbits_nav) for s in sc.sats}


def nav_bit_index(sat, t):
    """Index into nav_bits_of(...)[prn] of the data bit on the air at receiver time t (seconds)."""
    fcode = CODE_FREQ * (1.0 + sat.doppler / L1_FREQ)
    chip = np.floor(fcode * np.asarray(t, dtype=np.float64) - sat.delay_chips).astype(np.int64)
    return np.floor_divide(chip, 20 * CODE_CHIPS) + 1


def to_complex(iq: np.ndarray) -> np.ndarray:
    """What RFSignal.readFile returns: I + 1j*Q as complex128 (rfsignal.py:127-130)."""
    return iq[0::2] + 1j * iq[1::2]


def write_file(path: str, iq: np.ndarray) -> None:
    iq.tofile(path)


def generate_iq_torch(sc: Scenario, device="cuda", chunk: int = 1 << 22):
    """Same signal model evaluated with torch on `device` (bench set-up only: fast enough for
    seconds of 25 MS/s signal; statistically, not bit-wise, identical to generate_iq).
    Returns an interleaved int8/int16 torch tensor on `device`."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(sc.seed + 7919)
    n = sc.n_samples
    tdt = torch.int8 if sc.nbits == 8 else torch.int16
    out = torch.empty(2 * n, dtype=tdt, device=device)
    lim = 127 if sc.nbits == 8 else 32767
    rng = np.random.default_rng(sc.seed + 7919)
    nbits_nav = int(sc.duration * 50) + 3
    codes = {s.prn: torch.from_numpy(ca_code_pm1(s.prn).astype(np.float64)).to(device) for s in sc.sats}
    nav = {s.prn: torch.from_numpy(rng.choice(np.array([-1.0, 1.0]), nbits_nav)).to(device) for s in sc.sats}
    amp = {s.prn: sc.sigma * np.sqrt(2.0 * 10.0 ** (s.cn0 / 10.0) / sc.fs) for s in sc.sats}
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        t = torch.arange(lo, hi, dtype=torch.float64, device=device) / sc.fs
        xr = sc.sigma * torch.randn(hi - lo, dtype=torch.float64, device=device, generator=gen)
        xi = sc.sigma * torch.randn(hi - lo, dtype=torch.float64, device=device, generator=gen)
        for s in sc.sats:
            fcode = CODE_FREQ * (1.0 + s.doppler / L1_FREQ)
            chip = torch.floor(fcode * t - s.delay_chips).to(torch.int64)
            c = codes[s.prn][torch.remainder(chip, CODE_CHIPS)]
            d = nav[s.prn][torch.div(chip, 20 * CODE_CHIPS, rounding_mode="floor") + 1]
            ph = 2 * np.pi * torch.remainder((sc.inter_freq + s.doppler) * t, 1.0) + s.phase
            a = amp[s.prn] * c * d
            xr += a * torch.cos(ph)
            xi += a * torch.sin(ph)
        out[2 * lo:2 * hi:2] = torch.clamp(torch.round(xr), -lim, lim).to(tdt)
        out[2 * lo + 1:2 * hi:2] = torch.clamp(torch.round(xi), -lim, lim).to(tdt)
    return out
