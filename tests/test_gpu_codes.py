"""K-CODE on the device against the reference's code tables and code spectra."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ca_codes(golden):
    from sydr_b200.signal.gnsssignal import GenerateGPSGoldCode
    g = golden("codes.npz")
    for prn in range(1, 33):
        c = GenerateGPSGoldCode(prn)
        assert c.dtype == np.float64 and np.array_equal(c.astype(np.int8), g["codes"][prn - 1])


@pytest.mark.parametrize("fs", [4e6, 10e6, 25e6, 50e6])
def test_code_spectrum(golden, fs):
    from sydr_b200.signal.gnsssignal import CodeSpectrum, GenerateGPSGoldCode
    g = golden("codes.npz")
    sel = g[f"sel_{int(fs)}"]
    for prn in (1, 19, 32):
        spec = CodeSpectrum(prn, fs)
        ref = g[f"spec_{int(fs)}_{prn}"]
        # tolerance of the parity protocol is 1e-5 of the largest bin (FP32 tables); FP64 K-CODE does far better
        assert np.abs(spec[sel] - ref).max() <= 1e-9 * float(g[f"specabsmax_{int(fs)}_{prn}"])
        up = GenerateGPSGoldCode(prn, fs)
        assert np.array_equal(np.array([up.sum(), (up * np.arange(len(up))).sum()]), g[f"upsum_{int(fs)}_{prn}"])


def test_code_spectrum_non_smooth_length():
    """fs = 4.092 MHz -> 4092 = 4*3*11*31 samples per code: direct-DFT path."""
    from oracle import sydr_oracle as O
    from sydr_b200.signal.gnsssignal import CodeSpectrum
    fs = 4.092e6
    spec = CodeSpectrum(7, fs)
    ref = O.code_spectrum(7, fs)
    assert np.abs(spec - ref).max() <= 1e-9 * np.abs(ref).max()
