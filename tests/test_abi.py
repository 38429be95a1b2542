"""The C-ABI library loads on a CPU-only box, exports every symbol include/sydr_b200.h declares,
and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import os
import re

import numpy as np
import pytest

import helpers as H
from sydr_b200 import _lib as L


def header_functions():
    src = open(os.path.join(H.ROOT, "include", "sydr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(?:int|void|long long|const char\*)\s+\*?([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = header_functions()
    assert len(names) >= 29
    for n in names:
        assert hasattr(lib, n), f"libsydr_b200.so does not export {n}"
        assert n in L.SIGNATURES, f"_lib.py has no signature for {n}"
    assert lib.sydr_abi_version() == 1


def test_struct_layouts_match_header():
    assert L.ACQ_PEAK_DTYPE.itemsize == 24 and L.ACQ_ROW_DTYPE.itemsize == 16
    assert L.EPL_ARGS_DTYPE.itemsize == 72 and L.TRK_STATE_DTYPE.itemsize == 192 and L.TRK_EPOCH_DTYPE.itemsize == 128
    assert L.TRK_STATE_DTYPE.fields["carrier_freq"][1] == 48 and L.TRK_STATE_DTYPE.fields["spacing"][1] == 168


def test_no_cpu_fallback():
    lib = L.load()
    if lib.sydr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    out = np.zeros(1023)
    assert lib.sydr_ca_code(1, out.ctypes.data) != 0 and L.last_error()
    with pytest.raises(L.SydrError):
        L.require_device()
    from sydr_b200.dsp.tracking import EPL
    with pytest.raises(L.SydrError):
        EPL(np.zeros(4000, dtype=np.complex128), np.ones(1025), 4e6, 0.0, 0.0, 0.0, 0.25575, [-0.5, 0.0, 0.5])
    from sydr_b200.dsp.acquisition import PCPS
    with pytest.raises(L.SydrError):
        PCPS(np.zeros(4000, dtype=np.complex128), 0.0, 4e6, np.zeros(4000, dtype=complex), 5000, 250, 4000)


def test_receiver_needs_the_gpu(tmp_path):
    """main.py's receiver has no CPU path either: building it without a CUDA device fails loudly."""
    import configparser
    lib = L.load()
    if lib.sydr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from sydr_b200.receiver.receiver_gps_l1ca import ReceiverGPSL1CA
    cfg = configparser.ConfigParser()
    assert cfg.read(os.path.join(H.ROOT, "config", "receiver.ini"))
    cfg["DEFAULT"]["outfolder"] = str(tmp_path)
    with pytest.raises(L.SydrError):
        ReceiverGPSL1CA(cfg, overwrite=True)
    for name in ("channel_GPS_L1CA_borre.ini", "channel_GPS_L1CA_kaplan.ini"):
        ch = configparser.ConfigParser()
        assert ch.read(os.path.join(H.ROOT, "config", "channels", name)) and "TRACKING" in ch and "ACQUISITION" in ch


def test_product_never_imports_the_oracle():
    pkg = os.path.join(H.ROOT, "sydr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
