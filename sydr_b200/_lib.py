"""ctypes binding of libsydr_b200.so (include/sydr_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is
called without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsydr_b200.so")

SYDR_OK = 0
IQ_I8, IQ_I16, IQ_F32, IQ_F64 = 0, 1, 2, 3


class SydrError(RuntimeError):
    pass


class AcqPeak(C.Structure):
    _fields_ = [("prn", C.c_int32), ("freq_idx", C.c_int32), ("code_idx", C.c_int32),
                ("peak1", C.c_float), ("peak2", C.c_float), ("ratio", C.c_float)]


class AcqRow(C.Structure):
    _fields_ = [("peak1", C.c_float), ("code_idx", C.c_int32), ("peak2", C.c_float), ("reserved", C.c_int32)]


class EplArgs(C.Structure):
    _fields_ = [("start", C.c_int64), ("n", C.c_int32), ("prn", C.c_int32),
                ("carrier_freq", C.c_double), ("rem_carrier", C.c_double), ("rem_code", C.c_double),
                ("code_step", C.c_double), ("spacing", C.c_double * 3)]


class TrkState(C.Structure):
    _fields_ = [("iq_base", C.c_int64), ("iq_len", C.c_int64), ("cur", C.c_int64), ("n_req", C.c_int64),
                ("epochs_done", C.c_int64), ("prn", C.c_int32), ("status", C.c_int32),
                ("carrier_freq", C.c_double), ("code_freq", C.c_double), ("code_step", C.c_double),
                ("rem_carrier", C.c_double), ("rem_code", C.c_double),
                ("nco_code", C.c_double), ("nco_code_err", C.c_double),
                ("nco_carrier", C.c_double), ("nco_carrier_err", C.c_double),
                ("dll_tau1", C.c_double), ("dll_tau2", C.c_double), ("dll_pdi", C.c_double),
                ("pll_tau1", C.c_double), ("pll_tau2", C.c_double), ("pll_pdi", C.c_double),
                ("spacing", C.c_double * 3)]


class TrkEpoch(C.Structure):
    _fields_ = [("corr", C.c_double * 6), ("dll", C.c_double), ("pll", C.c_double),
                ("carrier_freq", C.c_double), ("code_freq", C.c_double), ("code_err", C.c_double),
                ("carrier_err", C.c_double), ("start", C.c_double), ("n", C.c_double),
                ("rem_code", C.c_double), ("rem_carrier", C.c_double)]


class TrkConfig(C.Structure):
    _fields_ = [("cluster", C.c_int32), ("threads", C.c_int32), ("use_tma", C.c_int32), ("append", C.c_int32),
                ("iq_len", C.c_int64), ("min_tap_gap", C.c_double),
                ("iq_base", C.c_int64), ("use_iq_base", C.c_int32), ("dense", C.c_int32),
                ("kernel", C.c_int32), ("group", C.c_int32), ("rec_channels", C.c_int32), ("reserved", C.c_int32)]


# numpy views of the same layouts (device buffers are torch uint8 tensors reinterpreted)
ACQ_PEAK_DTYPE = np.dtype([("prn", "<i4"), ("freq_idx", "<i4"), ("code_idx", "<i4"),
                           ("peak1", "<f4"), ("peak2", "<f4"), ("ratio", "<f4")])
ACQ_ROW_DTYPE = np.dtype([("peak1", "<f4"), ("code_idx", "<i4"), ("peak2", "<f4"), ("reserved", "<i4")])
EPL_ARGS_DTYPE = np.dtype([("start", "<i8"), ("n", "<i4"), ("prn", "<i4"), ("carrier_freq", "<f8"),
                           ("rem_carrier", "<f8"), ("rem_code", "<f8"), ("code_step", "<f8"),
                           ("spacing", "<f8", (3,))])
TRK_STATE_DTYPE = np.dtype([("iq_base", "<i8"), ("iq_len", "<i8"), ("cur", "<i8"), ("n_req", "<i8"),
                            ("epochs_done", "<i8"), ("prn", "<i4"), ("status", "<i4"),
                            ("carrier_freq", "<f8"), ("code_freq", "<f8"), ("code_step", "<f8"),
                            ("rem_carrier", "<f8"), ("rem_code", "<f8"), ("nco_code", "<f8"),
                            ("nco_code_err", "<f8"), ("nco_carrier", "<f8"), ("nco_carrier_err", "<f8"),
                            ("dll_tau1", "<f8"), ("dll_tau2", "<f8"), ("dll_pdi", "<f8"),
                            ("pll_tau1", "<f8"), ("pll_tau2", "<f8"), ("pll_pdi", "<f8"),
                            ("spacing", "<f8", (3,))])
TRK_EPOCH_DTYPE = np.dtype([("corr", "<f8", (6,)), ("dll", "<f8"), ("pll", "<f8"), ("carrier_freq", "<f8"),
                            ("code_freq", "<f8"), ("code_err", "<f8"), ("carrier_err", "<f8"),
                            ("start", "<f8"), ("n", "<f8"), ("rem_code", "<f8"), ("rem_carrier", "<f8")])
KAPLAN_STATE_DTYPE = np.dtype([(n, "<f8") for n in (
    "fll_bw_pullin", "fll_bw_wide", "fll_bw_narrow", "pll_bw_wide", "pll_bw_narrow", "fll_thr_wide", "fll_thr_narrow",
    "pll_thr_narrow", "dll_threshold", "ip_prev", "qp_prev", "fll_lock", "pll_lock", "cn0", "pdpn", "vel_memory",
    "fll_bw", "pll_bw")] + [("accum_counter", "<i4"), ("lock_state", "<i4"), ("flags", "<i4"), ("reserved", "<i4"),
                           ("code_counter", "<i8")])
KAPLAN_EPOCH_DTYPE = np.dtype([("fll", "<f8"), ("cn0", "<f8"), ("fll_lock", "<f8"), ("pll_lock", "<f8"),
                               ("lock_state", "<i4"), ("flags", "<i4")])
assert KAPLAN_STATE_DTYPE.itemsize == 168 and KAPLAN_EPOCH_DTYPE.itemsize == 40
NAV_STATE_DTYPE = np.dtype([("code_counter", "<i8"), ("sync_epoch", "<i8"), ("prev_iprompt", "<f8"),
                            ("row19", "<f8"), ("nav_sum", "<f8"), ("nav_count", "<i4"), ("n_bits", "<i4")])
assert NAV_STATE_DTYPE.itemsize == 48
assert ACQ_PEAK_DTYPE.itemsize == C.sizeof(AcqPeak) == 24
assert ACQ_ROW_DTYPE.itemsize == C.sizeof(AcqRow) == 16
assert EPL_ARGS_DTYPE.itemsize == C.sizeof(EplArgs) == 72
assert TRK_STATE_DTYPE.itemsize == C.sizeof(TrkState) == 192
assert TRK_EPOCH_DTYPE.itemsize == C.sizeof(TrkEpoch) == 128

_vp, _i, _ll, _d, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_size_t
_dp, _ip, _llp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_longlong)

# name -> (restype, argtypes); every symbol include/sydr_b200.h declares
SIGNATURES = {
    "sydr_abi_version": (_i, []),
    "sydr_last_error": (C.c_char_p, []),
    "sydr_clear_error": (None, []),
    "sydr_device_count": (_i, []),
    "sydr_set_device": (_i, [_i]),
    "sydr_measure_fp32_peak": (_i, [_dp, _dp]),
    "sydr_measure_fp64_peak": (_i, [_dp, _dp, _dp]),
    "sydr_launch_count": (_ll, []),
    "sydr_reset_launch_count": (None, []),
    "sydr_ca_code": (_i, [_i, _vp]),
    "sydr_code_spectrum": (_i, [_i, _d, _vp, _ll]),
    "sydr_acq_plan_create": (_i, [_d, _d, _d, _d, _i, _i, _vp, _i, _i, _i, C.POINTER(_vp)]),
    "sydr_acq_plan_destroy": (_i, [_vp]),
    "sydr_acq_plan_info": (_i, [_vp, _ip, _ip, _ip, _ip, _llp]),
    "sydr_acq_plan_set_spectrum": (_i, [_vp, _i, _vp]),
    "sydr_acq_run": (_i, [_vp, _vp, _i, _ll, _vp, _vp, _vp, _vp]),
    "sydr_acq_reduce_rows": (_i, [_vp, _vp, _i, _i, _vp]),
    "sydr_peak_compare": (_i, [_vp, _i, _i, _i, _ip, _ip, _dp]),
    "sydr_epl_batch": (_i, [_vp, _i, _ll, _d, _vp, _i, _vp, _vp]),
    "sydr_trk_run": (_i, [_vp, _i, _ll, _d, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "sydr_trk_run_kaplan": (_i, [_vp, _i, _ll, _d, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "sydr_trk_profile_buffer": (_i, [_vp]),
    "sydr_trk_set_mode": (_i, [_i]),
    "sydr_trkm_debug": (_i, [_i]),
    "sydr_trkm_shape": (_i, [_i, _i]),
    "sydr_trk_state_init": (_i, [_vp, _i, _d, _d, _ll] + [_d] * 11),
    "sydr_convert_to_f32": (_i, [_vp, _i, _ll, _vp, _vp]),
    "sydr_acq_handoff": (_i, [_vp, _i, _d, _d, _d, _ll, _ll, _ll, _d, _vp, _ll, _vp, _i, _vp, _vp]),
    "sydr_nav_state_init": (_i, [_vp]),
    "sydr_nav_bits": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "sydr_nav_bits_kaplan": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    # legacy per-call ABI (sydr/c_functions/*.c)
    "generateReplica": (None, [_vp, _sz, _d, _d, _vp, _vp]),
    "getCorrelator": (None, [_vp, _vp, _vp, _sz, _d, _d, _d, _vp, _vp]),
    "generateCarrier": (None, [_vp, _vp, _sz, _vp, _vp]),
    "delayLockLoop": (None, [_d] * 10 + [_vp] * 3),
    "phaseLockLoop": (None, [_d] * 8 + [_vp] * 3),
    "getLoopCoefficients": (None, [_d] * 3 + [_vp] * 2),
    "setSatellite": (None, [_vp, _sz, _vp]),
    "PCPS": (None, [_vp, _vp, _ll, _ll, _ll, _d, _d, _vp, _sz, _vp]),
    "twoCorrelationPeakComparison": (None, [_vp, _sz, _vp, _sz, _ll, _ll, _d] + [_vp] * 6),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (building it is `python -m sydr_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SydrError(f"{LIB_PATH} is missing: build it with `python -m sydr_b200.build` "
                        "(there is no CPU fallback for the SyDR hot paths)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.sydr_abi_version() != 1:
        raise SydrError("libsydr_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().sydr_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != SYDR_OK:
        raise SydrError(f"{what or 'libsydr_b200'} failed (code {rc}): {last_error()}")


def require_device() -> None:
    if load().sydr_device_count() <= 0:
        raise SydrError("no CUDA device: the SyDR hot paths have no CPU fallback")


def ptr(x) -> int:
    """Device/host pointer of a torch tensor or numpy array (0 for None)."""
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x.ctypes.data
