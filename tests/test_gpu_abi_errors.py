"""Argument validation at the C-ABI boundary (include/sydr_b200.h): a bad call returns a negative
status, leaves a message in sydr_last_error() and launches nothing; the Python mirror raises
SydrError with that message.  The reference raises from numpy / its own checks in the same
situations (a PRN outside its tap table in GenerateGPSGoldCode, gnsstools.py; too few samples for
PCPS's reshape, acquisition.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from sydr_b200 import _lib as L

pytestmark = pytest.mark.gpu

ERR_ARG, ERR_UNSUPPORTED = -2, -3


def raises_with(code, fragment):
    class Ctx:
        def __enter__(self):
            return self

        def __exit__(self, et, ev, tb):
            assert et is L.SydrError, f"expected SydrError, got {et}"
            assert f"(code {code})" in str(ev) and fragment in str(ev), str(ev)
            return True
    return Ctx()


def test_code_generation_rejects_bad_prns():
    lib = L.load()
    out = np.zeros(1023)
    for prn in (0, -1, 38, 1 << 20):          # the table holds PRN 1..37 like the reference's G2 tap list
        assert lib.sydr_ca_code(prn, out.ctypes.data) == ERR_ARG
        assert "out of range" in L.last_error()
    assert lib.sydr_ca_code(1, None) == ERR_ARG
    assert lib.sydr_ca_code(32, out.ctypes.data) == 0 and set(np.unique(out)) == {-1.0, 1.0}


def test_acquisition_plan_validation():
    from sydr_b200.engine import AcquisitionEngine
    with raises_with(ERR_ARG, "out of range"):
        AcquisitionEngine(4e6, 0.0, 5000, 250, 1, 1, [1, 40])
    with raises_with(ERR_ARG, "doppler"):
        AcquisitionEngine(4e6, 0.0, 5000, 0, 1, 1, [1])
    with raises_with(ERR_ARG, ">= 1"):
        AcquisitionEngine(4e6, 0.0, 5000, 250, 0, 1, [1])
    with raises_with(ERR_UNSUPPORTED, ""):
        AcquisitionEngine(4.001e6, 0.0, 5000, 250, 1, 1, [1])       # 4001 samples per code: 4001 is prime
    eng = AcquisitionEngine(4e6, 0.0, 5000, 250, 1, 2, [3, 7])
    short = torch.zeros(2 * 4000, dtype=torch.int8, device="cuda")   # noncoh=2 needs 8000 samples
    with raises_with(ERR_ARG, "needs 8000 samples"):
        eng.run(short)
    ok = torch.zeros(2 * 8000, dtype=torch.int8, device="cuda")
    assert len(eng.run(ok)["peaks"]) == 2                             # the plan is still usable after the error
    eng.close()


def test_tracking_launch_validation():
    from sydr_b200.engine import TrackingEngine, make_trk_states
    fs = 4e6
    chans = [dict(prn=5, carrier_freq=1000.0, start_sample=0, iq_len=40000)]
    iq = torch.randint(-20, 21, (2 * 40000,), dtype=torch.int8, device="cuda",
                       generator=torch.Generator("cuda").manual_seed(5))
    with raises_with(ERR_ARG, "threads"):
        TrackingEngine(fs, make_trk_states(fs, chans), 8, threads=100).run(iq)
    with raises_with(ERR_ARG, "cluster"):
        TrackingEngine(fs, make_trk_states(fs, chans), 8, cluster=3).run(iq)
    with raises_with(ERR_UNSUPPORTED, "below the supported"):
        TrackingEngine(1e6, make_trk_states(1e6, chans), 8).run(iq)
    # a state written with a PRN that has no code: the channel is aborted on the device (status
    # SYDR_ERR_STATE, no epochs, no table read), its neighbour in the same launch is unaffected
    both = make_trk_states(fs, [dict(prn=0, carrier_freq=0.0, start_sample=0, iq_len=40000), chans[0]])
    eng = TrackingEngine(fs, both, 8)
    eng.launch(iq)
    res = eng.fetch()
    assert len(res[0]) == 0 and len(res[1]) == 8
    assert list(eng.states()["status"]) == [-4, 0]
    with raises_with(-4, "aborted on channels [0]"):
        TrackingEngine(fs, both, 8).run(iq)                 # run() turns an aborted channel into an exception
    # a well-formed launch on the same (noise-only) input runs after the refused ones
    res = TrackingEngine(fs, make_trk_states(fs, chans), 8).run(iq)
    assert len(res) == 1 and len(res[0]) == 8


def test_nav_and_handoff_validation():
    lib = L.load()
    d = torch.zeros(64, dtype=torch.float64, device="cuda")
    assert lib.sydr_nav_state_init(None) == ERR_ARG
    rc = lib.sydr_acq_handoff(d.data_ptr(), 0, 0.0, 5000.0, 250.0, C.c_int64(4000), C.c_int64(4000), C.c_int64(0), 1.5,
                              d.data_ptr(), C.c_int64(40000), d.data_ptr(), 4, None, None)
    assert rc == ERR_ARG and L.last_error()
    rc = lib.sydr_acq_handoff(d.data_ptr(), 4, 0.0, 5000.0, 250.0, C.c_int64(4000), C.c_int64(4000), C.c_int64(0), 1.5,
                              d.data_ptr(), C.c_int64(40000), d.data_ptr(), 65, None, None)
    assert rc == ERR_ARG


def test_epl_batch_refuses_calls_outside_the_recording():
    from sydr_b200.engine import epl_batch
    fs = 4e6
    iq = torch.randint(-20, 21, (2 * 8000,), dtype=torch.int8, device="cuda", generator=torch.Generator("cuda").manual_seed(6))
    args = np.zeros(5, dtype=L.EPL_ARGS_DTYPE)
    args["prn"], args["start"], args["n"] = [4, 4, 4, 0, 38], [0, 4001, -8, 0, 0], [4000, 4000, 4000, 4000, 4000]
    args["n"][2] = 0
    args["start"][2] = 0
    args["carrier_freq"], args["code_step"], args["spacing"] = 500.0, 1.023e6 / fs, (-0.5, 0.0, 0.5)
    out = epl_batch(iq, fs, args)
    assert np.isfinite(out[0]).all() and np.abs(out[0]).max() > 0
    assert np.isnan(out[1:]).all()            # window past the end, n = 0, PRN 0, PRN 38


def test_tracking_stops_at_the_allocation_when_iq_len_overstates_it():
    """A state whose iq_len claims more samples than the buffer holds is tracked to the end of the buffer only."""
    from sydr_b200.engine import TrackingEngine, make_trk_states
    fs = 4e6
    iq = torch.randint(-20, 21, (2 * 40000,), dtype=torch.int8, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    good = TrackingEngine(fs, make_trk_states(fs, [dict(prn=9, carrier_freq=0.0, start_sample=0, iq_len=40000)]), 64).run(iq)
    over = TrackingEngine(fs, make_trk_states(fs, [dict(prn=9, carrier_freq=0.0, start_sample=0, iq_len=10 ** 9)]), 64).run(iq)
    assert 9 <= len(over[0]) <= 10 and over[0].tobytes() == good[0][:len(over[0])].tobytes()
