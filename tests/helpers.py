"""Shared test helpers: regenerate the synthetic inputs the golden fixtures were made from."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from sydr_b200 import synth  # noqa: E402
import make_golden as MG  # noqa: E402  (only its pure-python case tables / generators are used)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def acq_case(name):
    """(scenario, iq, n_code, params dict) of a golden acquisition case."""
    fs, nbits, present, seed, dr, ds, coh, noncoh, search, keep = MG.ACQ_CASES[name]
    sc, iq, n = MG.acq_input(name)
    return sc, iq, n, dict(fs=fs, nbits=nbits, present=present, seed=seed, doppler_range=dr, doppler_step=ds,
                           coh=coh, noncoh=noncoh, search=search, keep=keep, chip=round(fs / 1.023e6))


def epl_input(fs, nbits, seed):
    sc = synth.make_scenario(fs, nbits, 0.0045, (3, 7), seed, 250.0)
    return synth.generate_iq(sc)


EPL_SETS = ((4e6, 8, 11), (10e6, 8, 12), (25e6, 16, 13), (50e6, 16, 14))


def loop_input(meta, prns):
    fs, nbits, seed, ms, ds = meta
    sc = synth.make_scenario(float(fs), int(nbits), int(ms) * 1e-3, tuple(int(p) for p in prns), int(seed), float(ds))
    return sc, synth.generate_iq(sc)


def abs_code_phase(start, rem_code_before, code_step):
    """Absolute code phase of an epoch: start sample minus remCode/codeStep (SURVEY hard part 3)."""
    return start - rem_code_before / code_step
