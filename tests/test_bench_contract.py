"""bench.py's contract pieces that run without a GPU: the reference arm's JSON line (tier contract: same metric / unit /
config keys as the GPU arm, `impl`, `cpu_baseline` with kind / cores / sample, an `e2e` object with zero copy bytes), the
tracked DRAM-traffic file the GPU arm reads, and the legacy prototype table against the header."""
import json
import numpy as np
import os
import re
import subprocess
import sys

import helpers as H  # noqa: F401


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--chunk-seconds", "0.12", "--ref-budget-s", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["metric"] == "cold acquisition (32 PRN) + 12-channel tracking throughput"
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "PRNs acquired" in cb["sample"]
    assert cb["trk"]["channels"] == 12 and cb["acq"]["prns"] == 32
    # a short chunk fits the budget whole: nothing is extrapolated, and the clock is the one that ran
    assert line["sampled"] is False and abs(line["ms_per_step"] * 1e-3 - line["timed_s"]) < 0.05 * line["timed_s"] + 0.05
    assert "workload" in line["config"] and "model" not in line["config"]


def test_tracked_traffic_file():
    sys.path.insert(0, H.ROOT)
    import bench
    total, info = bench.ncu_traffic(60.0, 24)
    assert info["file"] == "profiles/ncu_traffic.json" and info["measured_in_this_run"] is False
    assert all(os.path.exists(os.path.join(H.ROOT, src)) for src in info["sources"])
    # the tracking launch of a step reads the samples of its 24 recordings once: 24 x 4 B x 25 MS/s x 60 s, plus the records
    assert 0.98 * 24 * 6.0e9 < info["trk_borre_kernel"] < 1.1 * 24 * 6.0e9 and total > info["trk_borre_kernel"]


def test_batch_slots_hold_different_samples():
    """bench.fill_batch: the slots beyond the generated recordings are those times j, -1, -j (CPU stand-in for the batch)."""
    import types
    import torch
    sys.path.insert(0, H.ROOT)
    import bench
    n = 64
    store = torch.zeros(4 * 2 * n, dtype=torch.int16)

    class FakeBatch:
        B = 4

        def slot(self, r):
            return store[r * 2 * n:(r + 1) * 2 * n]

    real_sync = torch.cuda.synchronize
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        host = torch.randint(-3000, 3000, (2 * n,), dtype=torch.int16)
        scs = bench.fill_batch(FakeBatch(), "sc0", host, 0, types.SimpleNamespace(seeds=1, chunk_seconds=0.0), "cpu")
    finally:
        torch.cuda.synchronize = real_sync
    assert scs == ["sc0"] * 4
    x = host.numpy().astype(np.float64).view(np.complex128)
    for r, f in enumerate((1, 1j, -1, -1j)):
        got = store[r * 2 * n:(r + 1) * 2 * n].numpy().astype(np.float64).view(np.complex128)
        assert np.array_equal(got, f * x), r


def test_legacy_prototypes_cover_the_header():
    from sydr_b200.old._legacy import PROTOTYPES
    hdr = open(os.path.join(H.ROOT, "include", "sydr_b200.h")).read()
    legacy = re.findall(r"^void\s+([A-Za-z]+)\(", hdr, re.M)
    legacy = [n for n in legacy if not n.startswith("sydr_")]
    assert sorted(legacy) == sorted(PROTOTYPES) and len(PROTOTYPES) == 9


def test_reference_checker_is_built_when_the_reference_is_mounted():
    """oracle/_ref/tracking.so (the reference's own tracking.c, compiled by oracle/Makefile) is what tests/test_gpu_legacy.py checks
    the legacy entry points against; it is built in the container that has /root/reference and travels with the repository."""
    if os.path.isdir("/root/reference/sydr/c_functions"):
        assert os.path.exists(os.path.join(H.ROOT, "oracle", "_ref", "tracking.so")), "oracle/Makefile did not produce oracle/_ref/tracking.so"
