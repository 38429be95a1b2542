"""Multi-GPU host logic on CPU: world_size 2 and 3 over gloo.  The cells a rank
owns, the byte-level all-gather of peak records / row summaries and the deterministic row
reduction must reproduce the single-rank table bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sydr_b200 import _lib as L
from sydr_b200 import shard as S


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 32, 41, 201):
        for w in (1, 2, 3, 4, 8):
            blocks = S.partition(n, w)
            assert len(blocks) == w and blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_plans():
    prns = list(range(1, 33))
    assert S.n_doppler_bins(5000.0, 250.0) == 41 and S.n_doppler_bins(5000.0, 50.0) == 201
    got = [S.plan_acquisition(prns, 41, r, 8) for r in range(8)]
    assert all(g.mode == "prn" and len(g.prns) == 4 and (g.bin_lo, g.bin_hi) == (0, 41) for g in got)
    assert sum((list(g.prns) for g in got), []) == prns
    few = [S.plan_acquisition([5, 9], 41, r, 4) for r in range(4)]
    assert all(g.mode == "bins" and g.prns == (5, 9) for g in few)
    assert [(g.bin_lo, g.bin_hi) for g in few] == S.partition(41, 4)
    assert [list(S.plan_recordings(32, r, 8)) for r in range(8)] == [list(range(4 * r, 4 * r + 4)) for r in range(8)]
    with pytest.raises(ValueError):
        S.plan_acquisition(prns, 41, 8, 8)


def _reference_reduce(rows, prns):
    """np.argmax order (acquisition.py:98): first maximum over rows wins."""
    peaks = np.zeros(len(prns), dtype=L.ACQ_PEAK_DTYPE)
    for p in range(len(prns)):
        best = int(np.argmax(rows["peak1"][p]))
        r = rows[p, best]
        peaks[p] = (prns[p], best, r["code_idx"], r["peak1"], r["peak2"], np.float32(r["peak1"]) / np.float32(r["peak2"]))
    return peaks


def _tables(n_prn, n_bins, seed=7):
    rng = np.random.default_rng(seed)
    rows = np.zeros((n_prn, n_bins), dtype=L.ACQ_ROW_DTYPE)
    rows["peak1"] = rng.uniform(10, 20, (n_prn, n_bins)).astype(np.float32)
    rows["peak2"] = rng.uniform(5, 9, (n_prn, n_bins)).astype(np.float32)
    rows["code_idx"] = rng.integers(0, 10000, (n_prn, n_bins))
    rows["peak1"][0, [3, n_bins - 2]] = 50.0          # a tie: the lower bin must win on every world size
    rows["peak1"][-1, n_bins - 1] = 60.0              # winner in the last rank's block
    return rows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = {}
    try:
        # (a) PRN blocks: all-gather of 24-byte peak records
        prns = list(range(1, 33))
        full = _reference_reduce(_tables(32, 41), prns)
        sh = S.plan_acquisition(prns, 41, rank, world)
        local = torch.from_numpy(full[sh.prn_lo:sh.prn_lo + len(sh.prns)].view(np.uint8).copy())
        ok["prn"] = S.gather_peak_table(local, sh, world).tobytes() == full.tobytes()
        # (b) fewer PRNs than ranks: Doppler-row blocks, all-gather of row summaries + reduction
        prns = [4] if world == 2 else [4, 17]
        rows = _tables(len(prns), 41)
        full = _reference_reduce(rows, prns)
        sh = S.plan_acquisition(prns, 41, rank, world)
        assert sh.mode == "bins"
        local = torch.from_numpy(np.ascontiguousarray(rows[:, sh.bin_lo:sh.bin_hi]).view(np.uint8).reshape(-1).copy())
        got_lib = S.gather_peak_table(local, sh, world)                          # C ABI reduction
        got_chk = S.gather_peak_table(local, sh, world, reduce_rows=_reference_reduce)
        ok["bins"] = got_lib.tobytes() == full.tobytes() and got_chk.tobytes() == full.tobytes()
        # (c) tracking: recordings per rank, records gathered by the host (no collective)
        mine = list(S.plan_recordings(5, rank, world))
        counts = [None] * world
        dist.all_gather_object(counts, mine)
        ok["rec"] = sorted(sum(counts, [])) == list(range(5))
    except Exception as e:                                  # noqa: BLE001
        ok["error"] = repr(e)
    finally:
        q.put((rank, ok))
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_gather_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world))
    for _, ok in res:
        assert ok == {"prn": True, "bins": True, "rec": True}, ok
