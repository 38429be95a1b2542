"""Multi-GPU sharding of the two hot paths: one process per GPU, no data
exchange inside either kernel.

* Acquisition splits (PRN, Doppler-bin) cells.  With at least as many PRNs as ranks every rank
  searches a contiguous block of PRNs over all Doppler rows and the ranks all-gather their
  24-byte peak records (`sydr_acq_peak`, 768 B for 32 PRNs).  With fewer PRNs than ranks the
  Doppler rows are split instead: every rank searches all PRNs over its rows, the 16-byte row
  summaries (`sydr_acq_row`) are all-gathered and every rank reduces them with the same
  deterministic rule as one GPU (lowest bin, then lowest code index wins ties, the order of
  np.argmax in sydr/dsp/acquisition.py:98).
* Tracking splits recordings (all channels of a recording stay on one GPU so its IQ is read
  once); there is no collective, the host gathers per-epoch records.

The reference has no counterpart: its parallelism is one OS process per channel
(sydr/channel/channel.py:21, channelManager.py:117-118).  The collectives go through
torch.distributed, NCCL over NVLink on the GPUs and gloo in the CPU tests; the functions here
only move bytes; the one computation, picking the winning row among <= 201 gathered 16-byte row
summaries per PRN in mode "bins", is sydr_acq_reduce_rows of the C ABI (the same code the
single-GPU path ends with).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L

PEAK_BYTES = L.ACQ_PEAK_DTYPE.itemsize      # 24
ROW_BYTES = L.ACQ_ROW_DTYPE.itemsize        # 16


def partition(n_units: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced [lo, hi) blocks of n_units over `world` ranks (the first
    n_units % world ranks get one more).  Ranks beyond n_units get empty blocks."""
    if world <= 0:
        raise ValueError("world must be positive")
    base, extra = divmod(max(0, int(n_units)), world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


@dataclass(frozen=True)
class AcqShard:
    mode: str            # "prn": a block of PRNs over all rows; "bins": all PRNs over a block of rows
    prns: tuple          # PRNs this rank searches
    bin_lo: int          # first Doppler row (inclusive)
    bin_hi: int          # last Doppler row (exclusive)
    prn_lo: int          # index of prns[0] in the global PRN list (mode "prn")
    n_prn_total: int
    n_bins_total: int


def n_doppler_bins(doppler_range: float, doppler_step: float) -> int:
    """len(np.arange(-R, R + 1, step)), sydr/dsp/acquisition.py:34."""
    return len(np.arange(-doppler_range, doppler_range + 1, doppler_step))


def plan_acquisition(prns, n_bins: int, rank: int, world: int) -> AcqShard:
    """The cells rank `rank` of `world` searches."""
    prns = tuple(int(p) for p in prns)
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    if len(prns) >= world:
        lo, hi = partition(len(prns), world)[rank]
        return AcqShard("prn", prns[lo:hi], 0, n_bins, lo, len(prns), n_bins)
    lo, hi = partition(n_bins, world)[rank]
    return AcqShard("bins", prns, lo, hi, 0, len(prns), n_bins)


def plan_recordings(n_recordings: int, rank: int, world: int) -> range:
    """Recordings (with all their channels) tracked by this rank."""
    lo, hi = partition(n_recordings, world)[rank]
    return range(lo, hi)


def _all_gather_bytes(local: torch.Tensor, counts: list[int], group=None) -> list[torch.Tensor]:
    """All-gather of uint8 tensors whose sizes (`counts`, bytes per rank) may differ: every rank
    pads to the largest, one all_gather_into_tensor, then the padding is cut off."""
    world = len(counts)
    width = max(max(counts), 1)
    send = torch.zeros(width, dtype=torch.uint8, device=local.device)
    send[:local.numel()] = local.reshape(-1)
    recv = torch.empty(world * width, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return [recv[r * width:r * width + counts[r]] for r in range(world)]


def gather_peak_table(local, shard: AcqShard, world: int, group=None, reduce_rows=None) -> np.ndarray:
    """Global peak table (one `sydr_acq_peak` per PRN, in the order of the global PRN list) from
    this rank's part; identical on every rank and for every world size.

    mode "prn":  `local` = this rank's peak records as a uint8 tensor (device for NCCL, CPU
                 for gloo), 24 B per local PRN.
    mode "bins": `local` = this rank's row summaries, uint8, laid out [n_prn][bin_hi-bin_lo][16 B].
                 `reduce_rows(rows[n_prn, n_bins], prns) -> peaks` replaces sydr_acq_reduce_rows
                 (tests pass a checker here).
    """
    if world == 1 and shard.mode == "prn":
        return local.cpu().numpy().view(L.ACQ_PEAK_DTYPE).copy()
    if shard.mode == "prn":
        counts = [(hi - lo) * PEAK_BYTES for lo, hi in partition(shard.n_prn_total, world)]
        parts = _all_gather_bytes(local, counts, group) if world > 1 else [local]
        return torch.cat(parts).cpu().numpy().view(L.ACQ_PEAK_DTYPE).copy()
    blocks = partition(shard.n_bins_total, world)
    counts = [shard.n_prn_total * (hi - lo) * ROW_BYTES for lo, hi in blocks]
    parts = _all_gather_bytes(local, counts, group) if world > 1 else [local]
    rows = np.concatenate(
        [p.cpu().numpy().view(L.ACQ_ROW_DTYPE).reshape(shard.n_prn_total, hi - lo) for p, (lo, hi) in zip(parts, blocks)],
        axis=1)
    rows = np.ascontiguousarray(rows)
    prns = np.asarray(shard.prns, dtype=np.int32)
    if reduce_rows is not None:
        return reduce_rows(rows, prns)
    peaks = np.zeros(shard.n_prn_total, dtype=L.ACQ_PEAK_DTYPE)
    L.check(L.load().sydr_acq_reduce_rows(rows.ctypes.data, prns.ctypes.data, shard.n_prn_total, shard.n_bins_total,
                                          peaks.ctypes.data), "sydr_acq_reduce_rows")
    return peaks


class ShardedAcquisition:
    """AcquisitionEngine of this rank's cells + the gather; `run()` returns the same table on
    every rank as a single-GPU AcquisitionEngine.run() over all PRNs."""

    def __init__(self, fs, inter_freq, doppler_range, doppler_step, coh, noncoh, prns, rank=None, world=None,
                 group=None, device=None):
        from .engine import AcquisitionEngine
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.group = group
        self.shard = plan_acquisition(prns, n_doppler_bins(doppler_range, doppler_step), self.rank, self.world)
        s = self.shard
        self.engine = None
        if len(s.prns) and s.bin_hi > s.bin_lo:
            self.engine = AcquisitionEngine(fs, inter_freq, doppler_range, doppler_step, coh, noncoh, list(s.prns),
                                            s.bin_lo, s.bin_hi, device=device)

    def run(self, iq_dev: torch.Tensor, stream=None) -> np.ndarray:
        if self.engine is None:
            local = torch.empty(0, dtype=torch.uint8, device=iq_dev.device)
        else:
            self.engine.launch(iq_dev, stream=stream)
            local = self.engine._peaks if self.shard.mode == "prn" else self.engine._rows
        return gather_peak_table(local, self.shard, self.world, self.group)

    def close(self):
        if self.engine is not None:
            self.engine.close()
