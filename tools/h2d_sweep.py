"""Raw pinned host -> device copy bandwidth with N ranks copying at the same time (one process per GPU), to name the
limiter of the end-to-end path on several GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_sweep.py [GB per copy]
Variants of the host buffer: torch pinned (cudaHostAlloc default), write-combined pinned (cudaHostAllocWriteCombined),
pinned memory allocated and first touched on the CPUs of the GPU's own NUMA node (sched_setaffinity from the device's
local_cpulist), and several copies in flight on several streams.  Device-timed (CUDA events), barrier on both sides,
per-rank and summed GB/s; rank 0 prints one line per variant."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = int(gb * (1 << 30))
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
rt = C.CDLL("libcudart.so")


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(copy, reps=6):
    copy(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        copy()
    e1.record(); torch.cuda.synchronize()
    gbs = reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    lo, sm = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    barrier()
    return float(lo[0]), float(sm[0])


def host_alloc(flags):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc flags {flags}: error {rc}")
    C.memset(p, 1, nbytes)                     # first touch
    return p


def raw_copy(p):
    s = torch.cuda.current_stream().cuda_stream
    return lambda: rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(nbytes), C.c_int(1), C.c_void_p(s))


def numa_cpus():
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dv = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dv:02x}.0/local_cpulist"
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus & os.sched_getaffinity(0), txt
    except Exception as exc:
        return set(), f"unavailable ({type(exc).__name__}: {exc})"


rows = []
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h.fill_(1)
rows.append(("torch pin_memory", *timed(lambda: d.copy_(h, non_blocking=True))))
streams = [torch.cuda.Stream() for _ in range(4)]
q = nbytes // 4


def four():
    for i, s in enumerate(streams):
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            d[i * q:(i + 1) * q].copy_(h[i * q:(i + 1) * q], non_blocking=True)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)


rows.append(("torch pin_memory, 4 pieces on 4 streams", *timed(four)))
del h
p = host_alloc(0)
rows.append(("cudaHostAlloc default", *timed(raw_copy(p))))
rt.cudaFreeHost(p)
p = host_alloc(4)
rows.append(("cudaHostAlloc write-combined", *timed(raw_copy(p))))
rt.cudaFreeHost(p)
cpus, where = numa_cpus()
if cpus:
    old = os.sched_getaffinity(0)
    os.sched_setaffinity(0, cpus)
    p = host_alloc(0)
    rows.append((f"cudaHostAlloc, allocated + first touched on the GPU's NUMA node (cpus {where})", *timed(raw_copy(p))))
    rt.cudaFreeHost(p)
    os.sched_setaffinity(0, old)
else:
    rows.append((f"NUMA-local variant skipped: {where}", float("nan"), float("nan")))
if rank == 0:
    print(f"== {world} rank(s) copying {gb:g} GiB each at the same time; host cpus available {len(os.sched_getaffinity(0))}")
    for name, lo, sm in rows:
        print(f"  {name:90s} slowest rank {lo:6.1f} GB/s   sum over ranks {sm:7.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
