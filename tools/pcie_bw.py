"""Raw pinned-memory device-to-host bandwidth of the box: the ceiling of bench.py's `e2e` figure, which has to
move 33.2 MB of decoded picture per 2160p picture.

    python tools/pcie_bw.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_bw.py

One process per GPU (as bench.py runs): every rank copies its own 2 GiB device buffer into its own pinned host
buffer, all ranks at the same time; the aggregate is total bytes over the slowest rank's device-timed interval.
Three variants per run: D2H alone, D2H next to a host-to-device stream, and D2H in 32 MiB pieces on two streams
(the shape of SequenceDecoder's copies).  One cudaMemcpyAsync per piece -- no batched-copy API.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    from vc2_conformance_b200 import sharding

    numa_cpus = sharding.bind_process_to_gpu(local)
except Exception:  # the tool must run even where the binding helper cannot
    numa_cpus = 0

n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
REPS = 4
PIECE = 32 << 20


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()


results = {}
for name in ("d2h only", "d2h + h2d", "d2h 32 MiB pieces on two streams"):
    for _ in range(2):
        h.copy_(d, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    for s in (s1, s2, s3):
        s.wait_stream(cur)
    for _ in range(REPS):
        if name.startswith("d2h 32"):
            for k, off in enumerate(range(0, n, PIECE)):
                with torch.cuda.stream(s1 if k % 2 == 0 else s3):
                    h[off:off + PIECE].copy_(d[off:off + PIECE], non_blocking=True)
        else:
            with torch.cuda.stream(s1):
                h.copy_(d, non_blocking=True)
        if name == "d2h + h2d":
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
    for s in (s1, s2, s3):
        cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    results[name] = {"per_gpu_GBps": REPS * n / e0.elapsed_time(e1) / 1e6,
                     "aggregate_GBps": world * REPS * n / float(ms.item()) / 1e6}
    barrier()
if rank == 0:
    print(json.dumps({"tool": "pcie_bw", "n_gpus": world, "bytes_per_copy": n, "numa_bound_cpus": numa_cpus,
                      "host_cpus": os.cpu_count(), "d2h": results}))
if dist is not None:
    dist.destroy_process_group()
