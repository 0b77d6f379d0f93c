"""
Picture-level sharding across GPUs.  Pictures of a VC-2 sequence are independent (intra-only,
/root/reference/docs/source/user_guide/limitations.rst:15-19), so ranks own disjoint pictures and
there is no collective on the data path: the only cross-rank value is the MAX of the device-timed
elapsed time (bench.py), and -- for callers that want them in one place -- a gather of results.
"""


def pictures_for_rank(num_pictures, rank, world_size, contiguous=True):
    """Indices of the pictures rank `rank` decodes.  contiguous=True gives each rank one block
    of the sequence (streams stay sequential per rank); False interleaves (picture i -> i % world)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    if contiguous:
        base, extra = divmod(num_pictures, world_size)
        start = rank * base + min(rank, extra)
        return list(range(start, start + base + (1 if rank < extra else 0)))
    return list(range(rank, num_pictures, world_size))


def max_over_ranks(value_ms, device=None):
    """MAX all-reduce of a per-rank elapsed time (milliseconds); a no-op without a process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(local_count):
    """Sum of per-rank picture counts (whole-job throughput numerator)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(local_count)
    t = torch.tensor([int(local_count)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def bind_process_to_gpu(device_index, uuid=None):
    """Restricts this process to the CPUs NVML reports as local to the GPU (its NUMA node), so that the
    pinned host buffers allocated afterwards live in that node's memory and the host<->device copies of
    different ranks do not all cross the same socket interconnect.  One process per GPU; call it before
    allocating pinned memory.  Returns the number of CPUs bound to, 0 if nothing was changed (NVML or the
    affinity call unavailable)."""
    import os

    try:
        import pynvml as N

        N.nvmlInit()
        h = None
        if uuid:
            try:
                h = N.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                h = None
        if h is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = int(vis.split(",")[device_index]) if vis else device_index
            h = N.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        mask = N.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1} & allowed
        if not cpus or cpus == allowed:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
