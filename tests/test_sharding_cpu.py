"""world_size-2 gloo test of the N>1 host path (no GPU): disjoint, complete picture ownership and
the max-over-ranks timing reduction."""

import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_disjoint_and_complete():
    from vc2_conformance_b200.sharding import pictures_for_rank

    for n in (0, 1, 7, 120, 121):
        for world in (1, 2, 3, 8):
            for contiguous in (True, False):
                owned = [pictures_for_rank(n, r, world, contiguous) for r in range(world)]
                flat = sorted(i for o in owned for i in o)
                assert flat == list(range(n))
                assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_two_rank_gloo_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import torch.distributed as dist
        from vc2_conformance_b200.sharding import pictures_for_rank, max_over_ranks, gather_counts
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        mine = pictures_for_rank(121, rank, world)
        total = gather_counts(len(mine))
        slowest = max_over_ranks(10.0 + 5.0 * rank)
        assert total == 121, total
        assert slowest == 15.0, slowest
        if rank == 0:
            print("OK", total, slowest)
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 121 15.0" in out.stdout
