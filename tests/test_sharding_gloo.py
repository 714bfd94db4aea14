"""CPU, world_size 2, gloo: the host-side multi-GPU plumbing (batch sharding,
barrier, max-over-ranks timing).  The data path itself has no collective."""
import os
import socket
import subprocess
import sys
import textwrap

from heongpu_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 1024, 4097):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_plumbing():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import torch.distributed as dist
        from heongpu_b200 import sharding as S
        rank, world, _ = S.env_rank_world()
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lo, hi = S.shard_range(9, rank, world)
        S.barrier(world)
        total = S.sum_over_ranks(hi - lo, world)
        slow = S.max_over_ranks(10.0 + rank, world)
        assert total == 9.0, total
        assert slow == 11.0, slow
        S.barrier(world)
        dist.destroy_process_group()
        print("rank", rank, "ok", lo, hi)
    """ % ROOT)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, out + err
