"""World-size-2 test of the multi-GPU host logic on CPU (gloo): sharding, gather, max-over-ranks.  The engine each rank
runs here is the oracle (there is no GPU on this box); on a GPU box bench.py --gpus N runs the same code over NCCL."""
import os
import subprocess
import sys

import numpy as np

import _oracle as O
from sapling_b200.dist import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import _oracle as O, _fixtures as F
from sapling_b200.dist import sharded_query, max_over_ranks, my_shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=rank, world_size=world)
g = F.small_genomes()["rand20k"]
port = O.Port.from_memory(g, k=21)                      # every rank holds a full replica of the index
kmers = F.query_mix(g, 21, 5001)                         # odd length: ragged shards
lo, hi, local, full = sharded_query(lambda q: port.query_batch(q), kmers, rank, world, dist)
assert (lo, hi) == my_shard(len(kmers), rank, world) and len(local) == hi - lo
t = max_over_ranks(1.0 + rank, dist)
assert t == float(world), t
if rank == 0:
    np.save(sys.argv[2], full)
else:
    assert full is None
dist.barrier(); dist.destroy_process_group()
'''


def test_shard_bounds_cover_and_balance():
    for nq in (0, 1, 7, 1000, 50_000_001):
        for world in (1, 2, 3, 8):
            b = shard_bounds(nq, world)
            assert b[0][0] == 0 and b[-1][1] == nq
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in b]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_query_gloo(oracle_built, tmp_path):
    import _fixtures as F
    out = str(tmp_path / "full.npy")
    port_no = 29600 + os.getpid() % 200
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER, ROOT, out], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    g = F.small_genomes()["rand20k"]
    port = O.Port.from_memory(g, k=21)
    kmers = F.query_mix(g, 21, 5001)
    assert np.array_equal(np.load(out), port.query_batch(kmers))
    port.close()
