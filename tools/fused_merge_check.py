"""Run under torchrun on >= 2 GPUs: the in-kernel root merge against the NCCL all-reduce of the same statistics.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/fused_merge_check.py"""
import os
import sys

sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import torch
import torch.distributed as dist
from helpers import golden_tables, make_planner, random_world_and_bag
from pomdpy_b200 import dist as dist_mod

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
t = golden_tables("map-11-11")
actual, bag, _ = random_world_and_bag(t, 5)
pl = make_planner(t, seed=3, gen_w0=256, gen_growth=2, gen_wmax=8192, node_capacity=1 << 20, device=local)
pl.set_world(actual, 0)
pl.set_root_particles(bag, int(t["start_cell"]))
fm = dist_mod.try_fused_merge(pl)
assert fm is not None, "symmetric memory rendezvous failed"
for move in range(6):
    n_sims = 20000 + 1000 * move
    pl.search(n_sims, move=move, sim_offset=rank * n_sims)
    got = pl.merged_root_stats()
    loc = pl.root_stats()
    want_n, want_q = dist_mod.merge_device(pl, len(loc["n"]))
    assert np.array_equal(got["n"], want_n) and np.array_equal(got["qfix"], want_q), (rank, move, got["n"], want_n)
    if move == 2:                                        # searches with very different durations on the ranks
        extra = 200000 if rank == 0 else 100
        pl.search(extra, move=100, sim_offset=rank * 1000000)
        g2 = pl.merged_root_stats()
        w2n, w2q = dist_mod.merge_device(pl, len(loc["n"]))
        assert np.array_equal(g2["n"], w2n) and np.array_equal(g2["qfix"], w2q)
    a = int(np.argmax(np.where(got["n"] > 0, got["qfix"] / np.maximum(got["n"], 1), -1e30)))
    pl.advance(a, move & 1, 0, move=move)
dist.barrier()
if rank == 0:
    print("FUSED-MERGE-OK world", world)
dist.destroy_process_group()
