"""Root-parallel merge on the CPU: world_size 2 over gloo.  Each rank's 'planner' is the oracle's gs search with
the rank's share of the budget and stream offset; the merged statistics must equal the single-process sum and
both ranks must choose the same action."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from helpers import ROOT, golden_tables, oracle_model, random_world_and_bag


def _worker(rank, world, port, n_sims, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    from pomdpy_b200 import dist as pd
    from pomdpy_b200.solvers import greedy_action
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    assert pd.world() == (rank, world)
    t = golden_tables("map-7-8")
    actual, bag, _ = random_world_and_bag(t, 5)
    per, off = pd.shard_budget(n_sims, rank, world)
    g = oracle.GsPOMCP(oracle_model(t), actual, 0, bag, seed=11)
    g.search(per, move=2, sim_offset=off, w0=8, growth=2, wmax=64)
    s = g.root_stats()
    n, q = pd.merge_host(s["n"], s["qfix"])
    a = greedy_action(n, q, s["legal"], 11, 2)
    out.put((rank, n.tolist(), q.tolist(), a))
    dist.barrier()
    dist.destroy_process_group()


def test_root_parallel_merge_world2():
    import oracle
    n_sims, world, port = 601, 2, 29511 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_sims, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:] == res[1][1:]                       # identical merged statistics and action on both ranks
    # single-process reference: the two shards searched one after the other, summed
    t = golden_tables("map-7-8")
    actual, bag, _ = random_world_and_bag(t, 5)
    tot_n, tot_q = 0, 0
    for r in range(world):
        g = oracle.GsPOMCP(oracle_model(t), actual, 0, bag, seed=11)
        g.search(301, move=2, sim_offset=r * 301, w0=8, growth=2, wmax=64)
        s = g.root_stats()
        tot_n, tot_q = tot_n + s["n"], tot_q + s["qfix"]
    assert res[0][1] == tot_n.tolist() and res[0][2] == tot_q.tolist()
    assert sum(res[0][1]) == 602
    from pomdpy_b200.solvers import greedy_action
    assert res[0][3] == greedy_action(tot_n, tot_q, s["legal"], 11, 2) == \
        oracle.GsPOMCP(oracle_model(t), actual, 0, bag, seed=11).greedy(move=2, stats=dict(n=tot_n, qfix=tot_q, legal=s["legal"]))
