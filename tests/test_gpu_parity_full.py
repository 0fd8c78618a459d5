"""Bit-exact parity at the sizes the benchmark is credited for (BASELINE configs 3 and 4): the CUDA planner, through
the C ABI, against the CPU oracle -- root and depth-1 statistics, work counters, belief updates -- from a fresh tree
and over consecutive moves of an episode (the re-rooted tree is a deep chain there).  Plus a build with tiny
per-block tables (-DARR_HT=64 -DPC_BITS=2) so that the crowded-table fallbacks of the backup and the conflict path
of the plan cache run as well."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from helpers import ROOT, golden_tables, make_planner, oracle_model, random_world_and_bag

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500)]


def _compare_nodes(g, pl, n_actions):
    a, b = g.root_stats(), pl.root_stats()
    assert a["legal"] == b["legal"]
    assert np.array_equal(a["n"], b["n"]), (a["n"], b["n"])
    assert np.array_equal(a["qfix"], b["qfix"]), (a["qfix"], b["qfix"])
    for act in range(n_actions):
        for o in (0, 1):
            x, y = g.node_stats([(act, o)]), pl.node_stats([(act, o)])
            assert (x is None) == (y is None), (act, o)
            if x is not None:
                assert x["cell"] == y["cell"] and x["legal"] == y["legal"]
                assert np.array_equal(x["n"], y["n"]) and np.array_equal(x["qfix"], y["qfix"]), (act, o)
    co, cg = g.counters(), pl.counters()
    for k in ("levels", "rollout_steps", "created", "nodes", "generations"):
        assert co[k] == cg[k], (k, co[k], cg[k])
    return b


def run_episode_parity(tag, n_sims, sched, moves, seed, node_capacity):
    """Fresh search, then `moves - 1` searches in the re-rooted tree; everything identical to the oracle."""
    w0, growth, wmax = sched
    t = golden_tables(tag)
    actual, bag, rng = random_world_and_bag(t, seed)
    om = oracle_model(t)
    g = oracle.GsPOMCP(om, actual, 0, bag, seed=seed)
    pl = make_planner(t, seed=seed, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=node_capacity)
    pl.set_world(actual, 0)
    pl.set_root_particles(bag, int(t["start_cell"]))
    cell, rocks, sampled = int(t["start_cell"]), int(bag[0] >> 8), 0
    for mv in range(moves):
        g.search(n_sims, move=mv, w0=w0, growth=growth, wmax=wmax)
        assert pl.search(n_sims, move=mv) == 0
        st = _compare_nodes(g, pl, om.n_actions)
        assert int(st["n"].sum()) >= n_sims and st["sims_done"] == n_sims
        a = g.greedy(move=mv, stats=st)
        o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
        cell, rocks = o.next_cell, o.next_rocks
        if o.terminal or not o.legal:
            break
        if a == 4:
            sampled |= 1 << int(t["cell"][cell])
        assert g.update(a, o.obs_good, sampled, move=mv) == pl.advance(a, o.obs_good, sampled, move=mv)
        gb, gcell = pl.root_particles()
        assert gcell == g.root_cell() and np.array_equal(gb, g.bag())
    c = pl.counters()
    pl.close()
    return c


def test_config4_rs15_1m_sims_bench_schedule_fresh_and_in_episode():
    """RockSample(15,15), 1 000 000 simulations per move on bench.py's schedule (the solver's rule for this budget:
    125 000 x32 <= 1 000 000 -- two generations from the fresh tree, one per move afterwards): a fresh tree and three
    more moves with the belief update in between."""
    from pomdpy_b200.solvers.pomcp import auto_schedule
    assert auto_schedule(1_000_000) == (125000, 32, 1000000)
    run_episode_parity("map-15-15", 1_000_000, auto_schedule(1_000_000), 4, seed=123, node_capacity=(1 << 22) + 4096)


def test_config4_rs15_1m_sims_four_generation_schedule():
    """The same at round 1's schedule (32768, 4, 524288): four generations from the fresh tree, two per move afterwards."""
    run_episode_parity("map-15-15", 1_000_000, (32768, 4, 1 << 19), 3, seed=124, node_capacity=(1 << 22) + 4096)


def test_config3_rs11_100k_sims():
    """RockSample(11,11), 100 000 simulations per move (BASELINE config 3), five moves, on the solver's schedule for this
    budget (12 500 x32 <= 100 000) and on round 1's (3125, 4, 50000)."""
    from pomdpy_b200.solvers.pomcp import auto_schedule
    run_episode_parity("map-11-11", 100_000, auto_schedule(100_000), 5, seed=321, node_capacity=1 << 20)
    run_episode_parity("map-11-11", 100_000, (3125, 4, 50000), 5, seed=322, node_capacity=1 << 20)


SMALL_TABLES = r'''
import sys
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
from test_gpu_parity_full import run_episode_parity
c = run_episode_parity("map-15-15", 300000, (8192, 4, 1 << 17), 4, seed=5, node_capacity=1 << 21)
c = run_episode_parity("map-7-8", 200000, (4096, 2, 1 << 16), 4, seed=6, node_capacity=1 << 20)
print("SMALL-TABLES-OK")
'''


def test_crowded_block_tables_variant_build():
    """The same comparison with 64-slot backup tables and a 4-slot plan cache per block: slots collide constantly, so
    the direct-to-arena fallbacks (backup) and the bypass path (plan cache) carry real traffic."""
    lib = os.path.join(ROOT, "pomdpy_b200", "libpomcp_b200_small.so")
    if not os.path.exists(lib):
        from pomdpy_b200 import build
        build.build(force=True, defines=("ARR_HT=64", "PC_BITS=2"), out=lib)
    env = dict(os.environ, POMCP_B200_LIB=lib)
    res = subprocess.run([sys.executable, "-c", "ROOT = %r\n" % ROOT + SMALL_TABLES], env=env, capture_output=True,
                         text=True, timeout=1200)
    assert res.returncode == 0 and "SMALL-TABLES-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
