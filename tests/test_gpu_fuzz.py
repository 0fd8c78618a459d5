"""Randomised parity: the CUDA planner against the CPU oracle over many small random configurations (budgets, generation
schedules, horizons, depth caps, UCB constants, discounts, belief sizes, preferred-action modes, tabular shapes).
Every case must match bit for bit: root statistics, counters, belief updates."""
import numpy as np
import pytest

import oracle
from helpers import MAPS, golden_tables, make_planner, oracle_model, random_world_and_bag
from pomdpy_b200 import native
from pomdpy_b200.tabular import random_pomdp

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _same(a, b, what):
    assert a["legal"] == b["legal"], what
    assert np.array_equal(a["n"], b["n"]), (what, a["n"], b["n"])
    assert np.array_equal(a["qfix"], b["qfix"]), (what, a["qfix"], b["qfix"])


def _config(rng):
    w0 = int(rng.choice([1, 2, 3, 8, 32, 100, 500]))
    return dict(n_sims=int(rng.choice([1, 2, 17, 100, 700, 2500, 6000])), w0=w0, growth=int(rng.choice([1, 2, 3, 4])),
                wmax=int(max(w0, rng.choice([1, 16, 128, 1024, 4096]))), max_depth=int(rng.choice([1, 2, 5, 30, 100])),
                dcap=int(rng.choice([2, 3, 8, 64])), ucb=float(rng.choice([0.0, 0.5, 3.0, 10.0])),
                discount=float(rng.choice([0.5, 0.95, 1.0])), n_bag=int(rng.choice([1, 7, 500, 2000])),
                seed=int(rng.integers(0, 1 << 62)))


@pytest.mark.parametrize("case", range(30))
def test_fuzz_rocksample(case):
    rng = np.random.default_rng(9000 + case)
    c = _config(rng)
    tag = MAPS[case % len(MAPS)]
    pref = int(rng.choice([0, 0, 1, 2]))
    t = golden_tables(tag)
    actual, bag, _ = random_world_and_bag(t, 100 + case, c["n_bag"])
    sampled0 = int(rng.integers(0, 1 << (int(t["cell"].max()) + 1))) if case % 3 == 0 else 0
    om = oracle_model(t)
    g = oracle.GsPOMCP(om, actual, sampled0, bag, max_depth=c["max_depth"], dcap=c["dcap"], ucb_c=c["ucb"],
                       discount=c["discount"], seed=c["seed"])
    pl = make_planner(t, max_depth=c["max_depth"], tree_depth_cap=c["dcap"], ucb_coefficient=c["ucb"],
                      discount=c["discount"], seed=c["seed"], gen_w0=c["w0"], gen_growth=c["growth"], gen_wmax=c["wmax"],
                      node_capacity=1 << 16, max_particle_count=300)
    if pref:
        g.set_preferred(True, rollouts=pref == 2)
        pl.set_preferred_actions(True, t["prob"], rollouts=pref == 2)
    pl.set_world(actual, sampled0)
    pl.set_root_particles(bag, int(t["start_cell"]))
    what = (case, tag, pref, c)
    for move in range(3):
        g.search(c["n_sims"], move=move, w0=c["w0"], growth=c["growth"], wmax=c["wmax"])
        pl.search(c["n_sims"], move=move)
        a, b = g.root_stats(), pl.root_stats()
        _same(a, b, what)
        gc, pc = g.counters(), pl.counters()
        for k in ("levels", "rollout_steps", "created", "nodes"):
            assert gc[k] == pc[k], (what, k, gc[k], pc[k])
        act = g.greedy(move, a)
        obs = int(rng.integers(0, 2))
        u = g.update(act, obs, sampled0, move=move, need=300, max_tries=30000)
        v = pl.advance(act, obs, sampled0, move=move, max_tries=30000)
        assert (u["status"], u["tries"], u["accepted"]) == (v["status"], v["tries"], v["accepted"]), what
        if u["status"] == 2:
            break
        assert np.array_equal(pl.root_particles()[0], g.bag()), what
    pl.close()


@pytest.mark.parametrize("case", range(16))
def test_fuzz_tabular(case):
    _fuzz_tabular(case, [1, 2, 3, 8, 32])


@pytest.mark.parametrize("case", range(16, 26))
def test_fuzz_tabular_many_observations(case):
    """The same over observation counts beyond the direct-indexed child table (hashed table)."""
    _fuzz_tabular(case, [33, 64, 100, 513, 2048])


def _fuzz_tabular(case, z_choices):
    rng = np.random.default_rng(7000 + case)
    c = _config(rng)
    S, A, Z = int(rng.choice([1, 2, 5, 33, 80, 200])), int(rng.choice([1, 2, 3, 9, 32])), int(rng.choice(z_choices))
    m = random_pomdp(S, A, Z, seed=case, sparsity=float(rng.choice([0.0, 0.5, 0.9])))
    if case % 4 == 1:                                     # some episode-ending actions / absorbing states
        m.term_action_mask = 1 << (A - 1)
        m.term_state[rng.integers(0, S)] = 1
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    bag = m.sample_init_states_packed(c["n_bag"], np.random.RandomState(case))
    g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=c["max_depth"], dcap=c["dcap"], ucb_c=c["ucb"], discount=c["discount"],
                       seed=c["seed"])
    pl = native.Planner(max_depth=c["max_depth"], tree_depth_cap=c["dcap"], ucb_coefficient=c["ucb"],
                        discount=c["discount"], seed=c["seed"], gen_w0=c["w0"], gen_growth=c["growth"], gen_wmax=c["wmax"],
                        node_capacity=1 << 16, max_particle_count=300)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl.set_root_particles(bag)
    what = (case, S, A, Z, c)
    for move in range(3):
        g.search(c["n_sims"], move=move, w0=c["w0"], growth=c["growth"], wmax=c["wmax"])
        pl.search(c["n_sims"], move=move)
        a, b = g.root_stats(), pl.root_stats()
        _same(a, b, what)
        act = g.greedy(move, a)
        obs = int(rng.integers(0, Z))
        u = g.update(act, obs, 0, move=move, need=300, max_tries=30000)
        v = pl.advance(act, obs, 0, move=move, max_tries=30000)
        assert (u["status"], u["tries"], u["accepted"]) == (v["status"], v["tries"], v["accepted"]), what
        if u["status"] == 2:
            break
        assert np.array_equal(pl.root_particles()[0], g.bag()), what
    pl.close()
