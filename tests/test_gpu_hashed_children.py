"""GPU parity of the hashed child table (north_star: "hashed child table for generic |O|").  The
reference keeps the children of an action node in a dict keyed by observation
(pomdpy/discrete_pomdp/discrete_observation_mapping.py:12-47), so |Z| is unbounded; on the device a tabular model with
more than 32 observations gets an open-addressing table over (node, action, observation) instead of the direct-indexed
[node][action][Z] rows.  Checked here through the C ABI against the CPU oracle (whose observation map is a chain per
action node): bit-exact root / inner-node statistics, counters and beliefs for |Z| = 64, 300, 2048; the hashed table
forced on models with |Z| <= 32 gives what the direct table gives; wide generations (thousands of simulations insert
into the table at once)."""
import numpy as np
import pytest

import oracle
from pomdpy_b200 import native
from pomdpy_b200.tabular import random_pomdp, tiger

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

MODELS = {
    "tiger": lambda: tiger(),
    "rand-64-16-32": lambda: random_pomdp(64, 16, 32, seed=2),
    "rand-48-4-64": lambda: random_pomdp(48, 4, 64, seed=10),
    "rand-24-3-300": lambda: random_pomdp(24, 3, 300, seed=11),
    "rand-16-2-2048": lambda: random_pomdp(16, 2, 2048, seed=12),
    "rand-200-32-40": lambda: random_pomdp(200, 32, 40, seed=13),        # all 32 action lanes, guide tables
}


def _pair(name, seed=0, n_bag=2000, hashed=False, **kw):
    m = MODELS[name]()
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    bag = m.sample_init_states_packed(n_bag, np.random.RandomState(seed))
    g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=kw.get("max_depth", 100), dcap=kw.get("tree_depth_cap", 64),
                       ucb_c=kw.get("ucb_coefficient", 3.0), discount=kw.get("discount", 0.95), seed=kw.get("seed", 0))
    pl = native.Planner(hashed_children=hashed, **kw)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl.set_root_particles(bag)
    return m, om, g, pl


def _same(a, b):
    assert a["legal"] == b["legal"]
    assert np.array_equal(a["n"], b["n"]), (a["n"], b["n"])
    assert np.array_equal(a["qfix"], b["qfix"]), (a["qfix"], b["qfix"])


def _compare_tree(m, g, pl, max_children=400):
    """Root, every (action, observation) child of the root (up to max_children existing ones) and one grandchild layer."""
    _same(g.root_stats(), pl.root_stats())
    gc, pc = g.counters(), pl.counters()
    for k in ("levels", "rollout_steps", "created", "nodes", "generations"):
        assert gc[k] == pc[k], (k, gc[k], pc[k])
    checked = 0
    for act in range(m.num_actions):
        for z in range(m.num_observations):
            x, y = g.node_stats([(act, z)]), pl.node_stats([(act, z)])
            assert (x is None) == (y is None), (act, z)
            if x is None or checked >= max_children:
                continue
            _same(x, y)
            checked += 1
            for a2 in range(min(m.num_actions, 2)):
                for z2 in (0, m.num_observations // 2, m.num_observations - 1):
                    x2, y2 = g.node_stats([(act, z), (a2, z2)]), pl.node_stats([(act, z), (a2, z2)])
                    assert (x2 is None) == (y2 is None), (act, z, a2, z2)
                    if x2 is not None:
                        _same(x2, y2)
    return checked


@pytest.mark.parametrize("name", ["rand-48-4-64", "rand-24-3-300", "rand-16-2-2048", "rand-200-32-40"])
def test_generate_step_many_observations(name):
    m = MODELS[name]()
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl = native.Planner()
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    rng = np.random.RandomState(5)
    n = 3000
    s = rng.randint(0, m.num_states, size=n).astype(np.uint32)
    a = rng.randint(0, m.num_actions, size=n).astype(np.uint8)
    w = rng.randint(0, 1 << 32, size=(2, n), dtype=np.uint64).astype(np.uint32)
    w[1, :8] = [0, 0xFFFFFFFF, 0xFFFFFFFE, 1 << 31, (1 << 31) - 1, 1, 2, 3]
    ns, ob, rw, tm = pl.tabular_steps(s, a, w[0], w[1])
    assert int(ob.max()) > 32
    for i in range(n):
        assert (int(ns[i]), int(ob[i]), float(rw[i]), int(tm[i])) == om.step(int(s[i]), int(a[i]), int(w[0, i]), int(w[1, i]))
    pl.close()


CASES = [
    # model, n_sims, (w0, growth, wmax), forced hash, planner kwargs
    ("rand-48-4-64", 400, (1, 1, 1), False, dict(max_depth=30)),
    ("rand-48-4-64", 30000, (32, 2, 4096), False, dict(max_depth=30, ucb_coefficient=10.0)),
    ("rand-48-4-64", 3000, (8, 2, 64), False, dict(max_depth=20, tree_depth_cap=4)),
    ("rand-24-3-300", 30000, (32, 4, 8192), False, dict(max_depth=25, ucb_coefficient=6.0, seed=3)),
    ("rand-16-2-2048", 20000, (16, 4, 4096), False, dict(max_depth=20, ucb_coefficient=4.0, discount=0.9)),
    ("rand-200-32-40", 40000, (64, 4, 8192), False, dict(max_depth=25, ucb_coefficient=8.0)),
    # the hashed table on models the direct table serves: same specification, same results
    ("tiger", 50000, (64, 4, 16384), True, dict(ucb_coefficient=30.0, seed=7)),
    ("rand-64-16-32", 30000, (32, 2, 4096), True, dict(max_depth=20, ucb_coefficient=10.0)),
]


@pytest.mark.parametrize("name,n_sims,sched,hashed,kw", CASES)
def test_hashed_search_bit_exact(name, n_sims, sched, hashed, kw):
    w0, growth, wmax = sched
    m, om, g, pl = _pair(name, hashed=hashed, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 17, **kw)
    g.search(n_sims, w0=w0, growth=growth, wmax=wmax)
    pl.search(n_sims)
    assert int(pl.root_stats()["n"].sum()) == n_sims
    checked = _compare_tree(m, g, pl)
    assert checked > 0
    pl.close()


def test_hashed_equals_direct_table():
    """Without the oracle in the loop: the same planner with the direct-indexed and with the hashed table."""
    m = MODELS["rand-64-16-32"]()
    bag = m.sample_init_states_packed(2000, np.random.RandomState(4))
    out = []
    for hashed in (False, True):
        pl = native.Planner(hashed_children=hashed, gen_w0=64, gen_growth=4, gen_wmax=16384, node_capacity=1 << 18,
                            max_depth=20, ucb_coefficient=10.0, max_particle_count=1000, seed=21)
        pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
        pl.set_root_particles(bag)
        rec = []
        for move in range(3):
            pl.search(100000, move=move)
            st = pl.root_stats()
            rec.append((st["n"].tolist(), st["qfix"].tolist(), pl.counters()["nodes"]))
            kids = []
            for act in range(m.num_actions):
                for z in range(m.num_observations):
                    y = pl.node_stats([(act, z)])
                    kids.append(None if y is None else (y["n"].tolist(), y["qfix"].tolist()))
            rec.append(kids)
            act = int(np.argmax(st["n"]))
            u = pl.advance(act, move % m.num_observations, 0, move=move)
            rec.append((u["status"], u["tries"], u["accepted"]))
            if u["status"] == 2:
                break
            rec.append(pl.root_particles()[0].tolist())
        out.append(rec)
        pl.close()
    assert out[0] == out[1]


@pytest.mark.parametrize("name,kw", [("rand-48-4-64", dict(max_depth=30, ucb_coefficient=10.0)),
                                     ("rand-24-3-300", dict(max_depth=25, ucb_coefficient=6.0))])
def test_hashed_search_update_loop_bit_exact(name, kw):
    """Several real steps: search, greedy action, belief update with a sampled observation (looked up in the hashed
    table by the belief-update kernel), re-root; then a tree reset (the whole table is cleared) and a search again."""
    m, om, g, pl = _pair(name, gen_w0=32, gen_growth=2, gen_wmax=2048, node_capacity=1 << 17, max_particle_count=1500,
                         **kw)
    rng = np.random.RandomState(11)
    state = int(rng.randint(m.num_states))
    seen_existing = False
    for move in range(6):
        g.search(6000, move=move, w0=32, growth=2, wmax=2048)
        pl.search(6000, move=move)
        a = g.root_stats()
        _same(a, pl.root_stats())
        act = g.greedy(move, a)
        ns, z, _, _ = om.step(state, act, int(rng.randint(1 << 32)), int(rng.randint(1 << 32)))
        state = ns
        u = g.update(act, z, 0, move=move, need=1500)
        v = pl.advance(act, z, 0, move=move)
        assert (u["status"], u["tries"], u["accepted"]) == (v["status"], v["tries"], v["accepted"])
        if u["status"] == 2:
            break
        seen_existing |= u["status"] == 0
        bag, _ = pl.root_particles()
        assert np.array_equal(bag, g.bag())
        _same(g.root_stats(), pl.root_stats())
    assert seen_existing                                   # at least one re-root found its child in the table
    # a fresh tree over a new belief: stale keys of the old tree must be gone
    bag = m.sample_init_states_packed(1500, np.random.RandomState(99))
    om2 = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    g2 = oracle.GsPOMCP(om2, 0, 0, bag, max_depth=kw["max_depth"], ucb_c=kw["ucb_coefficient"], seed=0)
    pl.set_root_particles(bag)
    g2.search(8000, move=0, w0=32, growth=2, wmax=2048)
    pl.search(8000, move=0)
    _same(g2.root_stats(), pl.root_stats())
    assert g2.counters()["nodes"] == pl.counters()["nodes"]
    pl.close()


def test_hashed_wide_generation():
    """One generation of 2^18 simulations over |Z| = 64: every block inserts into the table at once (the oracle takes a
    few seconds).  Root and depth-1 statistics bit-exact, node counts equal."""
    w0, growth, wmax = 16384, 16, 262144
    m, om, g, pl = _pair("rand-48-4-64", gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 20,
                         max_depth=30, ucb_coefficient=10.0, seed=5)
    n_sims = 400000
    g.search(n_sims, w0=w0, growth=growth, wmax=wmax)
    pl.search(n_sims)
    checked = _compare_tree(m, g, pl, max_children=256)
    assert checked >= 64
    pl.close()


def test_observation_count_limits():
    """|Z| = 2048 is the ceiling (11 bits of a path entry); above it the ABI refuses the model."""
    rng = np.random.RandomState(0)
    S, A, Z = 2, 1, 2049
    Tc = np.full((A, S, S), 0xFFFFFFFF, dtype=np.uint32)
    Tc[:, :, 0] = 1 << 31
    Oc = np.minimum((np.arange(1, Z + 1, dtype=np.float64) / Z * 4294967296.0), 4294967295.0).astype(np.uint32)
    Oc = np.ascontiguousarray(np.broadcast_to(Oc, (A, S, Z)))
    Oc[:, :, -1] = 0xFFFFFFFF
    R = rng.rand(A, S)
    pl = native.Planner()
    with pytest.raises(native.PomcpError):
        pl.set_model_tabular(Tc, Oc, R, 0, None)
    pl.set_model_tabular(Tc, np.ascontiguousarray(Oc[:, :, 1:]), R, 0, None)        # 2048 observations: accepted
    pl.close()


def test_switching_models_on_one_handle():
    """One planner handle, three models in turn: hashed table (|Z| = 64) -> direct table (Tiger) -> hashed again; the
    child table is re-allocated on every switch and each search equals the oracle's."""
    pl = native.Planner(gen_w0=32, gen_growth=4, gen_wmax=4096, node_capacity=1 << 16, max_depth=30, ucb_coefficient=10.0)
    for name in ("rand-48-4-64", "tiger", "rand-24-3-300", "rand-64-16-32"):
        m = MODELS[name]()
        om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
        bag = m.sample_init_states_packed(1000, np.random.RandomState(3))
        g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=30, ucb_c=10.0, seed=0)
        pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
        pl.set_root_particles(bag)
        g.search(15000, w0=32, growth=4, wmax=4096)
        pl.search(15000)
        _same(g.root_stats(), pl.root_stats())
        assert g.counters()["nodes"] == pl.counters()["nodes"]
        z = int(m.num_observations) - 1
        for act in range(min(m.num_actions, 3)):
            x, y = g.node_stats([(act, z)]), pl.node_stats([(act, z)])
            assert (x is None) == (y is None)
            if x is not None:
                _same(x, y)
    pl.close()
