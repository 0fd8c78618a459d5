"""GPU parity tests: the CUDA planner, called through the C ABI, against the CPU oracle (bit-exact) and the
golden vectors produced by the executed reference."""
import numpy as np
import pytest

import oracle
from helpers import MAPS, golden_tables, make_planner, oracle_model, random_world_and_bag

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("tag", MAPS)
def test_generate_step_golden(tag):
    """rock_model.py:451-465 -- device generate_step == the executed reference on the golden grid."""
    t = golden_tables(tag)
    pl = make_planner(t)
    si, so = t["step_in"], t["step_out"]
    state = (si[:, 0] | (si[:, 1] << np.uint64(8))).astype(np.uint32)
    ns, fl, rw = pl.generate_steps(state, si[:, 2].astype(np.uint8), si[:, 3], si[:, 4].astype(np.uint32),
                                   si[:, 5].astype(np.uint32))
    want_ns = (so[:, 0].astype(np.uint32) | (so[:, 1].astype(np.uint32) << np.uint32(8)))
    want_fl = (so[:, 2].astype(np.uint8) | (so[:, 3].astype(np.uint8) << 1) | (so[:, 5].astype(np.uint8) << 2)
               | (so[:, 6].astype(np.uint8) << 3))
    assert np.array_equal(ns, want_ns)
    assert np.array_equal(fl, want_fl)
    assert np.array_equal(rw, so[:, 4])
    pl.close()


def _pair(tag, wseed, n_bag=2000, **kw):
    t = golden_tables(tag)
    actual, bag, rng = random_world_and_bag(t, wseed, n_bag)
    okw = dict(max_depth=kw.get("max_depth", 100), dcap=kw.get("tree_depth_cap", 64),
               ucb_c=kw.get("ucb_coefficient", 3.0), discount=kw.get("discount", 0.95), seed=kw.get("seed", 0))
    om = oracle_model(t)
    g = oracle.GsPOMCP(om, actual, 0, bag, **okw)
    pl = make_planner(t, **kw)
    pl.set_world(actual, 0)
    pl.set_root_particles(bag, int(t["start_cell"]))
    return t, om, g, pl, actual, bag, rng


def _assert_same_root(g, pl, n_sims=None):
    a, b = g.root_stats(), pl.root_stats()
    assert a["legal"] == b["legal"]
    assert np.array_equal(a["n"], b["n"]), (a["n"], b["n"])
    assert np.array_equal(a["qfix"], b["qfix"]), (a["qfix"], b["qfix"])
    if n_sims is not None:
        assert int(b["n"].sum()) == n_sims and b["sims_done"] == n_sims
    return b


SEARCH_CASES = [
    # tag, seed, n_sims, (w0, growth, wmax), extra planner kwargs
    ("map-7-8", 1, 1, (1, 1, 1), {}),
    ("map-7-8", 2, 300, (1, 1, 1), {}),                       # strictly sequential POMCP
    ("map-7-8", 3, 500, (8, 2, 64), {}),
    ("map-7-8", 4, 2000, (32, 1, 32), dict(ucb_coefficient=5.0, discount=0.9, max_depth=20)),
    ("map-11-11", 5, 3000, (16, 2, 512), {}),
    ("map-15-15", 6, 5000, (64, 2, 2048), {}),
    ("map-15-15", 7, 20000, (256, 4, 8192), {}),
    ("map-7-8", 8, 3000, (4, 2, 256), dict(tree_depth_cap=3)),  # tree depth cap reached
    ("map-7-8", 9, 1500, (4, 2, 128), dict(max_depth=2)),       # search horizon reached inside the tree
    ("map-7-8", 10, 800, (8, 2, 64), dict(ucb_coefficient=0.0)),
    ("map-15-15", 11, 4097, (1000, 3, 3000), {}),               # ragged generations
]


@pytest.mark.parametrize("tag,seed,n_sims,sched,kw", SEARCH_CASES)
def test_search_root_stats_bit_exact(tag, seed, n_sims, sched, kw):
    """One search: root (visit_count, sum q) identical to the oracle's generation-synchronous POMCP, as are the
    statistics of the depth-1 nodes and the work counters."""
    w0, growth, wmax = sched
    kw = dict(kw, seed=100 + seed, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 16)
    t, om, g, pl, actual, bag, rng = _pair(tag, seed, **kw)
    g.search(n_sims, move=3, w0=w0, growth=growth, wmax=wmax)
    pl.search(n_sims, move=3)
    _assert_same_root(g, pl, n_sims)
    for a in range(om.n_actions):
        for o in (0, 1):
            x, y = g.node_stats([(a, o)]), pl.node_stats([(a, o)])
            assert (x is None) == (y is None)
            if x is not None:
                assert x["cell"] == y["cell"] and x["legal"] == y["legal"]
                assert np.array_equal(x["n"], y["n"]) and np.array_equal(x["qfix"], y["qfix"])
    co, cg = g.counters(), pl.counters()
    for k in ("levels", "rollout_steps", "created", "nodes", "generations"):
        assert co[k] == cg[k], (k, co[k], cg[k])
    pl.close()


def test_search_is_deterministic_and_offsets_differ():
    t = golden_tables("map-11-11")
    actual, bag, _ = random_world_and_bag(t, 77)
    res = []
    for off in (0, 0, 1 << 20):
        pl = make_planner(t, seed=5, gen_w0=64, gen_growth=2, gen_wmax=4096, node_capacity=1 << 16)
        pl.set_world(actual, 0)
        pl.set_root_particles(bag, int(t["start_cell"]))
        pl.search(10000, move=0, sim_offset=off)
        res.append(pl.root_stats())
        pl.close()
    assert np.array_equal(res[0]["n"], res[1]["n"]) and np.array_equal(res[0]["qfix"], res[1]["qfix"])
    assert not np.array_equal(res[0]["qfix"], res[2]["qfix"])


@pytest.mark.parametrize("tag,seed,n_sims,sched", [("map-7-8", 21, 400, (8, 2, 64)), ("map-15-15", 22, 3000, (64, 2, 1024))])
def test_episode_with_belief_updates_bit_exact(tag, seed, n_sims, sched):
    """Several moves: search -> greedy action -> real step -> belief update (K2) -> search in the re-rooted tree."""
    w0, growth, wmax = sched
    kw = dict(seed=seed, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 17)
    t, om, g, pl, actual, bag, rng = _pair(tag, seed, **kw)
    cell, rocks, sampled = int(t["start_cell"]), int(bag[0] >> 8), 0
    for mv in range(12):
        g.search(n_sims, move=mv, w0=w0, growth=growth, wmax=wmax)
        pl.search(n_sims, move=mv)
        st = _assert_same_root(g, pl)
        a = g.greedy(move=mv, stats=st)
        o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
        cell, rocks = o.next_cell, o.next_rocks
        if o.terminal or not o.legal:
            break
        if a == 4:
            sampled |= 1 << int(t["cell"][cell])
        uo = g.update(a, o.obs_good, sampled, move=mv)
        ug = pl.advance(a, o.obs_good, sampled, move=mv)
        assert uo["status"] == ug["status"] and uo["accepted"] == ug["accepted"] and uo["tries"] == ug["tries"]
        gb, gcell = pl.root_particles()
        assert gcell == g.root_cell()
        assert np.array_equal(gb, g.bag())
    pl.close()


def test_belief_update_low_acceptance_and_fresh_root():
    """CHECK of a far rock whose real observation was wrong: acceptance ~ 1 - p; the observed child was never
    expanded (no search) -> fresh root node."""
    t = golden_tables("map-15-15")
    actual, bag, rng = random_world_and_bag(t, 31)
    om = oracle_model(t)
    g = oracle.GsPOMCP(om, actual, 0, bag, seed=9)
    pl = make_planner(t, seed=9)
    pl.set_world(actual, 0)
    pl.set_root_particles(bag, int(t["start_cell"]))
    k = 14
    wrong = 0 if (actual >> k) & 1 else 1
    uo = g.update(5 + k, wrong, 0, move=0)
    ug = pl.advance(5 + k, wrong, 0, move=0)
    assert uo == ug and ug["status"] == 1 and ug["tries"] > 2000
    assert np.array_equal(pl.root_particles()[0], g.bag())
    # impossible observation: rock under the robot checked at distance 0 (p = 1), observation negated
    t7 = golden_tables("map-7-8")
    actual, bag, rng = random_world_and_bag(t7, 32)
    cell = int(np.argmax(t7["cell"] == 0))
    bag = (bag & np.uint32(0xFFFFFF00)) | np.uint32(cell)
    g = oracle.GsPOMCP(oracle_model(t7), actual, 0, bag, root_cell=cell, seed=9)
    pl2 = make_planner(t7, seed=9)
    pl2.set_world(actual, 0)
    pl2.set_root_particles(bag, cell)
    wrong = 0 if actual & 1 else 1
    uo = g.update(5, wrong, 0, move=0, max_tries=50000)
    ug = pl2.advance(5, wrong, 0, move=0, max_tries=50000)
    assert uo == ug and ug["status"] == 2 and ug["accepted"] == 0 and ug["tries"] == 50000
    pl.close()
    pl2.close()


def test_full_size_properties():
    """RockSample(15,15), 1M simulations (BASELINE config 4 at one GPU): size-independent invariants."""
    t = golden_tables("map-15-15")
    actual, bag, _ = random_world_and_bag(t, 41)
    n_sims = 1_000_000
    pl = make_planner(t, seed=1, gen_w0=1024, gen_growth=2, gen_wmax=1 << 18, node_capacity=1 << 21)
    pl.set_world(actual, 0)
    pl.set_root_particles(bag, int(t["start_cell"]))
    pl.search(n_sims)
    r = pl.root_stats()
    assert int(r["n"].sum()) == n_sims and r["sims_done"] == n_sims
    legal = np.array([(r["legal"] >> a) & 1 for a in range(pl.n_actions)])
    assert np.all(r["n"][legal == 0] == 0) and np.all(r["n"][legal == 1] > 0)
    c = pl.counters()
    assert c["created"] + 1 == c["nodes"] and c["created"] <= n_sims
    # flow conservation: visits entering a child <= visits of the edge above it, children of one edge sum to <= it
    for a in range(pl.n_actions):
        tot = 0
        for o in (0, 1):
            ch = pl.node_stats([(a, o)])
            if ch is not None:
                tot += int(ch["n"].sum())
        assert tot <= int(r["n"][a])
    # |mean q| bounded by the largest discounted return of the model
    mean = r["qfix"] / 2.0 ** 24 / np.maximum(r["n"], 1)
    assert np.all(np.abs(mean) <= 10 * 16 + 1e-9)
    pl.close()


def test_errors_are_loud():
    from pomdpy_b200 import native
    t = golden_tables("map-7-8")
    pl = native.Planner()
    with pytest.raises(native.PomcpError):
        pl.set_root_particles(np.zeros(4, dtype=np.uint32), 0)      # model not set
    pl.set_model_rocksample(int(t["rows"]), int(t["cols"]), t["cell"], int(t["start_cell"]), t["thr53"], t["rewards"])
    with pytest.raises(native.PomcpError):
        pl.search(10)                                               # no root belief
    with pytest.raises(native.PomcpError):
        pl.set_root_particles(np.zeros(0, dtype=np.uint32), 0)      # empty belief
    with pytest.raises(native.PomcpError):
        native.Planner(tree_depth_cap=1)
    pl.close()


@pytest.mark.parametrize("tag,seed,n_sims,sched,rollouts", [("map-7-8", 51, 1000, (8, 2, 128), False),
                                                            ("map-15-15", 52, 20000, (64, 4, 8192), False),
                                                            ("map-7-8", 53, 1000, (8, 2, 128), True),
                                                            ("map-11-11", 54, 6000, (1, 1, 1), True),
                                                            ("map-15-15", 55, 20000, (64, 4, 8192), True)])
def test_preferred_actions_episode_bit_exact(tag, seed, n_sims, sched, rollouts):
    """--preferred_actions (rock_position_history.py): per-node rock statistics and smart action sets on the
    device == oracle, over several moves with belief updates (fresh and existing roots); with `rollouts` the rollout
    policy follows the heuristic on the rollout's own history as well."""
    w0, growth, wmax = sched
    if rollouts and n_sims == 6000:
        n_sims = 600                                        # strictly sequential case: keep the oracle quick
    kw = dict(seed=seed, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 19)
    t = golden_tables(tag)
    actual, bag, rng = random_world_and_bag(t, seed)
    om = oracle_model(t)
    g = oracle.GsPOMCP(om, actual, 0, bag, seed=seed)
    g.set_preferred(True, rollouts=rollouts)
    pl = make_planner(t, **kw)
    pl.set_preferred_actions(True, t["prob"], rollouts=rollouts)
    pl.set_world(actual, 0)
    pl.set_root_particles(bag, int(t["start_cell"]))
    cell, rocks, sampled = int(t["start_cell"]), int(bag[0] >> 8), 0
    for mv in range(10):
        g.search(n_sims, move=mv, w0=w0, growth=growth, wmax=wmax)
        assert pl.search(n_sims, move=mv) == 0                     # 1 would mean: arena full, subtree dropped
        st = _assert_same_root(g, pl)
        assert bin(st["legal"]).count("1") < om.n_actions          # the in-tree set is restricted
        for a in range(om.n_actions):
            for o in (0, 1):
                x, y = g.node_stats([(a, o)]), pl.node_stats([(a, o)])
                assert (x is None) == (y is None)
                if x is not None:
                    assert x["legal"] == y["legal"] and np.array_equal(x["n"], y["n"]) and np.array_equal(x["qfix"], y["qfix"])
        a = g.greedy(move=mv, stats=st)
        o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
        cell, rocks = o.next_cell, o.next_rocks
        if o.terminal or not o.legal:
            break
        if a == 4:
            sampled |= 1 << int(t["cell"][cell])
        assert g.update(a, o.obs_good, sampled, move=mv) == pl.advance(a, o.obs_good, sampled, move=mv)
        assert np.array_equal(pl.root_particles()[0], g.bag())
    pl.close()


def test_cuda_planner_reproduces_committed_gs_pins():
    """The CUDA planner against the COMMITTED fixtures of the specification (tests/golden/gs_pins.json): no oracle in
    the loop -- root statistics, greedy actions, counters, belief updates of eleven fixed searches."""
    import json
    import os
    from helpers import GOLDEN
    from pomdpy_b200 import native
    from pomdpy_b200.solvers.pomcp import greedy_action
    from pomdpy_b200.tabular import random_pomdp, tiger
    with open(os.path.join(GOLDEN, "gs_pins.json")) as f:
        pins = json.load(f)["pins"]
    for pin in pins:
        c = pin["case"]
        w0, growth, wmax = c["sched"]
        kw = dict(seed=c["seed"], gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 17,
                  max_particle_count=500)
        if c["kind"] == "rs":
            t = golden_tables(c["tag"])
            actual, bag, _ = random_world_and_bag(t, c["seed"])
            pl = make_planner(t, **kw)
            if c["pref"]:
                pl.set_preferred_actions(True, t["prob"], rollouts=c["pref"] == 2)
            pl.set_world(actual, 0)
            pl.set_root_particles(bag, int(t["start_cell"]))
        else:
            m = tiger() if c["model"] == "tiger" else random_pomdp(*[int(x) for x in c["model"].split("-")[1:]], seed=c["seed"])
            bag = m.sample_init_states_packed(1000, np.random.RandomState(c["seed"]))
            pl = native.Planner(ucb_coefficient=c["ucb"], max_depth=40, **kw)
            pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
            pl.set_root_particles(bag)
        for move, want in enumerate(pin["moves"]):
            pl.search(c["n_sims"], move=move)
            st = pl.root_stats()
            assert st["n"].tolist() == want["n"] and st["qfix"].tolist() == want["qfix"] and st["legal"] == want["legal"], c
            a = greedy_action(st["n"], st["qfix"], st["legal"], c["seed"], move)
            cnt = pl.counters()
            assert a == want["action"] and cnt["nodes"] == want["nodes"], c
            u = pl.advance(a, move & 1, 0, move=move)
            assert [u["status"], u["tries"], u["accepted"]] == want["update"], c
            if u["status"] == 2:
                break
            assert int(pl.root_particles()[0].astype(np.uint64).sum()) == want["bag_sum"], c
        pl.close()
