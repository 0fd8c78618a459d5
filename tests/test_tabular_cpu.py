"""CPU suite of the tabular-model path: the integer sampling rule, the oracle's generative step against the host
model, the Tiger matrices against the executed reference's (tests/golden/vi_tiger.json), planner invariants."""
import json
import os

import numpy as np

import oracle
from helpers import GOLDEN
from pomdpy_b200.tabular import TabularState, cumulative_thresholds, draw, random_pomdp, tiger


def oracle_tab(m):
    return oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)


def test_tiger_matrices_match_executed_reference():
    with open(os.path.join(GOLDEN, "vi_tiger.json")) as f:
        gold = json.load(f)
    m = tiger()
    assert np.array_equal(m.T, np.array(gold["T"]))
    assert np.array_equal(m.O, np.array(gold["O"]))
    assert np.array_equal(m.R, np.array(gold["R"]))
    assert m.term_action_mask == 0b110 and m.num_observations == 2


def test_cumulative_thresholds_and_draw_rule():
    t = cumulative_thresholds([[0.5, 0.5, 0.0], [0.0, 1.0, 0.0], [0.25, 0.0, 0.75], [0.0, 0.0, 1.0]])
    assert t[:, -1].tolist() == [0xFFFFFFFF] * 4
    assert t[0].tolist() == [1 << 31, 0xFFFFFFFF, 0xFFFFFFFF]
    for row, zero_cols in ((0, [2]), (1, [0, 2]), (2, [1]), (3, [0, 1])):
        for w in (0, 1, (1 << 30) - 1, 1 << 30, (1 << 31) - 1, 1 << 31, 0xFFFFFFFE, 0xFFFFFFFF):
            assert draw(t[row], w) not in zero_cols                     # zero-probability entries are never drawn
    assert draw(t[0], (1 << 31) - 1) == 0 and draw(t[0], 1 << 31) == 1
    # frequencies follow the probabilities
    rng = np.random.RandomState(0)
    p = rng.dirichlet(np.ones(7))
    thr = cumulative_thresholds(p)
    w = rng.randint(0, 1 << 32, size=200000, dtype=np.uint64)
    freq = np.bincount(np.searchsorted(thr, np.minimum(w, 0xFFFFFFFE), side="right"), minlength=7) / w.size
    assert np.abs(freq - p).max() < 5e-3


def test_oracle_step_equals_host_model_step():
    for m in (tiger(), random_pomdp(19, 5, 7, seed=4)):
        om = oracle_tab(m)
        rng = np.random.RandomState(1)
        for _ in range(3000):
            s, a = rng.randint(m.num_states), rng.randint(m.num_actions)
            w = rng.randint(0, 1 << 32, size=2, dtype=np.uint64)
            if rng.rand() < 0.05:
                w[rng.randint(2)] = rng.choice([0, 0xFFFFFFFF, 0xFFFFFFFE, 1 << 31])
            r, _ = m.generate_step(TabularState(s), a, words=w)
            ns, ob, rew, term = om.step(s, a, int(w[0]), int(w[1]))
            assert (ns, ob, rew, bool(term)) == (r.next_state.index, r.observation.bin_number, r.reward, r.is_terminal)


def test_gs_tabular_invariants_and_width_one():
    m = random_pomdp(12, 4, 5, seed=2)
    om = oracle_tab(m)
    bag = m.sample_init_states_packed(500, np.random.RandomState(0))
    g = oracle.GsPOMCP(om, 0, 0, bag, seed=9, max_depth=30)
    g.search(4000, w0=16, growth=2, wmax=512)
    st = g.root_stats()
    assert int(st["n"].sum()) == 4000 and st["legal"] == 0b1111
    c = g.counters()
    assert c["created"] == c["nodes"] - 1 and c["max_tree_depth"] <= 30
    # children statistics: visits of (a, z) children never exceed the edge count
    for a in range(4):
        tot = 0
        for z in range(5):
            ch = g.node_stats([(a, z)])
            if ch is not None:
                tot += int(ch["n"].sum())
        assert tot <= int(st["n"][a])
    # deterministic: same seed, same statistics; the schedule changes them
    g2 = oracle.GsPOMCP(om, 0, 0, bag, seed=9, max_depth=30)
    g2.search(4000, w0=16, growth=2, wmax=512)
    assert np.array_equal(g2.root_stats()["qfix"], st["qfix"])
    # belief update keeps only particles consistent with the observation
    a = g.greedy(0, st)
    u = g.update(a, 3, 0, need=200)
    assert u["status"] in (0, 1) and u["accepted"] == 200 and g.bag().max() < 12


def test_gs_observation_map_is_unbounded():
    """More than 32 observations: the oracle's observation map is a chain per action node, as the reference's is a dict
    (discrete_observation_mapping.py:12-47); children are found again under the observation they were created for,
    re-rooting lands on them, and the index range ends at 2048 (11 bits of a device path entry)."""
    m = random_pomdp(24, 3, 300, seed=11)
    om = oracle_tab(m)
    bag = m.sample_init_states_packed(500, np.random.RandomState(0))
    g = oracle.GsPOMCP(om, 0, 0, bag, seed=4, max_depth=25, ucb_c=6.0)
    g.search(8000, w0=16, growth=4, wmax=2048)
    st = g.root_stats()
    assert int(st["n"].sum()) == 8000
    kids = {(a, z): g.node_stats([(a, z)]) for a in range(3) for z in range(300)}
    seen = {k: v for k, v in kids.items() if v is not None}
    assert len(seen) > 32 and max(z for _, z in seen) > 32
    for a in range(3):                                   # every visit of edge a ended in one of its children (or was
        below = sum(int(v["n"].sum()) for (aa, _), v in seen.items() if aa == a)   # the visit that created one)
        made = sum(1 for (aa, _) in seen if aa == a)
        assert below <= int(st["n"][a]) and below + made >= int(st["n"][a]) - 300
    a0 = int(np.argmax(st["n"]))
    z0 = max((k for k in seen if k[0] == a0), key=lambda k: int(seen[k]["n"].sum()))[1]
    want = seen[(a0, z0)]
    u = g.update(a0, z0, 0, need=50)
    assert u["status"] == 0 and np.array_equal(g.root_stats()["n"], want["n"])
    # limits
    two = random_pomdp(2, 1, 2048, seed=0)
    assert oracle_tab(two) is not None
    import pytest
    bad = random_pomdp(2, 1, 2049, seed=0)
    with pytest.raises(Exception):
        oracle_tab(bad)


def test_tiger_planner_prefers_listening_then_the_right_door():
    m = tiger()
    om = oracle_tab(m)
    bag = m.sample_init_states_packed(2000, np.random.RandomState(0))
    g = oracle.GsPOMCP(om, 0, 0, bag, seed=3, ucb_c=30.0)
    g.search(20000, w0=32, growth=2, wmax=1024)
    assert g.greedy(0) == 0                                    # listen at the uniform belief
    for move in range(2):
        g.update(0, 1, 0, move=move, need=2000)                # two roars from door 2 (state 1)
        g.search(20000, move=move + 1, w0=32, growth=2, wmax=1024)
    # state 1 pays +10 for action 1 (tiger_model.py:118-121); posterior p(state 1) ~ 0.97
    assert g.greedy(2) == 1
