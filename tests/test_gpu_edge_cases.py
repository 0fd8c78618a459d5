"""Edge cases of the search / belief update on the GPU, each against the oracle: obstacles, non-square and
maximum-size grids, custom rewards, degenerate beliefs and budgets, the capacity and timeout guards."""
import os

import numpy as np
import pytest

import oracle
from helpers import make_planner

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

OBSTACLE_MAP = """6 9
.o..X...G
..X...o.G
S.....X.G
.Xo.....G
....o.X.G
o.X.....G
"""


def _tables_from_map(tmp_path, text, **cfg):
    from pomdpy_b200.rock_sample import RockModel
    f = tmp_path / "custom-map.txt"
    f.write_text(text)
    m = RockModel(dict(map_file=str(f), discount=0.95), rock_config=cfg or None)
    s = m.device_spec()
    return dict(rows=s["rows"], cols=s["cols"], cell=s["cell"], start_cell=s["start_cell"], thr53=s["thr53"],
                prob=s["prob"], rewards=s["rewards"])


def _pair(t, wseed, n_bag=2000, root_cell=None, **kw):
    rng = np.random.default_rng(wseed)
    K = int(t["cell"].max()) + 1
    actual = int(rng.integers(0, 1 << K))
    rc = int(t["start_cell"]) if root_cell is None else root_cell
    bag = ((rng.integers(0, 1 << K, size=n_bag, dtype=np.uint64).astype(np.uint32) << np.uint32(8)) | np.uint32(rc)).astype(np.uint32)
    om = oracle.Model(int(t["rows"]), int(t["cols"]), t["cell"], int(t["start_cell"]), t["thr53"], t["rewards"])
    om.set_prob(t["prob"])
    g = oracle.GsPOMCP(om, actual, 0, bag, root_cell=rc, max_depth=kw.get("max_depth", 100), seed=kw.get("seed", 0))
    pl = make_planner(t, **kw)
    pl.set_world(actual, 0)
    pl.set_root_particles(bag, rc)
    return om, g, pl, actual, bag, rng


def _same(g, pl):
    a, b = g.root_stats(), pl.root_stats()
    assert a["legal"] == b["legal"] and np.array_equal(a["n"], b["n"]) and np.array_equal(a["qfix"], b["qfix"])
    return b


def test_obstacles_nonsquare_custom_rewards(tmp_path):
    t = _tables_from_map(tmp_path, OBSTACLE_MAP, good_rock_reward=7.5, bad_rock_penalty=3.25, exit_reward=12.0,
                         illegal_move_penalty=50.0, half_efficiency_distance=5.0)
    kw = dict(seed=3, gen_w0=16, gen_growth=2, gen_wmax=512, node_capacity=1 << 15)
    om, g, pl, actual, bag, rng = _pair(t, 3, **kw)
    # legal sets around obstacles: the device derives them from the cell codes on its own
    cell, rocks, sampled = int(t["start_cell"]), int(bag[0] >> 8), 0
    for mv in range(6):
        g.search(3000, move=mv, w0=16, growth=2, wmax=512)
        pl.search(3000, move=mv)
        st = _same(g, pl)
        a = g.greedy(move=mv, stats=st)
        o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
        assert o.legal and o.reward in (0.0, 7.5, -3.25, 12.0)
        cell, rocks = o.next_cell, o.next_rocks
        if o.terminal:
            break
        if a == 4:
            sampled |= 1 << int(t["cell"][cell])
        assert g.update(a, o.obs_good, sampled, move=mv) == pl.advance(a, o.obs_good, sampled, move=mv)
        assert np.array_equal(pl.root_particles()[0], g.bag())
    pl.close()


def test_maximum_size_grid_and_rock_count(tmp_path):
    """16 x 16 = 256 cells, 24 rocks (|A| = 29): the limits of the packed state and of lane = action."""
    rows = [["."] * 15 + ["G"] for _ in range(16)]
    rows[8][0] = "S"
    rng = np.random.default_rng(0)
    placed = 0
    while placed < 24:
        i, j = int(rng.integers(0, 16)), int(rng.integers(0, 15))
        if rows[i][j] == ".":
            rows[i][j] = "o"
            placed += 1
    t = _tables_from_map(tmp_path, "16 16\n" + "\n".join("".join(r) for r in rows) + "\n")
    kw = dict(seed=4, gen_w0=32, gen_growth=4, gen_wmax=2048, node_capacity=1 << 15)
    om, g, pl, actual, bag, _ = _pair(t, 4, **kw)
    assert om.n_actions == 29
    g.search(8000, move=0, w0=32, growth=4, wmax=2048)
    pl.search(8000, move=0)
    _same(g, pl)
    pl.close()


def test_degenerate_beliefs_and_budgets(golden_dir):
    from helpers import golden_tables
    t = golden_tables("map-7-8")
    # one particle; zero simulations; a root next to the goal column (terminal steps inside the tree)
    om, g, pl, actual, bag, _ = _pair(t, 5, n_bag=1, seed=5, gen_w0=4, gen_growth=2, gen_wmax=64, node_capacity=1 << 14)
    pl.search(0)
    assert int(pl.root_stats()["n"].sum()) == 0
    g.search(700, move=1, w0=4, growth=2, wmax=64)
    pl.search(700, move=1)
    _same(g, pl)
    pl.close()
    near_goal = 3 * 7 + 5                      # row 3, column 5: EAST exits
    om, g, pl, actual, bag, _ = _pair(t, 6, root_cell=near_goal, seed=6, gen_w0=8, gen_growth=2, gen_wmax=128,
                                      node_capacity=1 << 14)
    g.search(2000, move=0, w0=8, growth=2, wmax=128)
    pl.search(2000, move=0)
    st = _same(g, pl)
    assert st["n"][1] > 0                      # EAST was tried; its child never exists (terminal)
    assert pl.node_stats([(1, 0)]) is None and g.node_stats([(1, 0)]) is None
    pl.close()


def test_capacity_guard_drops_subtree_and_continues():
    from helpers import golden_tables
    t = golden_tables("map-7-8")
    om, g, pl, actual, bag, _ = _pair(t, 7, seed=7, gen_w0=8, gen_growth=2, gen_wmax=64, node_capacity=1200)
    assert pl.search(500, move=0) == 0
    n1 = pl.counters()["nodes"]
    assert pl.search(500, move=1) == 0 and pl.counters()["nodes"] >= n1
    rc = pl.search(500, move=2)                # 1200 - used < 502: subtree dropped, belief kept
    assert rc == 1
    st = pl.root_stats()
    assert int(st["n"].sum()) == 500 and pl.counters()["nodes"] <= 501
    from pomdpy_b200 import native
    with pytest.raises(native.PomcpError):
        pl.search(5000)                        # can never fit
    pl.close()


def test_timeout_guard_stops_between_generations():
    from helpers import golden_tables
    t = golden_tables("map-15-15")
    om, g, pl, actual, bag, _ = _pair(t, 8, seed=8, gen_w0=1, gen_growth=1, gen_wmax=1, node_capacity=1 << 18)
    pl.search(200000, move=0, timeout_s=0.05)  # strictly sequential generations: ~14k simulations/s
    st = pl.root_stats()
    assert 0 < st["sims_done"] < 200000 and int(st["n"].sum()) == st["sims_done"]
    pl.close()


@pytest.mark.parametrize("max_depth", [300, 700])
def test_long_horizons_beyond_256_steps(golden_dir, max_depth):
    """--max_depth above 255: rollout step counters, the discount table in shared memory and the Philox block index
    all run past 8 bits (the ABI accepts max_depth up to 4096).  A tabular model without terminal states rolls out
    the full horizon every time; RockSample(11,11) at the same horizon as the second case.  Bit-exact vs the oracle."""
    from helpers import golden_tables
    from pomdpy_b200 import native
    from pomdpy_b200.tabular import random_pomdp
    m = random_pomdp(32, 6, 8, seed=1)
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    bag = m.sample_init_states_packed(1000, np.random.RandomState(2))
    g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=max_depth, ucb_c=10.0, discount=0.995, seed=5)
    pl = native.Planner(max_depth=max_depth, ucb_coefficient=10.0, discount=0.995, seed=5, gen_w0=16, gen_growth=4,
                        gen_wmax=2048, node_capacity=1 << 16)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl.set_root_particles(bag)
    g.search(6000, w0=16, growth=4, wmax=2048)
    pl.search(6000)
    _same(g, pl)
    gc, pc = g.counters(), pl.counters()
    assert gc["rollout_steps"] == pc["rollout_steps"] and gc["levels"] == pc["levels"]
    assert pc["rollout_steps"] > 6000 * (max_depth - 10)            # no terminal states: (almost) full-length rollouts
    pl.close()
    t = golden_tables("map-11-11")
    kw = dict(seed=9, gen_w0=16, gen_growth=4, gen_wmax=2048, node_capacity=1 << 16, max_depth=max_depth)
    om, g, pl, actual, bag, rng = _pair(t, 9, **kw)
    g.search(5000, w0=16, growth=4, wmax=2048)
    pl.search(5000)
    _same(g, pl)
    assert g.counters()["rollout_steps"] == pl.counters()["rollout_steps"]
    pl.close()
