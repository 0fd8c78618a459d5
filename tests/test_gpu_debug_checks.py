"""The -DPOMCP_CHECKS build (device-side bounds and invariant checks; compute-sanitizer is closed on the GPU pool):
a multi-move episode with and without preferred actions, and a tabular model on the hashed child table, must trip no
check and still equal the oracle."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

SCRIPT = r'''
import sys
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np, oracle
from helpers import golden_tables, make_planner, oracle_model, random_world_and_bag
for tag, n_sims, sched, pref in (("map-15-15", 60000, (256, 4, 32768), False), ("map-7-8", 3000, (8, 2, 256), True),
                                 ("map-11-11", 20000, (1, 16, 65536), False)):
    t = golden_tables(tag); om = oracle_model(t)
    actual, bag, rng = random_world_and_bag(t, 77)
    g = oracle.GsPOMCP(om, actual, 0, bag, seed=3)
    pl = make_planner(t, seed=3, gen_w0=sched[0], gen_growth=sched[1], gen_wmax=sched[2], node_capacity=1 << 18)
    if pref:
        g.set_preferred(True); pl.set_preferred_actions(True, t["prob"])
    pl.set_world(actual, 0); pl.set_root_particles(bag, int(t["start_cell"]))
    cell, rocks, sampled = int(t["start_cell"]), int(bag[0] >> 8), 0
    for mv in range(5):
        g.search(n_sims, move=mv, w0=sched[0], growth=sched[1], wmax=sched[2]); pl.search(n_sims, move=mv)
        a, b = g.root_stats(), pl.root_stats()
        assert np.array_equal(a["n"], b["n"]) and np.array_equal(a["qfix"], b["qfix"])
        act = g.greedy(move=mv, stats=b)
        o = om.step(cell, rocks, act, int(rng.integers(0, 1 << 53)), actual, sampled)
        cell, rocks = o.next_cell, o.next_rocks
        if o.terminal or not o.legal: break
        if act == 4: sampled |= 1 << int(t["cell"][cell])
        assert g.update(act, o.obs_good, sampled, move=mv) == pl.advance(act, o.obs_good, sampled, move=mv)
    c = pl.counters()
    assert c["check_failed"] == 0, c["check_failed"]
    pl.close()
# a tabular model with more than 32 observations: the hashed child table under the same checks
from pomdpy_b200 import native
from pomdpy_b200.tabular import random_pomdp
m = random_pomdp(48, 4, 64, seed=10)
om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
bag = m.sample_init_states_packed(2000, np.random.RandomState(1))
g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=30, ucb_c=10.0, seed=3)
pl = native.Planner(max_depth=30, ucb_coefficient=10.0, seed=3, gen_w0=32, gen_growth=4, gen_wmax=8192,
                    node_capacity=1 << 17, max_particle_count=1000)
pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state); pl.set_root_particles(bag)
for mv in range(3):
    g.search(20000, move=mv, w0=32, growth=4, wmax=8192); pl.search(20000, move=mv)
    a, b = g.root_stats(), pl.root_stats()
    assert np.array_equal(a["n"], b["n"]) and np.array_equal(a["qfix"], b["qfix"])
    act = g.greedy(move=mv, stats=b)
    u, v = g.update(act, 7 * mv + 1, 0, move=mv, need=1000), pl.advance(act, 7 * mv + 1, 0, move=mv)
    assert (u["status"], u["tries"], u["accepted"]) == (v["status"], v["tries"], v["accepted"])
    if u["status"] == 2: break
c = pl.counters()
assert c["check_failed"] == 0, c["check_failed"]
pl.close()
print("CHECKS-OK")
'''


def test_checks_build_trips_nothing():
    lib = os.path.join(ROOT, "pomdpy_b200", "libpomcp_b200_checks.so")
    if not os.path.exists(lib):
        from pomdpy_b200 import build
        build.build(force=True, checks=True)
    env = dict(os.environ, POMCP_B200_LIB=lib)
    res = subprocess.run([sys.executable, "-c", "ROOT = %r\n" % ROOT + SCRIPT], env=env, capture_output=True, text=True,
                         timeout=500)
    if res.returncode != 0 or "CHECKS-OK" not in res.stdout:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "checks_child.log"), "a") as f:
            f.write("==== rc %d\n%s\n%s\n" % (res.returncode, res.stdout[-3000:], res.stderr[-6000:]))
    assert res.returncode == 0 and "CHECKS-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
