"""Level-step profile of searches inside an episode (re-rooted, deep trees), RockSample(15,15), 1e6 simulations/move."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np, oracle
from helpers import golden_tables, make_planner, oracle_model
t = golden_tables("map-15-15"); om = oracle_model(t); K = om.n_rocks
pl = make_planner(t, seed=7, gen_w0=32768, gen_growth=4, gen_wmax=1 << 19, node_capacity=1 << 24)
rng = np.random.default_rng(1000)
actual = int(rng.integers(0, 1 << K))
bag = (rng.integers(0, 1 << K, size=2000).astype(np.uint32) << np.uint32(8)) | np.uint32(om.start_cell)
pl.set_world(actual, 0); pl.set_root_particles(bag, om.start_cell)
cell, rocks, sampled = om.start_cell, int(bag[0] >> 8), 0
for mv in range(int(sys.argv[1]) if len(sys.argv) > 1 else 25):
    c0 = pl.counters()
    pl.search(1000000, move=mv); st = pl.root_stats()
    c1 = pl.counters()
    print("move %d: %.3f ms, level steps %d, tree levels/sim %.2f, nodes %d" % (
        mv, c1["last_search_ns"] * 1e-6, c1["level_steps"] - c0["level_steps"], (c1["levels"] - c0["levels"]) / 1e6, c1["nodes"]))
    mean = np.where(st["n"] > 0, st["qfix"] / 2.0 ** 24 / np.maximum(st["n"], 1), -1e30)
    a = int(np.argmax(mean))
    o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
    cell, rocks = o.next_cell, o.next_rocks
    if o.terminal or not o.legal: break
    if a == 4: sampled |= 1 << int(t["cell"][cell])
    if pl.advance(a, o.obs_good, sampled, move=mv)["status"] == 2: break
L = pl.steplog()
n = min(63, c1["level_steps"] - c0["level_steps"])
for i in range(n):
    print(i, "wait %.1f alloc %.1f wait2 %.1f expand %.1f us  nwl %d" % (L[i, 0] / 1e3, L[i, 1] / 1e3, L[i, 2] / 1e3, L[i, 3] / 1e3, L[i, 4]))
print("phase totals us [gen start, waits, allocate + barrier, expand, final]:", [round(float(x) / 1e3, 1) for x in L[126, :5]])
