"""Per-move search time inside an episode (re-rooted, deep trees), RockSample(15,15), 1e6 simulations/move, and the
per-generation phase log of the last search.  usage: episode_steplog.py [moves] [w0 growth wmax] [n_sims]"""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np, oracle
from helpers import golden_tables, make_planner, oracle_model
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 25
sched = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (32768, 4, 1 << 19)
n_sims = int(sys.argv[5]) if len(sys.argv) > 5 else 1000000
t = golden_tables("map-15-15"); om = oracle_model(t); K = om.n_rocks
pl = make_planner(t, seed=7, gen_w0=sched[0], gen_growth=sched[1], gen_wmax=sched[2], node_capacity=1 << 24)
rng = np.random.default_rng(1000)
actual = int(rng.integers(0, 1 << K))
bag = (rng.integers(0, 1 << K, size=2000).astype(np.uint32) << np.uint32(8)) | np.uint32(om.start_cell)
pl.set_world(actual, 0); pl.set_root_particles(bag, om.start_cell)
cell, rocks, sampled = om.start_cell, int(bag[0] >> 8), 0
times = []
for mv in range(moves):
    c0 = pl.counters()
    pl.search(n_sims, move=mv); st = pl.root_stats()
    c1 = pl.counters()
    times.append(c1["last_search_ns"] * 1e-6)
    if mv > 0 and times[-1] >= max(times[1:]):
        slow = (mv, pl.steplog().copy(), c1["generations"] - c0["generations"])
    print("move %d: %.3f ms, generations %d, tree levels/sim %.2f, plans %d, nodes %d" % (
        mv, times[-1], c1["generations"] - c0["generations"], (c1["levels"] - c0["levels"]) / n_sims,
        c1["alloc_nodes"] - c0["alloc_nodes"], c1["nodes"]), end="")
    mean = np.where(st["n"] > 0, st["qfix"] / 2.0 ** 24 / np.maximum(st["n"], 1), -1e30)
    a = int(np.argmax(mean))
    print(", root visits before %d, action %d (n %d)" % (int(st["n"].sum()) - n_sims, a, int(st["n"][a])))
    o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
    cell, rocks = o.next_cell, o.next_rocks
    if o.terminal or not o.legal: break
    if a == 4: sampled |= 1 << int(t["cell"][cell])
    if pl.advance(a, o.obs_good, sampled, move=mv)["status"] == 2: break
print("schedule", sched, "n_sims", n_sims, "median ms/move (after move 0): %.3f" % float(np.median(times[1:] or times)))
print("phase log of the slowest move after move 0: move %d" % slow[0])
L = slow[1]
for g in range(min(120, slow[2])):
    print("gen %d: start %.1f  descend+rollout %.1f  backup %.1f  flush+barrier %.1f us" % ((g,) + tuple(L[g, k] / 1e3 for k in range(4))))
print("phase totals us [start, descend+rollout, backup, flush+barrier]:", [round(float(x) / 1e3, 1) for x in L[126, :4]])
if L[127].any():
    tot = float(L[127, :5].sum())
    print("section shares (clock64 over all threads): plan %.3f  choose+step %.3f  child %.3f  rollout %.3f  backup %.3f" % tuple(L[127, k] / tot for k in range(5)))
    print("scout log of block 0, generation 0, per queued hot node (cycles): [allocation, cache fill + children + queue, 100 + warp, k]")
    for i in range(100):
        if L[i, 4] or L[i, 5]: print(i, L[i, 4:8].tolist())
