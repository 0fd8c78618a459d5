"""Per-move comparison CUDA vs oracle for one preferred-action case (first divergence, counters)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np, oracle
from helpers import golden_tables, make_planner, oracle_model, random_world_and_bag
tag, seed, n_sims, sched = "map-15-15", 55, 20000, (64, 4, 8192)
for rollouts in (False, True):
    w0, growth, wmax = sched
    t = golden_tables(tag); actual, bag, rng = random_world_and_bag(t, seed); om = oracle_model(t)
    g = oracle.GsPOMCP(om, actual, 0, bag, seed=seed); g.set_preferred(True, rollouts=rollouts)
    pl = make_planner(t, seed=seed, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 17)
    pl.set_preferred_actions(True, t["prob"], rollouts=rollouts)
    pl.set_world(actual, 0); pl.set_root_particles(bag, int(t["start_cell"]))
    cell, rocks, sampled = int(t["start_cell"]), int(bag[0] >> 8), 0
    for mv in range(10):
        for part, ns in enumerate((64, 256, 1024, 4096, n_sims - 5440)):
            pass
        g.search(n_sims, move=mv, w0=w0, growth=growth, wmax=wmax); pl.search(n_sims, move=mv)
        a, b = g.root_stats(), pl.root_stats()
        gc, pc = g.counters(), pl.counters()
        same = np.array_equal(a["n"], b["n"]) and np.array_equal(a["qfix"], b["qfix"])
        print("rollouts", rollouts, "move", mv, "same" if same else "DIFF", {k: (gc[k], pc[k]) for k in ("levels", "rollout_steps", "created", "nodes", "alloc_nodes")})
        if not same:
            print(" oracle n", a["n"].tolist()); print(" cuda   n", b["n"].tolist())
            print(" oracle q", a["qfix"].tolist()); print(" cuda   q", b["qfix"].tolist())
            break
        act = g.greedy(move=mv, stats=a)
        o = om.step(cell, rocks, act, int(rng.integers(0, 1 << 53)), actual, sampled)
        cell, rocks = o.next_cell, o.next_rocks
        if o.terminal or not o.legal: break
        if act == 4: sampled |= 1 << int(t["cell"][cell])
        u, v = g.update(act, o.obs_good, sampled, move=mv), pl.advance(act, o.obs_good, sampled, move=mv)
        print("   update", u, v)
    pl.close()
