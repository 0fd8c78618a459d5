"""Return vs generation schedule on the GPU (RockSample maps), env stepped on the host with the oracle's step.
PREF=1: --preferred_actions (in-tree smart actions), PREF=2: also in the rollouts."""
import os, sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np, oracle
from helpers import golden_tables, make_planner, oracle_model
tag, n_sims, neps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
scheds = [tuple(int(x) for x in s.split(',')) for s in sys.argv[4:]]
t = golden_tables(tag); om = oracle_model(t); K = om.n_rocks
for (w0, g, wmax) in scheds:
    pl = make_planner(t, seed=7, gen_w0=w0, gen_growth=g, gen_wmax=wmax, node_capacity=min(1 << 24, max(1 << 16, n_sims * 40)))
    pref = int(os.environ.get("PREF", "0"))
    if pref: pl.set_preferred_actions(True, t["prob"], rollouts=pref == 2)
    rets, steps, t0, tsearch, nsearch, nchecks = [], [], time.time(), 0.0, 0, 0
    for ep in range(neps):
        rng = np.random.default_rng(1000 + ep)
        actual = int(rng.integers(0, 1 << K))
        bag = (rng.integers(0, 1 << K, size=2000).astype(np.uint32) << np.uint32(8)) | np.uint32(om.start_cell)
        pl.set_world(actual, 0); pl.set_root_particles(bag, om.start_cell)
        cell, rocks, sampled, ret, disc = om.start_cell, int(bag[rng.integers(0, 2000)] >> 8), 0, 0.0, 1.0
        for mv in range(200):
            ts = time.perf_counter(); pl.search(n_sims, move=ep * 256 + mv); st = pl.root_stats(); tsearch += time.perf_counter() - ts; nsearch += 1
            mean = np.where(st["n"] > 0, st["qfix"] / 2.0 ** 24 / np.maximum(st["n"], 1), 0.0)
            legal = np.array([(st["legal"] >> a) & 1 for a in range(om.n_actions)], dtype=bool)
            mean[~legal] = -1e30
            best = np.flatnonzero(mean == mean.max()); a = int(best[rng.integers(0, len(best))])
            o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
            nchecks += a > 4
            ret += disc * o.reward; disc *= 0.95; cell, rocks = o.next_cell, o.next_rocks
            if o.terminal or not o.legal: break
            if a == 4: sampled |= 1 << int(t["cell"][cell])
            if pl.advance(a, o.obs_good, sampled, move=ep * 256 + mv)["status"] == 2: break
        rets.append(ret); steps.append(mv + 1)
    r = np.array(rets)
    print(json.dumps(dict(map=tag, preferred=int(os.environ.get("PREF", "0")), n_sims=n_sims, sched=(w0, g, wmax), episodes=neps, ret=round(float(r.mean()), 3), stderr=round(float(r.std() / np.sqrt(neps)), 3),
                          steps=float(np.mean(steps)), capped=float(np.mean(np.array(steps) >= 200)), check_moves=round(nchecks / nsearch, 3), search_ms=round(1e3 * tsearch / nsearch, 3), wall_s=round(time.time() - t0, 1))), flush=True)
    pl.close()
