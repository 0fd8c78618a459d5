import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
from helpers import golden_tables, make_planner, random_world_and_bag
t = golden_tables("map-7-8"); actual, bag, _ = random_world_and_bag(t, 1)
for n_sims, sched in ((1, (1, 1, 1)), (1, (1, 1, 1)), (300, (1, 1, 1)), (600, (8, 2, 64)), (600, (8, 2, 64))):
    pl = make_planner(t, seed=101, gen_w0=sched[0], gen_growth=sched[1], gen_wmax=sched[2], node_capacity=1 << 16)
    pl.set_world(actual, 0); pl.set_root_particles(bag, int(t["start_cell"]))
    for rep in range(2):
        pl.search(n_sims, move=3 + rep); r = pl.root_stats(); c = pl.counters()
        print(n_sims, sched, rep, "n", int(r["n"].sum()), "done", r["sims_done"], {k: c[k] for k in ("levels", "rollout_steps", "created", "nodes", "generations", "level_steps", "check_failed")})
    pl.close()
