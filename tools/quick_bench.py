import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
from helpers import golden_tables, make_planner, random_world_and_bag
tag = sys.argv[1] if len(sys.argv) > 1 else "map-15-15"
n_sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
t = golden_tables(tag)
actual, bag, _ = random_world_and_bag(t, 41)
for (w0, g, wmax) in [(1024, 2, 1 << 17), (1024, 2, 1 << 18), (4096, 4, 1 << 19), (16384, 4, 1 << 19), (32768, 4, 1 << 19), (1 << 18, 1, 1 << 18), (1 << 20, 1, 1 << 20)]:
    pl = make_planner(t, seed=1, gen_w0=w0, gen_growth=g, gen_wmax=wmax, node_capacity=1 << 22)
    pl.set_world(actual, 0)
    ts = []
    for rep in range(4):
        pl.set_root_particles(bag, int(t["start_cell"]))
        t0 = time.perf_counter(); pl.search(n_sims, move=rep); r = pl.root_stats(); t1 = time.perf_counter()
        c = pl.counters()
        ts.append((t1 - t0, c["last_search_ns"] * 1e-9))
    print(json.dumps(dict(sched=(w0, g, wmax), wall_ms=[round(x[0] * 1e3, 3) for x in ts], dev_ms=[round(x[1] * 1e3, 3) for x in ts],
                          sims_per_s=n_sims / min(x[1] for x in ts), levels=c["levels"], rollout=c["rollout_steps"], gens=c["generations"],
                          descend_ms=round(c["phase_descend_ns"]*1e-6,3), blocks=c["coop_blocks"], nodes=c["nodes"])), flush=True)
    pl.close()
