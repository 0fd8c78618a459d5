import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
from helpers import golden_tables, make_planner, random_world_and_bag
t = golden_tables("map-15-15"); actual, bag, _ = random_world_and_bag(t, 41)
w0, g, wmax = [int(x) for x in sys.argv[1:4]]
pl = make_planner(t, seed=1, gen_w0=w0, gen_growth=g, gen_wmax=wmax, node_capacity=1 << 22)
pl.set_world(actual, 0)
for rep in range(3):
    pl.set_root_particles(bag, int(t["start_cell"])); pl.search(1000000, move=rep); pl.root_stats()
c = pl.counters(); print(c["last_search_ns"] * 1e-6, c["level_steps"] // 3, c["alloc_nodes"] // 3)
L = pl.steplog()
for i in range(min(128, c["level_steps"] // 3)):
    print(i, "wait %.1f alloc %.1f wait2 %.1f expand %.1f us  nwl %d" % (L[i, 0] / 1e3, L[i, 1] / 1e3, L[i, 2] / 1e3, L[i, 3] / 1e3, L[i, 4]))

print("tsec cycles [choice, step+path, child, leaf-or-descend, syncwarp, arrive, load-sim, load-node]:", L[127].tolist())
