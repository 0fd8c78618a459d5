import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
from helpers import golden_tables, make_planner, random_world_and_bag
t = golden_tables("map-15-15"); actual, bag, _ = random_world_and_bag(t, 41)
w0, g, wmax = [int(x) for x in sys.argv[1:4]]
pl = make_planner(t, seed=1, gen_w0=w0, gen_growth=g, gen_wmax=wmax, node_capacity=1 << 22)
pl.set_world(actual, 0)
for rep in range(1):
    pl.set_root_particles(bag, int(t["start_cell"])); pl.search(1000000, move=rep); pl.root_stats()
c = pl.counters(); print(c["last_search_ns"] * 1e-6, c["level_steps"], c["alloc_nodes"])
L = pl.steplog()
for i in range(min(63, c["level_steps"])):
    print(i, "wait %.1f alloc %.1f wait2 %.1f expand %.1f us  nwl %d" % (L[i, 0] / 1e3, L[i, 1] / 1e3, L[i, 2] / 1e3, L[i, 3] / 1e3, L[i, 4]),
          " tsec Mcyc [load choice step child leaf sync arrive]:", [round(float(x) / 1e6, 1) for x in L[64 + i, :7]])

tot = float(L[127, :7].sum()) or 1.0
print("tsec share [load, choice, step, child, leaf-or-descend, syncwarp, arrive]:", [round(float(x) / tot, 3) for x in L[127, :7]], "total Gcycles", tot / 1e9)
print("phase totals us [gen start, barrier wait (working), allocate + barrier, expand, rollout + backup + flush]:", [round(float(x) / 1e3, 1) for x in L[126, :5]])
print("final phase Gcycles over all threads: rollouts %.2f backup %.2f (expand sections total %.2f)" % (L[127, 7] / 1e9, L[125, 0] / 1e9, L[127, :7].sum() / 1e9))
