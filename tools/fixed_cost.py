"""Kernel time of tiny searches: the fixed cost of a launch (prologue, barriers, epilogue)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
from helpers import golden_tables, make_planner, random_world_and_bag
t = golden_tables("map-15-15"); actual, bag, _ = random_world_and_bag(t, 41)
pl = make_planner(t, seed=1, gen_w0=4096, gen_growth=4, gen_wmax=1 << 19, node_capacity=1 << 22)
pl.set_world(actual, 0)
for n in (1, 256, 4096, 20480, 151552, 151552 * 2):
    ts = []
    for rep in range(4):
        pl.set_root_particles(bag, int(t["start_cell"])); pl.search(n, move=rep); pl.root_stats()
        ts.append(pl.counters()["last_search_ns"] * 1e-3)
    print(n, "sims:", [round(x, 1) for x in ts], "us (device globaltimer inside the kernel)")
