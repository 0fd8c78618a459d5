"""A handful of 1e6-simulation searches at bench.py's schedule (profiling target for ncu)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
from helpers import golden_tables, make_planner, random_world_and_bag
t = golden_tables("map-15-15"); actual, bag, _ = random_world_and_bag(t, 41)
pl = make_planner(t, seed=1, gen_w0=4096, gen_growth=4, gen_wmax=1 << 19, node_capacity=1 << 22)
pl.set_world(actual, 0)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    pl.set_root_particles(bag, int(t["start_cell"])); pl.search(1000000, move=rep); pl.root_stats()
print(pl.counters()["last_search_ns"] * 1e-6, "ms")
