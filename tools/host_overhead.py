"""Host-side cost of the C-ABI calls of one move (wall clock, 1-simulation searches so that the kernel is ~35 us)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np, torch
from helpers import golden_tables, make_planner, random_world_and_bag
t = golden_tables("map-15-15"); actual, bag, _ = random_world_and_bag(t, 1)
pl = make_planner(t, seed=1, gen_w0=31250, gen_growth=4, gen_wmax=500000, node_capacity=1 << 22)
pl.set_world(actual, 0); pl.set_root_particles(bag, int(t["start_cell"]))
stream = torch.cuda.current_stream().cuda_stream
def timeit(name, fn, n=300):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print("%-40s %.1f us" % (name, dt * 1e6))
timeit("search(1) async", lambda i: pl.search(1, move=i, stream=stream))
timeit("search(1) + root_stats", lambda i: (pl.search(1, move=i, stream=stream), pl.root_stats()))
timeit("root_stats only", lambda i: pl.root_stats())
timeit("advance(check 5)", lambda i: pl.advance(5, 0, 0, move=i))
timeit("root_particles(fetch=False)", lambda i: pl.root_particles(fetch=False))
timeit("counters", lambda i: pl.counters())
pl.set_root_particles(bag, int(t["start_cell"]))
timeit("search(1M) + root_stats", lambda i: (pl.reset_tree(stream), pl.search(1000000, move=i, stream=stream), pl.root_stats()), n=30)
c = pl.counters(); print("device ms of last search", c["last_search_ns"] * 1e-6)
