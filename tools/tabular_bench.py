"""Timing of the tabular-model planner (not the headline benchmark): 1e6 simulations per search on Tiger and on
random POMDPs, bench.py's generation schedule."""
import json
import sys
import time

sys.path.insert(0, '/root/repo')
import numpy as np
from pomdpy_b200 import native
from pomdpy_b200.tabular import random_pomdp, tiger

n_sims = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m256 = random_pomdp(256, 16, 16, seed=1)
for name, m, kw in (("tiger", tiger(), dict(ucb_coefficient=30.0)),
                    ("rand-256-16-16", m256, dict(ucb_coefficient=10.0)),
                    ("rand-256-16-16 hashed child table", m256, dict(ucb_coefficient=10.0, hashed_children=True)),
                    ("rand-256-8-64 (hashed)", random_pomdp(256, 8, 64, seed=3), dict(ucb_coefficient=10.0)),
                    ("rand-64-4-1024 (hashed)", random_pomdp(64, 4, 1024, seed=4), dict(ucb_coefficient=10.0)),
                    ("rand-4096-8-8", random_pomdp(4096, 8, 8, seed=2, sparsity=0.98), dict(ucb_coefficient=10.0))):
    pl = native.Planner(seed=1, gen_w0=4096, gen_growth=4, gen_wmax=1 << 19, node_capacity=1 << 22, **kw)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    bag = m.sample_init_states_packed(2000, np.random.RandomState(0))
    ts = []
    for rep in range(4):
        pl.set_root_particles(bag)
        t0 = time.perf_counter(); pl.search(n_sims, move=rep); pl.root_stats(); t1 = time.perf_counter()
        c = pl.counters()
        ts.append((t1 - t0, c["last_search_ns"] * 1e-9))
    print(json.dumps(dict(model=name, n_sims=n_sims, dev_ms=[round(x[1] * 1e3, 3) for x in ts],
                          sims_per_s=n_sims / min(x[1] for x in ts), levels=c["levels"], rollout=c["rollout_steps"],
                          nodes=c["nodes"], blocks=c["coop_blocks"])), flush=True)
    pl.close()
