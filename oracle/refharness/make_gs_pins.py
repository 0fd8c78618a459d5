"""Regression pins of the generation-synchronous oracle (the specification the CUDA kernels are held to): root statistics,
counters and belief checksums of a dozen fixed searches, written to tests/golden/gs_pins.json.  Not a reference
comparison (that is validate.py for the "ref" mode) -- a guard against unintended changes of the specification itself;
regenerate deliberately, with the change described in DESIGN.md.
    python oracle/refharness/make_gs_pins.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import golden_tables, oracle_model, random_world_and_bag  # noqa: E402
from pomdpy_b200.tabular import random_pomdp, tiger  # noqa: E402

CASES = [
    dict(kind="rs", tag="map-7-8", seed=1, n_sims=2000, sched=(8, 2, 128), pref=0),
    dict(kind="rs", tag="map-7-8", seed=2, n_sims=300, sched=(1, 1, 1), pref=0),
    dict(kind="rs", tag="map-11-11", seed=3, n_sims=5000, sched=(32, 4, 2048), pref=0),
    dict(kind="rs", tag="map-15-15", seed=4, n_sims=20000, sched=(64, 4, 8192), pref=0),
    dict(kind="rs", tag="map-7-8", seed=5, n_sims=1500, sched=(8, 2, 128), pref=1),
    dict(kind="rs", tag="map-15-15", seed=6, n_sims=6000, sched=(64, 4, 4096), pref=2),
    dict(kind="tab", model="tiger", seed=7, n_sims=20000, sched=(32, 2, 2048), ucb=30.0),
    dict(kind="tab", model="rand-40-6-8", seed=8, n_sims=10000, sched=(16, 4, 1024), ucb=10.0),
    dict(kind="tab", model="rand-300-32-3", seed=9, n_sims=10000, sched=(64, 4, 4096), ucb=8.0),
    # more than 32 observations: the observation map is unbounded (hashed child table on the device)
    dict(kind="tab", model="rand-48-4-64", seed=10, n_sims=12000, sched=(32, 4, 4096), ucb=10.0),
    dict(kind="tab", model="rand-24-3-300", seed=11, n_sims=8000, sched=(16, 4, 2048), ucb=6.0),
]


def run_case(c):
    w0, growth, wmax = c["sched"]
    if c["kind"] == "rs":
        t = golden_tables(c["tag"])
        actual, bag, _ = random_world_and_bag(t, c["seed"])
        g = oracle.GsPOMCP(oracle_model(t), actual, 0, bag, seed=c["seed"])
        if c["pref"]:
            g.set_preferred(True, rollouts=c["pref"] == 2)
    else:
        m = tiger() if c["model"] == "tiger" else random_pomdp(*[int(x) for x in c["model"].split("-")[1:]], seed=c["seed"])
        om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
        bag = m.sample_init_states_packed(1000, np.random.RandomState(c["seed"]))
        g = oracle.GsPOMCP(om, 0, 0, bag, seed=c["seed"], ucb_c=c["ucb"], max_depth=40)
    out = []
    for move in range(2):
        g.search(c["n_sims"], move=move, w0=w0, growth=growth, wmax=wmax)
        st = g.root_stats()
        a = g.greedy(move, st)
        cnt = g.counters()
        u = g.update(a, move & 1, 0, move=move, need=500)
        out.append(dict(n=st["n"].tolist(), qfix=st["qfix"].tolist(), legal=int(st["legal"]), action=a,
                        levels=cnt["levels"], rollout_steps=cnt["rollout_steps"], nodes=cnt["nodes"],
                        update=[u["status"], u["tries"], u["accepted"]],
                        bag_sum=int(g.bag().astype(np.uint64).sum()) if u["status"] != 2 else -1))
        if u["status"] == 2:
            break
    return out


def main():
    pins = [dict(case=c, moves=run_case(c)) for c in CASES]
    path = os.path.join(ROOT, "tests", "golden", "gs_pins.json")
    with open(path, "w") as f:
        json.dump(dict(generator="oracle/refharness/make_gs_pins.py", pins=pins), f, indent=0)
    print("wrote", path, len(pins), "cases")


if __name__ == "__main__":
    main()
