"""Return of the SEQUENTIAL planners on the CPU as a function of the simulation budget (iso-quality evidence, VERDICT r1
item 5): the oracle's "ref" mode (the reference's algorithm, quirks on) and "gs" with generation width 1 (canonical
sequential POMCP), paired worlds (seeds 1000 + episode, as tools/gpu_quality.py), all host cores.
usage: [PLANNERS=ref,gs_w1] [WORKERS=n] cpu_quality_curve.py map episodes n_sims [n_sims ...]
(at 1e5 simulations a "ref" episode holds several GB of tree: lower WORKERS or leave it out)"""
import os, sys, time, json
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo/oracle')
from multiprocessing import Pool
import numpy as np
from quality import load, episode_gs, episode_ref

def run(a):
    kind, tag, n_sims, seed = a
    om, z = load(tag)
    t0 = time.time()
    r = episode_ref(om, z, seed, n_sims) if kind == "ref" else episode_gs(om, z, seed, n_sims, 1, 1, 1)
    return r[0], r[1], time.time() - t0

if __name__ == "__main__":
    tag, neps = sys.argv[1], int(sys.argv[2])
    with Pool(int(os.environ.get("WORKERS", "0")) or None) as pool:
        for n_sims in [int(x) for x in sys.argv[3:]]:
            for kind in os.environ.get("PLANNERS", "ref,gs_w1").split(","):
                R = pool.map(run, [(kind, tag, n_sims, 1000 + i) for i in range(neps)], chunksize=1)
                r = np.array([x[0] for x in R]); st = np.array([x[1] for x in R]); sec = np.array([x[2] for x in R])
                print(json.dumps(dict(map=tag, planner=kind, n_sims=n_sims, episodes=neps, ret=round(float(r.mean()), 3),
                                      stderr=round(float(r.std() / np.sqrt(neps)), 3), steps=float(st.mean()),
                                      ms_per_move_one_core=round(1e3 * float(sec.sum() / st.sum()), 2))), flush=True)
