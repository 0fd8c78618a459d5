"""Shared helpers of the test-suite: golden model tables -> oracle model / CUDA planner pairs."""
import os

import numpy as np

import oracle  # oracle/oracle.py (test infrastructure)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MAPS = ("map-7-8", "map-11-11", "map-15-15")


def golden_tables(tag):
    z = np.load(os.path.join(GOLDEN, "rocksample_%s.npz" % tag))
    return {k: z[k] for k in z.files}


def oracle_model(t):
    m = oracle.Model(int(t["rows"]), int(t["cols"]), t["cell"], int(t["start_cell"]), t["thr53"], t["rewards"])
    m.set_prob(t["prob"])
    return m


def make_planner(t, **kw):
    from pomdpy_b200 import native
    pl = native.Planner(**kw)
    pl.set_model_rocksample(int(t["rows"]), int(t["cols"]), t["cell"], int(t["start_cell"]), t["thr53"], t["rewards"])
    return pl


def random_world_and_bag(t, seed, n_bag=2000):
    rng = np.random.default_rng(seed)
    K = int(t["cell"].max()) + 1
    actual = int(rng.integers(0, 1 << K))
    bag = (rng.integers(0, 1 << K, size=n_bag).astype(np.uint32) << np.uint32(8)) | np.uint32(int(t["start_cell"]))
    return actual, bag.astype(np.uint32), rng
