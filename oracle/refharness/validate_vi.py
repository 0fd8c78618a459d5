"""Pin the VI oracle to the EXECUTED reference (Tiger, horizons 1-3) and write tests/golden/vi_tiger.json.
TEST INFRASTRUCTURE ONLY; needs /root/reference.  Horizon 3 takes ~20 s in the reference (LP prune)."""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_env  # noqa: E402
import vi_oracle  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "golden")


def run_reference(horizon, discount=0.95):
    ref = ref_env.import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pomdpy.solvers import ValueIteration
        from examples.tiger import TigerModel

    class FakeAgent:
        pass
    ref["model"].pp = lambda *_a, **_k: None
    args = dict(env="Tiger", solver="ValueIteration", seed=1, discount=discount, n_epochs=1, max_steps=10,
                planning_horizon=horizon, save=False, use_tf=False)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = TigerModel(args)
        a = FakeAgent()
        a.model = m
        from pomdpy.pomdp.history import Histories
        a.histories = Histories()
        vi = ValueIteration(a)
        T, O, R = m.get_transition_matrix(), m.get_observation_matrix(), m.get_reward_matrix()
        vi.value_iteration(T, O, R, horizon)
    vecs = np.array([av.v for av in vi.gamma], dtype=np.float64)
    acts = np.array([av.action for av in vi.gamma])
    return T, O, R, vecs, acts


def main():
    grid = np.stack([np.linspace(0, 1, 201), 1 - np.linspace(0, 1, 201)], axis=1)
    out = dict(generator="oracle/refharness/validate_vi.py", discount=0.95, grid_points=201, horizons={})
    ok = True
    for h in (1, 2, 3):
        T, O, R, rv, ra = run_reference(h)
        ov, oa = vi_oracle.value_iteration(T, O, R, 0.95, h)
        env_r, env_o = vi_oracle.envelope(rv, grid), vi_oracle.envelope(ov, grid)
        cr, car = vi_oracle.canonical_set(rv, ra)
        co, cao = vi_oracle.canonical_set(ov, oa)
        same_set = cr.shape == co.shape and np.allclose(cr, co, rtol=1e-12, atol=1e-12) and np.array_equal(car, cao)
        print("horizon %d: reference %d vectors (canonical %d), oracle %d; envelope max diff %.3g; canonical sets equal: %s"
              % (h, len(rv), len(cr), len(ov), np.abs(env_r - env_o).max(), same_set))
        ok &= np.abs(env_r - env_o).max() <= 1e-9 and same_set
        out["horizons"][str(h)] = dict(reference_raw_count=int(len(rv)), canonical=cr.tolist(),
                                       canonical_actions=[int(x) for x in car], envelope=env_r.tolist())
    out["T"], out["O"], out["R"] = T.tolist(), O.tolist(), R.tolist()
    with open(os.path.join(GOLDEN, "vi_tiger.json"), "w") as f:
        json.dump(out, f)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
