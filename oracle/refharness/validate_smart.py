"""Preferred-action heuristic: pin the oracle's per-node functions to the EXECUTED reference
(examples/rock_sample/rock_position_history.py:96-137 create_child, :170-234 generate_smart_actions) on random
linear histories, and write tests/golden/smart_actions.json.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

On a linear chain the reference's shared rock-data list cannot leak between siblings, so the newest
node of the chain is exactly the clean per-node semantics."""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
import ref_env  # noqa: E402
from validate import tables_from_reference, oracle_model  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "golden")


def main():
    ref = ref_env.import_reference()
    ref["model"].pp = lambda *_a, **_k: None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from examples.rock_sample.rock_position_history import PositionAndRockData, RockData
        from examples.rock_sample.rock_observation import RockObservation
    rng = np.random.default_rng(99)
    out, bad, total = dict(generator="oracle/refharness/validate_smart.py", chains=[]), 0, 0
    for mp in ("map-7-8.txt", "map-11-11.txt", "map-15-15.txt"):
        m = ref_env.make_rock_model(ref, mp, preferred_actions=True)
        t = tables_from_reference(ref, m)
        om = oracle_model(t)
        om.set_prob(t["prob"])
        K, A = m.n_rocks, 5 + m.n_rocks
        for chain in range(60):
            data = PositionAndRockData(m, m.start_position.copy(), [RockData() for _ in range(K)], None)
            st = oracle.rockstats_init(om)
            cell = t["start_cell"]
            steps = []
            for step in range(14):
                legal = m.legal_actions_at(data.grid_position) if hasattr(m, "legal_actions_at") else \
                    data.generate_legal_actions()
                smart = data.generate_smart_actions()
                # mostly follow the heuristic, sometimes any legal action
                pool = [a for a in smart if a in legal] or legal
                a = int(pool[rng.integers(len(pool))]) if rng.random() < 0.8 else int(legal[rng.integers(len(legal))])
                good = bool(rng.integers(2)) if a >= 5 else False
                obs = RockObservation(good, False) if a >= 4 else RockObservation()
                data = data.create_child(ref["RockAction"](a), obs)
                st = oracle.rockstats_child(om, st, cell, a, good)
                cell = data.grid_position.i * m.n_cols + data.grid_position.j
                want_mask = sum(1 << x for x in data.generate_smart_actions())
                want = ([int(r.goodness_number) for r in data.all_rock_data], [float(r.chance_good) for r in data.all_rock_data])
                got = ([st.good[k] for k in range(K)], [st.chance[k] for k in range(K)])
                got_mask = oracle.smart_mask(om, st, cell)
                total += 1
                if want != got or want_mask != got_mask:
                    bad += 1
                    if bad < 4:
                        print("MISMATCH", mp, chain, step, a, good, want, got, want_mask, got_mask)
                steps.append(dict(action=a, obs_good=int(good), cell=int(cell), good=want[0], chance=want[1], mask=int(want_mask)))
            out["chains"].append(dict(map=mp.replace(".txt", ""), steps=steps))
    with open(os.path.join(GOLDEN, "smart_actions.json"), "w") as f:
        json.dump(out, f)
    print("%d / %d node states differ" % (bad, total))
    print("ALL OK" if bad == 0 else "FAILURES")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
