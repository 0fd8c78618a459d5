"""The device against the EXECUTED reference, bit for bit: the reference-exact mode (the reference's sequential POMCP
with all its quirks, one thread) replays the episodes that oracle/refharness/validate.py recorded from the executed
reference with injected random draws (tests/golden/pomcp_ref_episodes.json): root visit counts, total and mean Q
(float64, exact), legal mask, chosen action, the real step, the number of draws consumed -- move after move,
belief updates included."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_tables

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def test_device_reproduces_executed_reference_episodes():
    from pomdpy_b200 import native
    with open(os.path.join(GOLDEN, "pomcp_ref_episodes.json")) as f:
        eps = json.load(f)["episodes"]
    moves = 0
    for ep in eps:
        t = golden_tables(ep["map"].replace(".txt", ""))
        kw = ep["overrides"]
        rx = native.ReferenceExact(int(t["rows"]), int(t["cols"]), t["cell"], int(t["start_cell"]), t["thr53"], t["rewards"],
                                   ep["actual"], np.array(ep["bag"], dtype=np.uint32), max_depth=kw.get("max_depth", 100),
                                   max_particle_count=kw.get("max_particle_count", 2000),
                                   ucb_c=kw.get("ucb_coefficient", 3.0), discount=kw.get("discount", 0.95),
                                   seed=ep["stream_seed"])
        particle = rx.sample_root_particle()
        sampled = 0
        for rec in ep["trace"]:
            s = rx.search(ep["n_sims"])
            assert s["action"] == rec["action"]
            assert s["n"].tolist() == rec["stats"]["n"] and s["N"] == rec["stats"]["N"]
            assert s["total"].tolist() == rec["stats"]["total"]          # float64, exact
            assert s["mean"].tolist() == rec["stats"]["mean"]
            assert s["legal"] == rec["stats"]["legal"]
            o = rx.env_step(particle, s["action"])
            particle = o["particle"]
            assert (o["obs_good"], o["reward"], o["terminal"], o["legal"], o["next_cell"], o["next_rocks"]) == (
                rec["obs_good"], rec["reward"], rec["terminal"], rec["legal"], rec["next_cell"], rec["next_rocks"])
            assert o["draws"] == rec["draws"]
            if not o["terminal"] or not o["legal"]:
                if s["action"] == 4:
                    sampled |= 1 << int(t["cell"][o["next_cell"]])
                u = rx.update(s["action"], o["obs_good"], sampled)
                assert u["draws"] == rec["draws_after_update"] and u["disable_tree"] == rec["disable_tree"]
            moves += 1
        rx.close()
    assert moves >= 40
