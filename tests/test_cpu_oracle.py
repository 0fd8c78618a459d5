"""CPU suite: the oracle against the golden vectors of the executed reference, the host logic, and the C ABI
surface (no GPU needed)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import oracle
from helpers import GOLDEN, MAPS, ROOT, golden_tables, oracle_model


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    from pomdpy_b200 import philox
    for ctr, key, want in kats:
        assert tuple(int(v) for v in oracle.philox(ctr, key)) == want
        assert philox.philox4x32_10(ctr, key) == want
    # ... and for the 7-round variant the rollout stream of the "gs" planner uses (oracle/pomcp_oracle.c ORC_ROLLOUT_ROUNDS)
    kats7 = [((0, 0, 0, 0), (0, 0), (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)),
             ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
              (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a))]
    for ctr, key, want in kats7:
        assert tuple(int(v) for v in oracle.philox(ctr, key, rounds=7)) == want


@pytest.mark.parametrize("tag", MAPS)
def test_oracle_generate_step_and_legal_sets_golden(tag):
    """oracle step / legal masks == executed reference (rock_model.py:451-465, :218-247)."""
    t = golden_tables(tag)
    om = oracle_model(t)
    n_cells = int(t["rows"]) * int(t["cols"])
    assert [om.legal_mask(c) for c in range(n_cells)] == t["legal"].tolist()
    si, so = t["step_in"], t["step_out"]
    idx = np.arange(0, len(si), max(1, len(si) // 4000))
    for i in idx:
        c, rocks, a, u, actual, sampled = (int(x) for x in si[i])
        o = om.step(c, rocks, a, u, actual, sampled)
        got = (o.next_cell, o.next_rocks, o.obs_good, o.obs_empty, o.reward, o.terminal, o.legal, o.rewarded_sample)
        assert got == tuple(so[i].tolist()), (i, got, so[i])


def test_oracle_ref_mode_reproduces_reference_episodes():
    """The sequential 'ref' mode == the executed reference move by move (root statistics, chosen action, real
    step, random draws consumed), including the belief updates: tests/golden/pomcp_ref_episodes.json."""
    with open(os.path.join(GOLDEN, "pomcp_ref_episodes.json")) as f:
        eps = json.load(f)["episodes"]
    for ep in eps:
        tag = ep["map"].replace(".txt", "")
        t = golden_tables(tag)
        om = oracle_model(t)
        kw = ep["overrides"]
        o = oracle.RefPOMCP(om, ep["actual"], 0, np.array(ep["bag"], dtype=np.uint32),
                            max_depth=kw.get("max_depth", 100), max_particle_count=kw.get("max_particle_count", 2000),
                            ucb_c=kw.get("ucb_coefficient", 3.0), discount=kw.get("discount", 0.95),
                            seed=ep["stream_seed"])
        p = o.sample_root_particle()
        sampled = 0
        for rec in ep["trace"]:
            a = o.search(ep["n_sims"])
            s = o.root_stats()
            assert a == rec["action"]
            assert s["n"].tolist() == rec["stats"]["n"] and s["N"] == rec["stats"]["N"]
            assert s["total"].tolist() == rec["stats"]["total"] and s["mean"].tolist() == rec["stats"]["mean"]
            assert s["legal"] == rec["stats"]["legal"]
            p, so = o.env_step(p, a)
            cell, rocks = o.particle(p)
            assert (so.obs_good, so.reward, so.terminal, so.legal, cell, rocks) == (
                rec["obs_good"], rec["reward"], rec["terminal"], rec["legal"], rec["next_cell"], rec["next_rocks"])
            assert o.draws() == rec["draws"]
            if not so.terminal or not so.legal:
                if a == 4:
                    sampled |= 1 << int(t["cell"][cell])
                o.update(a, so.obs_good, sampled)
            assert o.draws() == rec["draws_after_update"] and o.disabled() == rec["disable_tree"]


def test_det_log_accuracy():
    import math
    for x in list(range(1, 2000)) + [10 ** 6, 2 ** 20, 2 ** 31 - 1, 2 ** 40 + 12345]:
        assert abs(oracle.det_log(x) - math.log(x)) <= 4e-16 * max(1.0, math.log(x))


def test_gs_width_one_is_sequential_and_invariants():
    """gs mode: visits add up, generation width only changes the interleaving, not the budget."""
    t = golden_tables("map-7-8")
    om = oracle_model(t)
    rng = np.random.default_rng(3)
    bag = (rng.integers(0, 256, size=500).astype(np.uint32) << np.uint32(8)) | np.uint32(int(t["start_cell"]))
    for sched in ((1, 1, 1), (8, 2, 64), (1000, 1, 1000)):
        g = oracle.GsPOMCP(om, 0b10110101, 0, bag, seed=5)
        g.search(1000, w0=sched[0], growth=sched[1], wmax=sched[2])
        s, c = g.root_stats(), g.counters()
        assert int(s["n"].sum()) == 1000 and c["created"] + 1 == c["nodes"]
        assert all(s["n"][a] == 0 for a in range(om.n_actions) if not (s["legal"] >> a) & 1)
    g1 = oracle.GsPOMCP(om, 0b10110101, 0, bag, seed=5)
    g1.search(1000, w0=1, growth=1, wmax=1)
    assert g1.counters()["alloc_nodes"] == 0 and g1.counters()["generations"] == 1000


def test_product_tables_match_reference_tables():
    """RockModel.device_spec() == the tables extracted from the executed reference."""
    from pomdpy_b200.rock_sample import GridPosition, RockModel
    for tag in MAPS:
        m = RockModel(dict(map_file=tag, discount=0.95))
        s, t = m.device_spec(), golden_tables(tag)
        assert np.array_equal(s["cell"], t["cell"]) and np.array_equal(s["thr53"], t["thr53"])
        assert s["start_cell"] == int(t["start_cell"]) and np.array_equal(s["rewards"], t["rewards"])
        legal = [sum(1 << a for a in m.legal_actions_at(GridPosition(c // m.n_cols, c % m.n_cols)))
                 for c in range(m.n_rows * m.n_cols)]
        assert legal == t["legal"].tolist()


def test_host_generate_step_matches_golden():
    """The Python RockModel.generate_step (the real environment step) on the golden grid."""
    from pomdpy_b200.rock_sample import GridPosition, RockModel, RockState
    t = golden_tables("map-7-8")
    m = RockModel(dict(map_file="map-7-8", discount=0.95))
    si, so = t["step_in"], t["step_out"]
    real = np.random.binomial
    try:
        for i in range(0, len(si), 3):
            c, rocks, a, u, actual, sampled = (int(x) for x in si[i])
            m.actual_rock_states = m.decode_rocks(actual)
            m.unique_rocks_sampled = [k for k in range(m.n_rocks) if (sampled >> k) & 1]
            thr = int(t["thr53"][c, a - 5]) if a >= 5 else 0
            np.random.binomial = lambda n, p, _u=u, _t=thr: 1 if _u <= _t else 0
            st = RockState(GridPosition(c // m.n_cols, c % m.n_cols), m.decode_rocks(rocks))
            res, legal = m.generate_step(st, a)
            got = (m.cell_index(res.next_state.position), m.encode_rocks(res.next_state.rock_states),
                   int(bool(res.observation.is_good)), int(bool(res.observation.is_empty)), float(res.reward),
                   int(bool(res.is_terminal)), int(bool(legal)))
            assert got == tuple(so[i][:7].tolist()), (i, got, so[i])
    finally:
        np.random.binomial = real


def test_c_abi_exports_every_declared_symbol():
    """libpomcp_b200.so loads without a GPU and exports everything include/pomcp_b200.h declares."""
    from pomdpy_b200 import native
    L = native.load()
    hdr = open(os.path.join(ROOT, "include", "pomcp_b200.h")).read()
    declared = set(re.findall(r"\b(pomcp_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    declared |= set(re.findall(r"\b(vi_[a-z_0-9]+)\s*\(", open(os.path.join(ROOT, "include", "vi_b200.h")).read()))
    assert {"vi_solve_reference", "vi_project_textbook_dev"} <= declared
    for name in declared:
        assert isinstance(getattr(L, name), ctypes._CFuncPtr)


def test_no_cpu_fallback():
    """Without a CUDA device the product fails loudly instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pomdpy_b200 import native
    with pytest.raises(native.PomcpError, match="no CUDA device"):
        native.Planner()


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pomdpy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "pomcp_oracle" not in src, f


def test_preferred_action_heuristic_golden():
    """create_child / generate_smart_actions (rock_position_history.py:96-137, 170-234): the oracle's per-node
    functions and the product's host classes against node states recorded from the executed reference."""
    from pomdpy_b200.rock_sample import PositionAndRockData, RockAction, RockData, RockModel, RockObservation
    with open(os.path.join(GOLDEN, "smart_actions.json")) as f:
        chains = json.load(f)["chains"]
    models, oms = {}, {}
    for ch in chains:
        tag = ch["map"]
        if tag not in models:
            models[tag] = RockModel(dict(map_file=tag, discount=0.95, preferred_actions=True))
            oms[tag] = oracle_model(golden_tables(tag))
        m, om = models[tag], oms[tag]
        K = m.n_rocks
        data = PositionAndRockData(m, m.start_position.copy(), [RockData() for _ in range(K)], None)
        st, cell = oracle.rockstats_init(om), om.start_cell
        for s in ch["steps"]:
            a, good = s["action"], bool(s["obs_good"])
            data = data.create_child(RockAction(a), RockObservation(good, False) if a >= 4 else RockObservation())
            st = oracle.rockstats_child(om, st, cell, a, good)
            cell = s["cell"]
            assert m.cell_index(data.grid_position) == cell
            assert [r.goodness_number for r in data.all_rock_data] == s["good"] == [st.good[k] for k in range(K)]
            assert [r.chance_good for r in data.all_rock_data] == s["chance"] == [st.chance[k] for k in range(K)]
            assert sum(1 << x for x in data.generate_smart_actions()) == s["mask"] == oracle.smart_mask(om, st, cell)


def test_gs_oracle_regression_pins():
    """The generation-synchronous oracle is the specification of the CUDA kernels; tests/golden/gs_pins.json pins its
    output on eleven fixed searches (RockSample with and without preferred actions / rollouts, tabular models) so that
    a change of the specification cannot slip in unnoticed (oracle/refharness/make_gs_pins.py regenerates them)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_gs_pins", os.path.join(ROOT, "oracle", "refharness", "make_gs_pins.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(GOLDEN, "gs_pins.json")) as f:
        pins = json.load(f)["pins"]
    assert len(pins) == len(mod.CASES)
    for pin in pins:
        case = dict(pin["case"])
        case["sched"] = tuple(case["sched"])
        assert mod.run_case(case) == pin["moves"], case
