"""pomdpy_b200.adapters against an UNMODIFIED reference RockModel (build container only: needs /root/reference; the
golden tables in tests/golden/ were produced by the executed reference).  SURVEY 8(b): isinstance(model, RockModel)
extraction instead of asking the maintainer to add device hooks to their model."""
import os
import sys

import numpy as np
import pytest

from helpers import MAPS, ROOT, golden_tables

sys.path.insert(0, os.path.join(ROOT, "oracle", "refharness"))
import ref_env  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_env.reference_available(), reason="reference tree not present on this box")


@pytest.fixture(autouse=True)
def _keep_sys_path():
    """ref_env puts the reference copy first on sys.path (its root holds a pomcp.py / vi.py of its own): undo that after
    every test so that later tests import this repository's CLI modules."""
    saved = list(sys.path)
    yield
    sys.path[:] = saved


@pytest.mark.parametrize("tag", MAPS)
def test_spec_extracted_from_the_reference_model_equals_the_golden_tables(tag):
    from pomdpy_b200 import adapters
    ref = ref_env.import_reference(map_file=tag + ".txt")
    model = ref_env.make_rock_model(ref, tag + ".txt")
    assert type(model).__module__.startswith("examples.rock_sample")      # the reference's class, not this package's
    spec, _ = adapters.spec_from_reference(model)
    t = golden_tables(tag)
    assert (spec["rows"], spec["cols"], spec["start_cell"]) == (int(t["rows"]), int(t["cols"]), int(t["start_cell"]))
    assert np.array_equal(spec["cell"], t["cell"].reshape(-1))
    assert np.array_equal(spec["thr53"], t["thr53"].reshape(spec["thr53"].shape))
    assert np.array_equal(spec["prob"], t["prob"].reshape(spec["prob"].shape))
    assert np.array_equal(spec["rewards"], t["rewards"])
    # the mutable world and the particle format
    np.random.seed(3)
    model.reset_for_epoch()
    h = adapters.hooks_for(model)
    assert isinstance(h, adapters.RockSampleHooks)
    actual, sampled = h.world_masks()
    assert sampled == 0 and [bool((actual >> k) & 1) for k in range(model.n_rocks)] == [bool(x) for x in model.actual_rock_states]
    model.unique_rocks_sampled = [1, 3]
    assert h.world_masks()[1] == 0b1010
    for _ in range(20):
        s = model.sample_an_init_state()
        p = h.pack_state(s)
        assert p & 0xFF == spec["start_cell"]
        back = h.unpack_state(p)
        assert type(back) is type(s) and back.position == s.position and list(back.rock_states) == [bool(x) for x in s.rock_states]


def test_this_packages_models_keep_their_own_hooks():
    from pomdpy_b200 import adapters
    from pomdpy_b200.rock_sample import RockModel
    m = RockModel(dict(map_file="map-7-8", discount=0.95))
    assert adapters.hooks_for(m) is m
    with pytest.raises(TypeError):
        adapters.hooks_for(object())


def test_saved_alpha_vectors_are_read_by_the_reference(tmp_path):
    """--save writes what the reference's own tooling reads: experiments/scripts/pickle_wrapper.load_pkl returns a set
    of the REFERENCE's pomdpy.solvers.alpha_vector.AlphaVector; and this package reads the reference's committed pickle."""
    import subprocess
    from pomdpy_b200 import vi
    gamma = [vi.AlphaVector(0, np.array([1.5, -2.0])), vi.AlphaVector(2, np.array([0.25, 7.0]))]
    path = str(tmp_path / "VI_planning_horizon_3.pkl")
    vi.save_alpha_vectors(gamma, path)
    back = sorted(vi.load_alpha_vectors(path), key=lambda a: a.action)
    assert [a.action for a in back] == [0, 2] and np.array_equal(back[1].v, [0.25, 7.0])
    ref, shim = ref_env.prepare_copy()
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from experiments.scripts.pickle_wrapper import load_pkl\n"
            "g = load_pkl(%r)\n"
            "assert isinstance(g, set) and len(g) == 2\n"
            "for a in g: assert type(a).__module__ == 'pomdpy.solvers.alpha_vector' and a.copy().action == a.action\n"
            "print(sorted((a.action, a.v.tolist()) for a in g))\n") % (shim, ref, path)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ref)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "[(0, [1.5, -2.0]), (2, [0.25, 7.0])]" in res.stdout
    theirs = vi.load_alpha_vectors(os.path.join(ref_env.REFERENCE_SRC, "experiments", "pickle_jar", "VI_planning_horizon_1.pkl"))
    assert len(theirs) >= 1 and all(a.v.dtype == np.float64 and a.v.size == 2 for a in theirs)
