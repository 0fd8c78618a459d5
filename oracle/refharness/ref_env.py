"""Load the *executed* reference (pemami4911/POMDPy) from a writable copy.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the C
oracle under ``oracle/`` against the real reference and (b) generate the golden
vectors committed under ``tests/golden/``.  ``/root/reference`` does not exist
on the GPU box, so nothing under ``pomdpy_b200/``, ``bench.py`` or the ``-m gpu``
tests may import this module.

Why a copy: ``Model.__init__`` (reference pomdpy/pomdp/model.py:34-51) mkdirs
``experiments/{pickle_jar,checkpoints,tensorboard/<n>}`` on every construction
and ``init_logger`` (pomdpy/log/log_pomdpy.py:13-16) opens a log file inside the
package, so the reference is never executed in place.

Why shims: 15 reference files import ``past.utils.old_div`` /
``future.utils.with_metaclass`` (e.g. pomdpy/solvers/pomcp.py:4); the ``future``
package is absent here.  Two tiny modules restate those helpers.
"""
import io
import json
import os
import shutil
import sys
import contextlib

REFERENCE_SRC = os.environ.get("POMDPY_REFERENCE", "/root/reference")
DEFAULT_WORKDIR = os.environ.get("POMDPY_REF_COPY", "/tmp/pomdpy_ref_copy")

_PAST_UTILS = '''
import numbers
def old_div(a, b):
    if isinstance(a, numbers.Integral) and isinstance(b, numbers.Integral):
        return a // b
    return a / b
'''
_FUTURE_UTILS = '''
def with_metaclass(meta, *bases):
    class metaclass(meta):
        def __new__(cls, name, this_bases, d):
            return meta(name, bases, d)
    return type.__new__(metaclass, "temporary_class", (), {})
'''


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "pomdpy"))


def prepare_copy(workdir=DEFAULT_WORKDIR, map_file="map-7-8.txt"):
    """Copy the reference tree, write the shims, select the RockSample map."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_SRC)
    ref = os.path.join(workdir, "ref")
    if not os.path.isdir(ref):
        os.makedirs(workdir, exist_ok=True)
        shutil.copytree(REFERENCE_SRC, ref)
    shim = os.path.join(workdir, "shim")
    os.makedirs(os.path.join(shim, "past"), exist_ok=True)
    os.makedirs(os.path.join(shim, "future"), exist_ok=True)
    for pkg, mod, body in (("past", "utils", _PAST_UTILS), ("future", "utils", _FUTURE_UTILS)):
        open(os.path.join(shim, pkg, "__init__.py"), "w").close()
        with open(os.path.join(shim, pkg, mod + ".py"), "w") as f:
            f.write(body)
    cfg_path = os.path.join(ref, "pomdpy", "config", "rock_sample_config.json")
    with open(cfg_path) as f:
        cfg = json.load(f)
    if cfg.get("map_file") != map_file:
        cfg["map_file"] = map_file
        with open(cfg_path, "w") as f:
            json.dump(cfg, f)
    return ref, shim


def import_reference(workdir=DEFAULT_WORKDIR, map_file="map-7-8.txt"):
    """Return the dict of reference classes, imported from the copy."""
    import warnings
    ref, shim = prepare_copy(workdir, map_file)
    for p in (shim, ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")           # `is` with a literal: SyntaxWarning on py3.12
        import pomdpy                              # noqa: F401
        from pomdpy import Agent
        from pomdpy.solvers import POMCP
        from pomdpy.pomdp import model as ref_model
        from examples.rock_sample import RockModel
        from examples.rock_sample.rock_state import RockState
        from examples.rock_sample.rock_action import RockAction
        from examples.rock_sample.grid_position import GridPosition
        import pomdpy.util.console as ref_console
    ref_console.VERBOSITY = 0
    return dict(Agent=Agent, POMCP=POMCP, RockModel=RockModel, RockState=RockState,
                RockAction=RockAction, GridPosition=GridPosition, model=ref_model,
                console=ref_console, root=ref)


POMCP_DEFAULT_ARGS = dict(
    env="RockSample", solver="POMCP", seed=1993, use_tf=False, discount=0.95, n_epochs=100,
    max_steps=200, save=False, test=10, epsilon_start=0.5, epsilon_minimum=0.1,
    epsilon_decay=0.95, epsilon_decay_step=20, n_sims=500, timeout=3600,
    preferred_actions=False, ucb_coefficient=3.0, n_start_states=2000,
    min_particle_count=1000, max_particle_count=2000, max_depth=100,
    action_selection_timeout=60)


def make_rock_model(ref, map_file, **overrides):
    """Construct the reference RockModel on `map_file` (stdout of its pprint swallowed)."""
    prepare_copy(os.path.dirname(ref["root"]), map_file)
    args = dict(POMCP_DEFAULT_ARGS)
    args.update(overrides)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref["RockModel"](args)
    return m
