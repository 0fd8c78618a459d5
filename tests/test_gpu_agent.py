"""The drop-in boundary on the GPU: the reference's CLI flags -> Agent -> CUDA POMCP solver."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_cli_runs_epochs_and_returns_are_sane(capsys):
    import pomcp as cli
    agent = cli.main(["--env", "RockSample", "--solver", "POMCP", "--max_steps", "200", "--n_epochs", "12",
                      "--n_sims", "500", "--seed", "123", "--verbosity", "1"])
    er = agent.experiment_results
    assert er.discounted_return.count == 12
    # reference on this configuration: 14.6 (executed, 40 epochs) / 15.75 +- 0.51 (restatement, 60 epochs)
    assert 8.0 < er.discounted_return.mean < 26.0
    assert agent.total_sims >= 12 * 500


def test_solver_surface_matches_what_agent_touches():
    from pomdpy_b200 import Agent
    from pomdpy_b200.rock_sample import RockModel, RockState
    from pomdpy_b200.solvers import POMCP
    from pomdpy_b200 import util
    util.VERBOSITY = 0
    np.random.seed(5)
    args = dict(env="RockSample", solver="POMCP", seed=5, discount=0.95, n_epochs=1, max_steps=50, n_sims=300,
                preferred_actions=False, ucb_coefficient=3.0, n_start_states=2000, min_particle_count=1000,
                max_particle_count=2000, max_depth=100, action_selection_timeout=60, timeout=3600,
                epsilon_start=0.5, epsilon_minimum=0.1, epsilon_decay=0.95, map_file="map-7-8")
    model = RockModel(args)
    model.reset_for_epoch()
    agent = Agent(model, POMCP)
    solver = agent.solver_factory(agent)
    state = solver.belief_tree_index.sample_particle()
    assert isinstance(state, RockState) and len(solver.belief_tree_index.state_particles) == 2000
    action = solver.select_eps_greedy_action(0.5, __import__("time").time())
    assert hasattr(action, "bin_number") and action.to_string() and action.copy().bin_number == action.bin_number
    amap = solver.belief_tree_index.action_map
    assert amap.total_visit_count == 300
    assert action.bin_number in amap.bin_sequence and action.bin_number in solver.belief_tree_index.data.legal_actions()
    best = max(e.mean_q_value for e in amap.entries.values() if e.is_legal)
    assert amap.entries[action.bin_number].mean_q_value == best
    res, legal = model.generate_step(state, action)
    solver.update(res)
    assert not solver.disable_tree and len(solver.belief_tree_index.state_particles) == 2000
    solver.history.add_entry()
    assert solver.history.get_length() == 1


def test_cli_preferred_actions_runs():
    """The README command of the reference (README.md:85): --preferred_actions."""
    import pomcp as cli
    agent = cli.main(["--env", "RockSample", "--solver", "POMCP", "--max_steps", "200", "--n_epochs", "20",
                      "--n_sims", "500", "--seed", "123", "--preferred_actions", "--verbosity", "0"])
    er = agent.experiment_results
    print("preferred actions: %.2f +- %.2f" % (er.discounted_return.mean, er.discounted_return.std_err()))
    assert er.discounted_return.count == 20
    assert -5.0 < er.discounted_return.mean < 26.0        # reference: 7.2 pooled over 40 epochs


def test_batched_epochs_match_serial_level_and_are_faster():
    """Episode-level batching (pomdpy_b200/batch.py): 256 epochs on 32 concurrent lanes."""
    import time
    import pomcp as cli
    t0 = time.time()
    agent = cli.main(["--env", "RockSample", "--solver", "POMCP", "--n_epochs", "256", "--n_sims", "500", "--seed", "123",
                      "--verbosity", "0", "--concurrent_epochs", "32"])
    dt = time.time() - t0
    st = agent.discounted_return
    print("batched: %.2f +- %.2f over %d epochs, %.1f ms/epoch (%.0f epochs/s)" % (st.mean, st.std_err(), st.count,
                                                                                  1e3 * agent.wall_seconds / 256, 256 / agent.wall_seconds))
    assert st.count == 256 and 14.62 - 2.0 < st.mean < 20.5
    assert agent.total_sims >= 256 * 500 and dt < 60


def test_cli_preferred_rollouts_improve_return():
    """Additive --preferred_rollouts: the rollout policy uses the smart actions of the rollout's own history.  On the
    reference's configuration the return rises well above the tree-only heuristic's (~21) and the plain planner's (~16)."""
    import pomcp as cli
    agent = cli.main(["--env", "RockSample", "--solver", "POMCP", "--max_steps", "200", "--n_epochs", "30",
                      "--n_sims", "500", "--seed", "123", "--preferred_actions", "--preferred_rollouts",
                      "--verbosity", "0"])
    er = agent.experiment_results
    print("preferred rollouts: %.2f +- %.2f" % (er.discounted_return.mean, er.discounted_return.std_err()))
    assert er.discounted_return.mean > 16.0


def _solver_on_rock(rock=0, n_sims=2000):
    """A solver whose root belief stands ON `rock` of map-7-8 (so that CHECK of that rock is certain, p = 1)."""
    from pomdpy_b200 import Agent, util
    from pomdpy_b200.rock_sample import RockModel
    from pomdpy_b200.solvers import POMCP
    util.VERBOSITY = 0
    np.random.seed(11)
    args = dict(env="RockSample", solver="POMCP", seed=11, discount=0.95, n_epochs=1, max_steps=50, n_sims=n_sims,
                preferred_actions=False, ucb_coefficient=3.0, n_start_states=2000, min_particle_count=200,
                max_particle_count=2000, max_depth=100, action_selection_timeout=60, timeout=3600,
                epsilon_start=0.5, epsilon_minimum=0.1, epsilon_decay=0.95, map_file="map-7-8")
    model = RockModel(args)
    model.reset_for_epoch()
    model.start_position = model.rock_positions[rock]               # the root belief is drawn at the start position
    agent = Agent(model, POMCP)
    return model, agent.solver_factory(agent)


def test_update_failure_recovers_with_uninformed_particles_like_the_reference():
    """belief_tree_solver.py:196-203: no particle of the old belief reproduces the observation -> a new particle set with
    uniformly drawn rocks that does.  Observations come from the true world, so the failure is forced by
    letting the device's world disagree with the host model's about rock 0 (robot on the rock: the CHECK is certain)."""
    import time
    from pomdpy_b200.model import StepResult
    from pomdpy_b200.rock_sample.types import RockAction, RockObservation, RockState
    model, solver = _solver_on_rock(0)
    _, cell = solver.planner.root_particles(fetch=False)
    model.actual_rock_states[0] = True                               # host: rock 0 is good ...
    actual, sampled = model.world_masks()
    solver.planner.set_world(actual & ~1, sampled)                   # ... device: bad, so "good" cannot be observed there
    solver.select_eps_greedy_action(0.5, time.time())
    res = StepResult()
    res.action, res.observation = RockAction(5), RockObservation(True, False)
    res.next_state = RockState(model.rock_positions[0], [True] * model.n_rocks)
    res.reward, res.is_terminal = 0.0, False
    solver.update(res)
    assert not solver.disable_tree                                   # recovered: tree search goes on from a fresh root
    nb, ncell = solver.planner.root_particles()
    assert ncell == cell and len(nb) == 200 and np.all((nb >> 8) & 1 == 1)     # min_particle_count particles, rock 0 good
    a = solver.select_eps_greedy_action(0.5, time.time())
    assert a.bin_number in model.legal_actions_at(model.rock_positions[0])


def test_impossible_observation_sets_disable_tree_and_the_next_action_is_legal_at_the_true_position():
    """The observation cannot be produced at all (uninformed refill fails too): disable_tree is set as in the reference
    (belief_tree_solver.py:205-208), and select_eps_greedy_action still returns an action that is legal at the TRUE
    position -- the old belief is pushed through the real action and planted there (the reference would go on
    from the stale node)."""
    import time
    from pomdpy_b200.model import StepResult
    from pomdpy_b200.rock_sample.types import RockAction, RockObservation, RockState
    model, solver = _solver_on_rock(0)
    _, cell0 = solver.planner.root_particles(fetch=False)
    solver.select_eps_greedy_action(0.5, time.time())
    east = model.make_next_position(model.rock_positions[0], 1)[0]
    res = StepResult()
    res.action, res.observation = RockAction(1), RockObservation(True, False)   # moving never observes "good"
    res.next_state = RockState(east, [True] * model.n_rocks)
    res.reward, res.is_terminal = 0.0, False
    solver.update(res)
    assert solver.disable_tree
    nb, ncell = solver.planner.root_particles()
    assert ncell == model.cell_index(east) != cell0 and len(nb) == 2000 and np.all((nb & 0xFF) == ncell)
    for _ in range(3):
        a = solver.select_eps_greedy_action(0.5, time.time())
        assert a.bin_number in model.legal_actions_at(east)


REF_DRIVER = r'''
import sys, io, contextlib, warnings, random, json
warnings.simplefilter("ignore")
sys.path.insert(0, ROOT); sys.path.insert(0, REF + "/shim"); sys.path.insert(0, REF + "/ref")
import numpy as np
from pomdpy import Agent                                   # the REFERENCE's Agent ...
from examples.rock_sample import RockModel                 # ... and RockModel, unmodified
import pomdpy.util.console as cons
from pomdpy_b200.solvers import POMCP                      # the CUDA solver: the one-line swap of INTEGRATION.md
cons.VERBOSITY = 0
args = dict(env="RockSample", solver="POMCP", seed=7, use_tf=False, discount=0.95, n_epochs=3, max_steps=60, save=False,
            test=10, epsilon_start=0.5, epsilon_minimum=0.1, epsilon_decay=0.95, epsilon_decay_step=20, n_sims=2000,
            timeout=3600, preferred_actions=False, ucb_coefficient=3.0, n_start_states=2000, min_particle_count=1000,
            max_particle_count=2000, max_depth=100, action_selection_timeout=60)
np.random.seed(7); random.seed(7)
with contextlib.redirect_stdout(io.StringIO()):
    model = RockModel(args)
    agent = Agent(model, POMCP)
    agent.discounted_return()
er = agent.experiment_results
print(json.dumps(dict(cls=type(model).__module__, epochs=er.discounted_return.count, mean=er.discounted_return.mean,
                      rocks=model.n_rocks)))
'''


def test_unmodified_reference_agent_and_model_drive_the_cuda_solver():
    """SURVEY 8(b) / INTEGRATION.md: the reference's own Agent and RockModel (the copy under baseline/_ref, which travels
    to the GPU box) run epochs with pomdpy_b200.solvers.POMCP plugged in; nothing is added to the reference model --
    pomdpy_b200.adapters extracts the device tables from its public attributes."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "ref", "pomdpy")):
        pytest.skip("baseline/_ref not installed (python tools/install_reference.py)")
    res = subprocess.run([sys.executable, "-c", "ROOT = %r\nREF = %r\n" % (root, ref) + REF_DRIVER], capture_output=True,
                         text=True, timeout=500, cwd=os.path.join(ref, "ref"))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["cls"].startswith("examples.rock_sample") and out["epochs"] == 3
    assert -40.0 < out["mean"] < 60.0
