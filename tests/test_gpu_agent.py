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
