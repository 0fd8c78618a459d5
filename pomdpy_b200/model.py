"""The Model plugin API (mirrors reference pomdpy/pomdp/model.py:13-293).

A model is the black-box generative simulator (s, a) -> (s', o, r) plus factories for the action /
observation pools.  Every key of the CLI argument dict becomes an attribute (model.py:30-31), which is how
the solvers read their hyper-parameters.  Unlike the reference, constructing a model has no file-system
side effects (the reference mkdirs experiments/* on every construction, model.py:34-51).

For the CUDA planner a model additionally implements ``device_spec()`` (tables uploaded once) and
``world_masks()`` (the mutable part: re-uploaded per epoch / per real step)."""
import abc
import random


class StepResult(object):
    """(action, observation, reward, next_state, is_terminal) of one generative step (model.py:280-293)."""

    def __init__(self):
        self.action = None
        self.observation = None
        self.reward = 0
        self.next_state = None
        self.is_terminal = 0

    def print_step_result(self):
        print("------- Step Result --------")
        print("Action: ", self.action.to_string())
        print("Observation: ", self.observation.to_string())
        print("Reward: ", self.reward)
        print("Next state: ", self.next_state.to_string())
        print("Is terminal: ", self.is_terminal)


class Model(abc.ABC):
    def __init__(self, args):
        for key, value in args.items():
            setattr(self, key, value)

    # ---- simulator life-cycle -------------------------------------------------------------------------
    @abc.abstractmethod
    def reset_for_simulation(self):
        """Called before each simulation (belief_tree_solver.py:55)."""

    @abc.abstractmethod
    def reset_for_epoch(self):
        """Called before each epoch (agent.py:135,143)."""

    @abc.abstractmethod
    def update(self, sim_data):
        """Fold a real StepResult into the simulator (belief_tree_solver.py:165)."""

    # ---- generative model -----------------------------------------------------------------------------
    @abc.abstractmethod
    def generate_step(self, state, action):
        """(state, action) -> (StepResult, is_legal)."""

    @abc.abstractmethod
    def sample_an_init_state(self):
        """A state drawn from the initial belief."""

    @abc.abstractmethod
    def sample_state_uninformed(self):
        """A state drawn from an uninformed prior."""

    @abc.abstractmethod
    def sample_state_informed(self, belief):
        """A state drawn from `belief`."""

    @abc.abstractmethod
    def belief_update(self, old_belief, action, observation):
        """Exact Bayes filter, for models that have one."""

    @abc.abstractmethod
    def get_all_states(self):
        pass

    @abc.abstractmethod
    def get_all_actions(self):
        pass

    @abc.abstractmethod
    def get_all_observations(self):
        pass

    @abc.abstractmethod
    def get_legal_actions(self, state):
        pass

    @abc.abstractmethod
    def is_terminal(self, state):
        pass

    @abc.abstractmethod
    def is_valid(self, state):
        pass

    @abc.abstractmethod
    def create_action_pool(self):
        pass

    @abc.abstractmethod
    def create_root_historical_data(self, solver):
        pass

    @abc.abstractmethod
    def create_observation_pool(self, solver):
        pass

    @abc.abstractmethod
    def get_max_undiscounted_return(self):
        pass

    # ---- optional matrix hooks (value iteration) --------------------------------------------------------
    @staticmethod
    def get_initial_belief_state():
        return None

    @staticmethod
    def get_transition_matrix():
        return None

    @staticmethod
    def get_observation_matrix():
        return None

    @staticmethod
    def get_reward_matrix():
        return None

    # ---- default particle filter (model.py:221-252), host version --------------------------------------
    def generate_particles(self, previous_belief, action, obs, n_particles, prev_particles, max_tries=None):
        """Rejection sampling on the host: kept for API compatibility and for models without a device
        spec.  The CUDA solver never calls it -- its belief update is the batched kernel behind
        pomcp_advance()."""
        out, tries = [], 0
        limit = max_tries if max_tries is not None else 1000 * max(int(n_particles), 1)
        while len(out) < n_particles and tries < limit:
            tries += 1
            result, _ = self.generate_step(random.choice(prev_particles), action)
            if result.observation == obs:
                out.append(result.next_state)
        return out
