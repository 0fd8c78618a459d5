"""The Tiger problem as the reference ships it (examples/tiger/tiger_model.py): the data source of the value
iteration kernel (T, O, R, initial belief) plus the small generative environment the VI experiment loop steps.

Actions: 0 LISTEN, 1 OPEN_DOOR_1, 2 OPEN_DOOR_2.  States: tiger behind door 1 / door 2.  Listening costs -1 and
is correct with probability 0.85; opening the tiger's door -20, the other door +10 (tiger_model.py:102-135)."""
import numpy as np

from ..discrete import DiscreteAction, DiscreteActionPool, DiscreteObservation, DiscreteObservationPool, DiscreteState
from ..model import Model, StepResult


class ActionType(object):
    LISTEN, OPEN_DOOR_1, OPEN_DOOR_2 = 0, 1, 2


class TigerAction(DiscreteAction):
    def copy(self):
        return TigerAction(self.bin_number)

    def to_string(self):
        return ("Listening", "Opening door 1", "Opening door 2")[self.bin_number] \
            if 0 <= self.bin_number <= 2 else "Unknown action type"


class TigerObservation(DiscreteObservation):
    """source_of_roar: [1, 0] roar heard from door 1, [0, 1] from door 2, None after opening a door."""

    def __init__(self, source_of_roar):
        super(TigerObservation, self).__init__(-1 if source_of_roar is None else (0 if source_of_roar[0] else 1))
        self.source_of_roar = source_of_roar

    def copy(self):
        return TigerObservation(self.source_of_roar)

    def to_string(self):
        if self.source_of_roar is None:
            return "The door was opened"
        return "Roaring is heard coming from door %d" % (1 if self.source_of_roar[0] else 2)


class TigerState(DiscreteState):
    def __init__(self, door_open, door_prizes):
        self.door_open = door_open
        self.door_prizes = door_prizes

    def copy(self):
        return TigerState(self.door_open, list(self.door_prizes))

    def as_list(self):
        return [self.door_open] + list(self.door_prizes)

    def to_string(self):
        return ("Door is open - " if self.door_open else "Door is closed - ") + str(self.door_prizes)


class TigerModel(Model):
    def __init__(self, args):
        super(TigerModel, self).__init__(args if isinstance(args, dict) else {})
        self.tiger_door = None
        self.num_doors, self.num_states, self.num_actions, self.num_observations = 2, 2, 3, 2
        if not hasattr(self, "discount"):
            self.discount = 0.95

    def start_scenario(self):
        self.tiger_door = int(np.random.randint(0, self.num_doors)) + 1

    # matrices (tiger_model.py:102-139)
    @staticmethod
    def get_transition_matrix():
        return np.array([[[1.0, 0.0], [0.0, 1.0]], [[0.5, 0.5], [0.5, 0.5]], [[0.5, 0.5], [0.5, 0.5]]])

    @staticmethod
    def get_observation_matrix():
        return np.array([[[0.85, 0.15], [0.15, 0.85]], [[0.5, 0.5], [0.5, 0.5]], [[0.5, 0.5], [0.5, 0.5]]])

    @staticmethod
    def get_reward_matrix():
        return np.array([[-1., -1.], [-20.0, 10.0], [10.0, -20.0]])

    @staticmethod
    def get_initial_belief_state():
        return np.array([0.5, 0.5])

    # Model interface
    def is_terminal(self, state):
        return bool(state.door_open)

    def is_valid(self, _):
        return True

    def sample_state_uninformed(self):
        cfg = [0, 1]
        if np.random.uniform(0, 1) <= 0.5:
            cfg.reverse()
        return TigerState(False, cfg)

    def sample_an_init_state(self):
        return self.sample_state_uninformed()

    def sample_state_informed(self, belief):
        return TigerState(False, [0, 1] if 100. * np.random.random() <= 100. * belief[0] else [1, 0])

    def get_all_states(self):
        return [[False, 0, 1], [False, 1, 0]]

    def get_all_actions(self):
        return [TigerAction(a) for a in range(3)]

    def get_all_observations(self):
        return [0, 1]

    def get_legal_actions(self, _):
        return self.get_all_actions()

    def reset_for_simulation(self):
        self.start_scenario()

    def reset_for_epoch(self):
        self.start_scenario()

    def update(self, sim_data):
        pass

    def get_max_undiscounted_return(self):
        return 10

    def create_action_pool(self):
        return DiscreteActionPool(self)

    def create_observation_pool(self, solver):
        return DiscreteObservationPool(solver)

    def create_root_historical_data(self, agent):
        return None

    def generate_step(self, action, state=None):
        """(action[, state]) -> StepResult: the signature of the reference's Tiger (tiger_model.py:151-164)."""
        if not isinstance(action, TigerAction):
            action = TigerAction(int(action))
        result = StepResult()
        result.is_terminal = action.bin_number != ActionType.LISTEN
        result.action = action.copy()
        result.observation = self.make_observation(action)
        result.reward = self.make_reward(action, result.is_terminal)
        return result

    def make_reward(self, action, is_terminal):
        if action.bin_number == ActionType.LISTEN:
            return -1.0
        return -20. if action.bin_number == self.tiger_door else 10.0

    def make_observation(self, action):
        if action.bin_number > 0:
            return TigerObservation(None)
        obs = [1, 0] if self.tiger_door == 1 else [0, 1]
        if np.random.uniform(0, 1) > 0.85:
            obs.reverse()
        return TigerObservation(obs)

    def belief_update(self, old_belief, action, observation):
        """Bayes filter for LISTEN (tiger_model.py:217-244)."""
        a = action if isinstance(action, (int, np.integer)) else action.bin_number
        if a > 0:
            return old_belief
        pc, pi = 0.85, 0.15
        l1, l2 = (pc, pi) if observation.source_of_roar[0] else (pi, pc)
        post = np.array([l1 * old_belief[0], l2 * old_belief[1]])
        return post / post.sum()
