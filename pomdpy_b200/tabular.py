"""Tabular POMDPs on the device planner: any model that can state its dynamics as matrices -- the reference's
``Model`` ABC asks for exactly these through get_transition_matrix / get_observation_matrix / get_reward_matrix
(pomdpy/pomdp/model.py:133-160; the Tiger example fills them in, examples/tiger/tiger_model.py:102-139).

The CUDA planner samples from 32-bit cumulative thresholds; `cumulative_thresholds` is the one place where
probabilities become integers, so the host environment step below, the device kernels and the test oracle all draw
the same outcome from the same 32-bit word."""
import numpy as np

from .discrete import DiscreteAction, DiscreteActionPool, DiscreteObservation, DiscreteObservationPool, DiscreteState
from .model import Model, StepResult


def cumulative_thresholds(prob):
    """[..., n] probabilities -> uint32 thresholds min(floor(2^32 * cumsum), 2^32 - 1); rows are renormalised and end
    at 0xFFFFFFFF.  A draw w picks the first index k with thr[k] > min(w, 2^32 - 2)."""
    p = np.asarray(prob, dtype=np.float64)
    if p.ndim < 1 or np.any(p < 0) or np.any(p.sum(axis=-1) <= 0):
        raise ValueError("rows must be non-negative with a positive sum")
    c = np.cumsum(p, axis=-1)
    c = c / c[..., -1:]
    t = np.minimum(np.floor(c * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
    last = np.argmax(np.cumsum(p > 0, axis=-1), axis=-1)         # index of the last positive entry of each row
    idx = np.arange(p.shape[-1])
    t[idx >= last[..., None]] = 0xFFFFFFFF
    return np.ascontiguousarray(t)


def draw(thr_row, w):
    """Host twin of the device sampler."""
    w = min(int(w), 0xFFFFFFFE)
    return int(np.searchsorted(thr_row, w, side="right"))


class TabularState(DiscreteState):
    def __init__(self, index):
        self.index = int(index)

    def copy(self):
        return TabularState(self.index)

    def as_list(self):
        return [self.index]

    def to_string(self):
        return "s%d" % self.index

    def __eq__(self, other):
        return isinstance(other, TabularState) and other.index == self.index

    def __hash__(self):
        return hash(self.index)


class TabularAction(DiscreteAction):
    def copy(self):
        return TabularAction(self.bin_number)

    def to_string(self):
        return "a%d" % self.bin_number


class TabularObservation(DiscreteObservation):
    def copy(self):
        return TabularObservation(self.bin_number)

    def to_string(self):
        return "z%d" % self.bin_number


class TabularModel(Model):
    """A POMDP given by T[a][s][s'], O[a][s'][z], R[a][s], an initial belief and the set of episode-ending actions /
    absorbing states.  Implements the reference's Model interface plus the device hooks (device_spec)."""

    def __init__(self, args, T, O, R, initial_belief, terminal_actions=(), terminal_states=(), name="Tabular"):
        super(TabularModel, self).__init__(args if isinstance(args, dict) else {})
        self.T = np.asarray(T, dtype=np.float64)
        self.O = np.asarray(O, dtype=np.float64)
        self.R = np.asarray(R, dtype=np.float64)
        self.b0 = np.asarray(initial_belief, dtype=np.float64)
        self.num_actions, self.num_states, _ = self.T.shape
        self.num_observations = self.O.shape[2]
        if self.T.shape != (self.num_actions, self.num_states, self.num_states) or \
                self.O.shape[:2] != (self.num_actions, self.num_states) or \
                self.R.shape != (self.num_actions, self.num_states) or self.b0.shape != (self.num_states,):
            raise ValueError("expected T[A][S][S], O[A][S][Z], R[A][S], b0[S]")
        self.terminal_actions = tuple(int(a) for a in terminal_actions)
        self.term_action_mask = sum(1 << a for a in self.terminal_actions)
        self.term_state = np.zeros(self.num_states, dtype=np.uint8)
        self.term_state[list(terminal_states)] = 1
        self.Tc = cumulative_thresholds(self.T)
        self.Oc = cumulative_thresholds(self.O)
        self.b0c = cumulative_thresholds(self.b0)
        self.env_name = name
        self.true_state = None
        if not hasattr(self, "discount"):
            self.discount = 0.95
        self._spec = None

    # ---- device hooks -------------------------------------------------------------------------------------
    def device_spec(self):
        if self._spec is None:
            self._spec = dict(kind="tabular", Tc=self.Tc, Oc=self.Oc, R=self.R,
                              term_action_mask=self.term_action_mask, term_state=self.term_state)
        return self._spec

    def world_masks(self):
        return 0, 0

    def pack_state(self, state):
        return int(state.index)

    def unpack_state(self, packed):
        return TabularState(int(packed))

    def sample_init_states_packed(self, n, rng=None):
        rng = np.random if rng is None else rng
        w = rng.randint(0, 1 << 32, size=int(n), dtype=np.uint64).astype(np.uint32)
        return np.searchsorted(self.b0c, np.minimum(w, 0xFFFFFFFE), side="right").astype(np.uint32)

    # ---- the reference's Model interface (pomdpy/pomdp/model.py) ---------------------------------------------
    def get_transition_matrix(self):
        return self.T

    def get_observation_matrix(self):
        return self.O

    def get_reward_matrix(self):
        return self.R

    def get_initial_belief_state(self):
        return self.b0

    def start_scenario(self):
        self.true_state = TabularState(int(self.sample_init_states_packed(1)[0]))

    def reset_for_simulation(self):
        self.start_scenario()

    def reset_for_epoch(self):
        self.start_scenario()

    def update(self, sim_data):
        pass

    def is_terminal(self, state):
        return bool(self.term_state[state.index])

    def is_valid(self, state):
        return 0 <= state.index < self.num_states

    def sample_an_init_state(self):
        return TabularState(int(self.sample_init_states_packed(1)[0]))

    def sample_state_uninformed(self):
        return TabularState(int(np.random.randint(0, self.num_states)))

    def sample_state_informed(self, belief):
        return TabularState(int(np.random.choice(self.num_states, p=np.asarray(belief) / np.sum(belief))))

    def get_all_states(self):
        return [TabularState(s) for s in range(self.num_states)]

    def get_all_actions(self):
        return [TabularAction(a) for a in range(self.num_actions)]

    def get_all_observations(self):
        return [TabularObservation(z) for z in range(self.num_observations)]

    def get_legal_actions(self, _):
        return self.get_all_actions()

    def get_max_undiscounted_return(self):
        return float(self.R.max())

    def create_action_pool(self):
        return DiscreteActionPool(self)

    def create_observation_pool(self, solver):
        return DiscreteObservationPool(solver)

    def create_root_historical_data(self, agent):
        return None

    def belief_update(self, old_belief, action, observation):
        """Exact Bayes filter (what the reference's VI loop uses for Tiger, tiger_model.py:166-199)."""
        a, z = int(getattr(action, "bin_number", action)), int(getattr(observation, "bin_number", observation))
        b = (np.asarray(old_belief) @ self.T[a]) * self.O[a][:, z]
        s = b.sum()
        return b / s if s > 0 else np.asarray(old_belief)

    def generate_step(self, state, action, words=None):
        """One generative step; `words` = (w1, w2) 32-bit draws (np.random when omitted).  The real environment and
        the device planner share the integer sampling rule, so a test can replay either on the other."""
        a = int(getattr(action, "bin_number", action))
        s = int(state.index)
        if words is None:
            words = np.random.randint(0, 1 << 32, size=2, dtype=np.uint64)
        ns = draw(self.Tc[a, s], words[0])
        z = draw(self.Oc[a, ns], words[1])
        result = StepResult()
        result.action = TabularAction(a)
        result.next_state = TabularState(ns)
        result.observation = TabularObservation(z)
        result.reward = float(self.R[a, s])
        result.is_terminal = bool((self.term_action_mask >> a) & 1) or bool(self.term_state[ns])
        return result, True


def tiger(args=None):
    """The reference's Tiger problem (examples/tiger/tiger_model.py:102-139) as a tabular model: listening (a0) costs
    1 and hears the tiger's door with probability 0.85; opening a door (a1, a2) ends the episode."""
    from .tiger.tiger_model import TigerModel
    return TabularModel(args or {}, TigerModel.get_transition_matrix(), TigerModel.get_observation_matrix(),
                        TigerModel.get_reward_matrix(), TigerModel.get_initial_belief_state(),
                        terminal_actions=(1, 2), name="Tiger")


def random_pomdp(n_states, n_actions, n_obs, seed=0, sparsity=0.5, args=None):
    """A seeded random POMDP (test and benchmark workload)."""
    rng = np.random.RandomState(seed)
    T = rng.gamma(0.5, size=(n_actions, n_states, n_states)) * (rng.rand(n_actions, n_states, n_states) > sparsity)
    T[..., 0] += 1e-3
    O = rng.gamma(0.7, size=(n_actions, n_states, n_obs)) + 1e-3
    R = np.round(rng.randn(n_actions, n_states) * 5.0, 2)
    b0 = rng.dirichlet(np.ones(n_states))
    T /= T.sum(-1, keepdims=True)
    O /= O.sum(-1, keepdims=True)
    return TabularModel(args or {}, T, O, R, b0, name="Random-%d-%d-%d" % (n_states, n_actions, n_obs))
