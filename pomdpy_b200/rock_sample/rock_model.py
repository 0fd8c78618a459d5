"""RockSample as the reference implements it (examples/rock_sample/rock_model.py), host side.

This class is the Python plugin the Agent drives (the real environment step, epoch resets, statistics) and
the source of the tables the CUDA planner runs on (``device_spec``).  The per-simulation work -- the
generative step, legal-action sets, rollouts -- runs on the device; the Python ``generate_step`` below is the
same function and is what the parity tests compare the device against.

Behaviours kept from the reference: CHECK observations are generated from the REAL
world (``actual_rock_states``) and overwrite the particle's rock bit (A.5); a rock that was really sampled
always reads "not good" without a random draw; observation identity is is_good only (A.6); the goal column is
part of the grid; rewards +10 exit / +10 good sample / -10 other sample / -100 illegal.
Not kept: ``make_reward`` mutating its input state (A.4) -- ``generate_step`` is pure here."""
import math
import os

import numpy as np

from ..discrete import DiscreteActionPool, DiscreteObservationPool
from ..model import Model, StepResult
from ..util import console
from . import maps
from .history_data import PositionAndRockData, RockData
from .types import ActionType, GridPosition, RockAction, RockObservation, RockState

module = "RockModel"

ROCK_CONFIG = {            # reference pomdpy/config/rock_sample_config.json
    "map_file": maps.DEFAULT_MAP,
    "good_rock_reward": 10.0,
    "bad_rock_penalty": 10.0,
    "exit_reward": 10.0,
    "illegal_move_penalty": 100.0,
    "half_efficiency_distance": 20.0,
}


class RSCellType:
    """Rocks are numbered 0, 1, 2, ...; every other cell type is negative."""
    ROCK, EMPTY, GOAL, OBSTACLE = 0, -1, -2, -3


_MOVES = {ActionType.NORTH: (-1, 0), ActionType.EAST: (0, 1), ActionType.SOUTH: (1, 0), ActionType.WEST: (0, -1)}


def binomial_threshold(p):
    """np.random.binomial(1.0, p) of the legacy generator returns 1 iff U <= exp(log(1 - (1 - p))) for
    p > 0.5 (inversion branch with q = 1 - p); verified against the real generator in
    oracle/refharness/validate.py."""
    return math.exp(math.log(1.0 - (1.0 - float(p))))


class RockModel(Model):
    def __init__(self, args, rock_config=None):
        super(RockModel, self).__init__(args)
        cfg = dict(ROCK_CONFIG)
        if rock_config:
            cfg.update(rock_config)
        map_name = args.get("map_file") or os.environ.get("POMDPY_MAP_FILE") or cfg["map_file"]
        self.rock_config = cfg
        self.map_file = map_name
        self.good_rock_reward = cfg["good_rock_reward"]
        self.bad_rock_penalty = cfg["bad_rock_penalty"]
        self.exit_reward = cfg["exit_reward"]
        self.illegal_move_penalty = cfg["illegal_move_penalty"]
        self.half_efficiency_distance = cfg["half_efficiency_distance"]
        if not hasattr(self, "discount"):
            self.discount = 0.95
        if not hasattr(self, "preferred_actions"):
            self.preferred_actions = False

        self.sampled_rock_yet = False
        self.any_good_rocks = False
        self.unique_rocks_sampled = []
        self.num_times_sampled = 0.0
        self.good_samples = 0.0
        self.num_reused_nodes = 0
        self.num_bad_rocks_sampled = 0
        self.num_good_checks = 0
        self.num_bad_checks = 0

        self.n_rows, self.n_cols, self.map_text = maps.load(map_name)
        self.n_rocks = 0
        self.start_position = GridPosition()
        self.rock_positions = []
        self.goal_positions = []
        self.env_map = []
        self.all_rock_data = []
        self.actual_rock_states = []
        self.initialize()

    def initialize(self):
        """Cell codes from the map text (rock_model.py:110-138)."""
        for i, row in enumerate(self.map_text):
            codes = []
            for j, ch in enumerate(row):
                code = RSCellType.EMPTY
                if ch == "o":
                    code = RSCellType.ROCK + self.n_rocks
                    self.rock_positions.append(GridPosition(i, j))
                    self.n_rocks += 1
                elif ch == "G":
                    code = RSCellType.GOAL
                    self.goal_positions.append(GridPosition(i, j))
                elif ch == "S":
                    self.start_position = GridPosition(i, j)
                elif ch == "X":
                    code = RSCellType.OBSTACLE
                codes.append(code)
            self.env_map.append(codes)
        self.num_states = 2 ** self.n_rocks
        self.min_val = -self.illegal_move_penalty / (1 - self.discount)
        self.max_val = self.good_rock_reward * self.n_rocks + self.exit_reward

    # ---- utilities ------------------------------------------------------------------------------------
    def get_cell_type(self, pos):
        return self.env_map[pos.i][pos.j]

    def get_sensor_correctness_probability(self, distance):
        """(1 + 2^(-d / half_efficiency_distance)) / 2, with the reference's numpy expression (:148-150)."""
        return (1 + np.power(2.0, -distance / self.half_efficiency_distance)) * 0.5

    def is_valid_pos(self, pos):
        return 0 <= pos.i < self.n_rows and 0 <= pos.j < self.n_cols and \
            self.get_cell_type(pos) != RSCellType.OBSTACLE

    def is_valid_state(self, rock_state):
        return self.is_valid_pos(rock_state.position)

    def is_valid(self, state):
        if isinstance(state, RockState):
            return self.is_valid_state(state)
        if isinstance(state, GridPosition):
            return self.is_valid_pos(state)
        return False

    def is_terminal(self, rock_state):
        return self.get_cell_type(rock_state.position) == RSCellType.GOAL

    # ---- sampling (numpy global generator, like the reference: pomcp.py:53 seeds it) --------------------
    def sample_rocks(self):
        return self.decode_rocks(int(np.random.randint(0, 1 << self.n_rocks)))

    def decode_rocks(self, value):
        return [value & (1 << k) for k in range(self.n_rocks)]

    def encode_rocks(self, rock_states):
        return sum(1 << k for k in range(self.n_rocks) if rock_states[k])

    def sample_an_init_state(self):
        self.sampled_rock_yet = False
        self.unique_rocks_sampled = []
        return RockState(self.start_position, self.sample_rocks())

    def sample_init_states_packed(self, n):
        """n draws of sample_an_init_state() in the device particle format (uint32 cell | rocks << 8), vectorised:
        the same distribution -- start cell, i.i.d. uniform rock masks (:156-159, :175-176) -- from one numpy call."""
        self.sampled_rock_yet = False
        self.unique_rocks_sampled = []
        rocks = np.random.randint(0, 1 << self.n_rocks, size=int(n)).astype(np.uint32)
        return (rocks << np.uint32(8)) | np.uint32(self.cell_index(self.start_position))

    def sample_position(self):
        return GridPosition(int(np.random.randint(0, self.n_rows)), int(np.random.randint(0, self.n_cols)))

    def sample_state_uninformed(self):
        while True:
            pos = self.sample_position()
            if self.get_cell_type(pos) != RSCellType.OBSTACLE:
                return RockState(pos, self.sample_rocks())

    def sample_state_informed(self, belief):
        return belief.sample_particle()

    # ---- Model interface ----------------------------------------------------------------------------
    def create_observation_pool(self, solver):
        return DiscreteObservationPool(solver)

    def create_action_pool(self):
        return DiscreteActionPool(self)

    def create_root_historical_data(self, solver):
        self.create_new_rock_data()
        return PositionAndRockData(self, self.start_position.copy(), self.all_rock_data, solver)

    def create_new_rock_data(self):
        self.all_rock_data = [RockData() for _ in range(self.n_rocks)]

    def get_all_states(self):
        return None, self.num_states

    def get_all_observations(self):
        return {"Empty": 0, "Bad": 1, "Good": 2}, 3

    def get_all_actions(self):
        return [RockAction(code) for code in range(5 + self.n_rocks)]

    def legal_actions_at(self, pos):
        """Moves into valid cells, SAMPLE iff standing on a rock, every CHECK (:218-247)."""
        legal = []
        for a in range(5 + self.n_rocks):
            if a in _MOVES:
                di, dj = _MOVES[a]
                if not self.is_valid_pos(GridPosition(pos.i + di, pos.j + dj)):
                    continue
            elif a == ActionType.SAMPLE:
                if not self.is_valid_pos(pos) or not 0 <= self.get_cell_type(pos) < self.n_rocks:
                    continue
            legal.append(a)
        return legal

    def get_legal_actions(self, state):
        return self.legal_actions_at(state.position)

    def get_max_undiscounted_return(self):
        return 10 + 10 * sum(1 for r in self.actual_rock_states if r)

    def reset_for_simulation(self):
        self.good_samples = 0.0
        self.num_reused_nodes = 0
        self.num_bad_rocks_sampled = 0
        self.num_bad_checks = 0
        self.num_good_checks = 0

    def reset_for_epoch(self):
        self.actual_rock_states = self.sample_rocks()
        console(2, module, "Actual rock states = " + str(self.actual_rock_states))

    def update(self, step_result):
        if step_result.action.bin_number == ActionType.SAMPLE:
            self.unique_rocks_sampled.append(self.get_cell_type(step_result.next_state.position))
            self.num_times_sampled = 0.0
            self.sampled_rock_yet = True

    def belief_update(self, old_belief, action, observation):
        pass

    # ---- the generative model (:323-465) ------------------------------------------------------------
    def make_next_position(self, pos, action_type):
        is_legal = True
        if action_type >= ActionType.CHECK:
            pass
        elif action_type == ActionType.SAMPLE:
            is_legal = self.is_valid_pos(pos) and 0 <= self.get_cell_type(pos) < self.n_rocks
        else:
            di, dj = _MOVES[action_type]
            moved = GridPosition(pos.i + di, pos.j + dj)
            if self.is_valid_pos(moved):
                pos = moved
            else:
                is_legal = False
        return pos, is_legal

    def make_next_state(self, state, action):
        next_position, is_legal = self.make_next_position(state.position.copy(), action.bin_number)
        if not is_legal:
            return state.copy(), False
        rocks = list(state.rock_states)
        self.any_good_rocks = any(rocks)
        if action.bin_number == ActionType.SAMPLE:
            self.num_times_sampled += 1.0
            rocks[self.get_cell_type(next_position)] = False
        return RockState(next_position, rocks), True

    def make_observation(self, action, next_state):
        if action.bin_number < ActionType.SAMPLE:
            return RockObservation()
        if action.bin_number == ActionType.SAMPLE:
            return RockObservation(False, False)
        if action.rock_no in self.unique_rocks_sampled:
            return RockObservation(False, False)
        observation = bool(self.actual_rock_states[action.rock_no])          # the REAL world (A.5)
        dist = next_state.position.euclidean_distance(self.rock_positions[action.rock_no])
        correct = np.random.binomial(1.0, self.get_sensor_correctness_probability(dist))
        if not correct:
            observation = not observation
        next_state.rock_states[action.rock_no] = bool(observation)          # particle follows the observation
        return RockObservation(observation, False)

    def make_reward(self, state, action, next_state, is_legal):
        if not is_legal:
            return -self.illegal_move_penalty
        if self.is_terminal(next_state):
            return self.exit_reward
        if action.bin_number == ActionType.SAMPLE:
            rock_no = self.get_cell_type(state.position)
            if 0 <= rock_no < self.n_rocks:
                if self.actual_rock_states[rock_no] and state.rock_states[rock_no]:
                    self.good_samples += 1.0
                    return self.good_rock_reward
                self.num_bad_rocks_sampled += 1.0
                return -self.bad_rock_penalty
            return -self.illegal_move_penalty
        return 0

    def generate_reward(self, state, action):
        next_state, is_legal = self.make_next_state(state, action)
        return self.make_reward(state, action, next_state, is_legal)

    def generate_step(self, state, action):
        if action is None:
            print("Tried to generate a step with a null action")
            return None
        if isinstance(action, (int, np.integer)):
            action = RockAction(int(action))
        result = StepResult()
        result.next_state, is_legal = self.make_next_state(state, action)
        result.action = action.copy()
        result.observation = self.make_observation(action, result.next_state)
        result.reward = self.make_reward(state, action, result.next_state, is_legal)
        result.is_terminal = self.is_terminal(result.next_state)
        return result, is_legal

    def generate_particles_uninformed(self, previous_belief, action, obs, n_particles):
        old_pos = previous_belief.state_particles[0].position
        particles = []
        while len(particles) < n_particles:
            result, _ = self.generate_step(RockState(old_pos, self.sample_rocks()), action)
            if obs == result.observation:
                particles.append(result.next_state)
        return particles

    def generate_particles_uninformed_packed(self, root_cell, action, obs, n_particles, max_tries=None):
        """generate_particles_uninformed (rock_model.py:252-261) from the cell of the old root, as packed device
        states; gives up after max_tries draws (the reference loops forever if the observation is impossible)."""
        old_pos = GridPosition(int(root_cell) // self.n_cols, int(root_cell) % self.n_cols)
        out, tries = [], 0
        max_tries = max_tries if max_tries is not None else 50 * int(n_particles)
        while len(out) < n_particles and tries < max_tries:
            tries += 1
            result, _ = self.generate_step(RockState(old_pos, self.sample_rocks()), action)
            if obs == result.observation:
                out.append(self.pack_state(result.next_state))
        return np.array(out, dtype=np.uint32)

    # ---- tables for the CUDA planner ------------------------------------------------------------------
    def cell_index(self, pos):
        return pos.i * self.n_cols + pos.j

    def pack_state(self, state):
        """RockState -> uint32 cell | rocks << 8 (the device particle format)."""
        return self.cell_index(state.position) | (self.encode_rocks(state.rock_states) << 8)

    def unpack_state(self, packed):
        cell, rocks = int(packed) & 0xFF, int(packed) >> 8
        return RockState(GridPosition(cell // self.n_cols, cell % self.n_cols), self.decode_rocks(rocks))

    def device_spec(self):
        """Immutable tables: grid, cell codes, start cell, 53-bit sensor thresholds [cell][rock], signed
        rewards (illegal, exit, good sample, bad sample).  Thresholds are computed with the reference's own
        numpy expressions (:148-150, grid_position.py:41-42) so that device decisions equal np.random.binomial
        decisions for the same uniform."""
        if getattr(self, "_device_spec", None) is not None:       # the tables never change after construction
            return self._device_spec
        n_cells = self.n_rows * self.n_cols
        if n_cells > 256 or self.n_rocks > 24 or 5 + self.n_rocks > 32:
            raise ValueError("map too large for the packed device state (<=256 cells, <=24 rocks)")
        cell = np.array(self.env_map, dtype=np.int8).reshape(-1)
        thr53 = np.zeros((n_cells, self.n_rocks), dtype=np.uint64)
        prob = np.zeros((n_cells, self.n_rocks), dtype=np.float64)
        for i in range(self.n_rows):
            for j in range(self.n_cols):
                for k, rp in enumerate(self.rock_positions):
                    d = np.sqrt(np.power(rp.j - j, 2.0) + np.power(rp.i - i, 2.0))
                    p = self.get_sensor_correctness_probability(d)
                    prob[i * self.n_cols + j, k] = p
                    thr53[i * self.n_cols + j, k] = int(math.floor(binomial_threshold(p) * 9007199254740992.0))
        rewards = np.array([-self.illegal_move_penalty, self.exit_reward, self.good_rock_reward,
                            -self.bad_rock_penalty], dtype=np.float64)
        self._device_spec = dict(kind="rocksample", rows=self.n_rows, cols=self.n_cols, n_rocks=self.n_rocks,
                                 n_actions=5 + self.n_rocks, cell=cell, start_cell=self.cell_index(self.start_position),
                                 thr53=thr53, prob=prob, rewards=rewards)
        return self._device_spec

    def world_masks(self):
        """(actual_rock_mask, sampled_rock_mask): the mutable world the device model reads (A.5)."""
        actual = sum(1 << k for k in range(self.n_rocks) if self.actual_rock_states[k])
        sampled = 0
        for k in self.unique_rocks_sampled:
            if 0 <= k < self.n_rocks:
                sampled |= 1 << k
        return actual, sampled

    def draw_env(self):
        for row in self.env_map:
            print(" ".join("G" if c == RSCellType.GOAL else "X" if c == RSCellType.OBSTACLE else
                           " . " if c == RSCellType.EMPTY else format(c, "x") for c in row))
            print()
