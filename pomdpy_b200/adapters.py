"""Device hooks for models that were not written for this package -- first of all the UNMODIFIED RockModel of the
reference (examples/rock_sample/rock_model.py:35-105), so that the reference's own Agent + RockModel can drive the CUDA
solver with a one-line change (``Agent(model, pomdpy_b200.solvers.POMCP)``).

SURVEY.md 8(b): "for the CUDA path a model must additionally expose its tables (a device_spec()-style hook, or
isinstance(model, RockModel) extraction)".  This module is the extraction: it reads the public attributes the reference
model sets in __init__ / initialize() (rock_model.py:45-138) and never touches the reference's code."""
import math
import sys

import numpy as np

ROCK_ATTRS = ("n_rows", "n_cols", "n_rocks", "env_map", "rock_positions", "start_position", "good_rock_reward",
              "bad_rock_penalty", "exit_reward", "illegal_move_penalty", "half_efficiency_distance",
              "actual_rock_states", "unique_rocks_sampled")


def _binomial_threshold(p):
    """np.random.binomial(1.0, p) with the legacy generator draws ONE double U and returns 1 iff U <= this value
    (numpy's inversion algorithm for n = 1; verified against the executed reference)."""
    p = float(p)
    if p >= 1.0:
        return 1.0
    if p <= 0.5:
        return p
    return float(np.exp(np.log(1.0 - (1.0 - p))))


class RockSampleHooks(object):
    """device_spec / world_masks / pack_state / unpack_state for any object with the reference RockModel's attributes."""

    def __init__(self, model):
        missing = [a for a in ROCK_ATTRS if not hasattr(model, a)]
        if missing:
            raise TypeError("not a RockSample model: missing %s" % ", ".join(missing))
        self.m = model
        self._spec = None
        mod = sys.modules.get(type(model).__module__)
        # the model's own state classes (the reference's RockState / GridPosition when the model is the reference's)
        self._RockState = getattr(mod, "RockState", None)
        self._GridPosition = getattr(mod, "GridPosition", None)
        if self._RockState is None or self._GridPosition is None:
            from .rock_sample.types import GridPosition, RockState
            self._RockState, self._GridPosition = RockState, GridPosition

    def cell_index(self, pos):
        return pos.i * self.m.n_cols + pos.j

    def device_spec(self):
        """Immutable tables (rock_model.py:110-138 map, :148-150 sensor): cell codes, start cell, 53-bit sensor
        thresholds and probabilities [cell][rock] computed with the reference's own expressions, signed rewards."""
        if self._spec is not None:
            return self._spec
        m = self.m
        n_cells = m.n_rows * m.n_cols
        if n_cells > 256 or m.n_rocks > 24 or 5 + m.n_rocks > 32:
            raise ValueError("map too large for the packed device state (<= 256 cells, <= 24 rocks)")
        cell = np.array([[int(c) for c in row] for row in m.env_map], dtype=np.int8).reshape(-1)
        thr53 = np.zeros((n_cells, m.n_rocks), dtype=np.uint64)
        prob = np.zeros((n_cells, m.n_rocks), dtype=np.float64)
        for i in range(m.n_rows):
            for j in range(m.n_cols):
                for k, rp in enumerate(m.rock_positions):
                    d = np.sqrt(np.power(rp.j - j, 2.0) + np.power(rp.i - i, 2.0))       # grid_position.py:41-42
                    p = m.get_sensor_correctness_probability(d)                            # rock_model.py:148-150
                    prob[i * m.n_cols + j, k] = p
                    thr53[i * m.n_cols + j, k] = int(math.floor(_binomial_threshold(p) * 9007199254740992.0))
        rewards = np.array([-m.illegal_move_penalty, m.exit_reward, m.good_rock_reward, -m.bad_rock_penalty],
                           dtype=np.float64)
        self._spec = dict(kind="rocksample", rows=int(m.n_rows), cols=int(m.n_cols), n_rocks=int(m.n_rocks),
                          n_actions=5 + int(m.n_rocks), cell=cell, start_cell=self.cell_index(m.start_position),
                          thr53=thr53, prob=prob, rewards=rewards)
        return self._spec

    def world_masks(self):
        """(actual_rock_mask, sampled_rock_mask): the mutable world the device model reads (rock_model.py:263-272)."""
        m = self.m
        actual = sum(1 << k for k, good in enumerate(m.actual_rock_states) if good)   # [] before reset_for_epoch
        sampled = 0
        for k in m.unique_rocks_sampled:
            if 0 <= k < m.n_rocks:
                sampled |= 1 << k
        return actual, sampled

    def pack_state(self, state):
        rocks = sum(1 << k for k, good in enumerate(state.rock_states) if good)
        return self.cell_index(state.position) | (rocks << 8)

    def unpack_state(self, packed):
        cell, rocks = int(packed) & 0xFF, int(packed) >> 8
        m = self.m
        return self._RockState(self._GridPosition(cell // m.n_cols, cell % m.n_cols),
                               [bool((rocks >> k) & 1) for k in range(m.n_rocks)])


def hooks_for(model):
    """The object that answers device_spec() / world_masks() / pack_state() / unpack_state() for `model`: the model
    itself if it has the hooks (this package's models), else an extraction for a reference RockModel."""
    if all(hasattr(model, a) for a in ("device_spec", "world_masks", "pack_state", "unpack_state")):
        return model
    return RockSampleHooks(model)


def spec_from_reference(model):
    """Tables and world masks of an unmodified reference RockModel: (device_spec dict, (actual, sampled))."""
    h = RockSampleHooks(model)
    return h.device_spec(), h.world_masks()
