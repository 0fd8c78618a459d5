"""Per-node heuristic data of RockSample: position + rock statistics -> legal / preferred action sets
(mirrors reference examples/rock_sample/rock_position_history.py).

Unlike the reference, a child's rock statistics are a private copy (the reference's ``deep_copy`` shares one
list across the whole tree)."""
import math

from .types import ActionType


class RockData(object):
    def __init__(self, check_count=0, goodness_number=0, chance_good=0.5):
        self.check_count = check_count          # times this rock has been checked
        self.goodness_number = goodness_number  # +1 per good observation, -1 per bad one
        self.chance_good = chance_good          # Bayes estimate that the rock is good

    def copy(self):
        return RockData(self.check_count, self.goodness_number, self.chance_good)

    def to_string(self):
        return " Check count: %s Goodness number: %s Probability that rock is good: %s" % (
            self.check_count, self.goodness_number, self.chance_good)


class PositionAndRockData(object):
    def __init__(self, model, grid_position, all_rock_data, solver):
        self.model = model
        self.solver = solver
        self.grid_position = grid_position
        self.all_rock_data = all_rock_data
        self.legal_actions = self.generate_smart_actions if getattr(model, "preferred_actions", False) \
            else self.generate_legal_actions

    def copy(self):
        return PositionAndRockData(self.model, self.grid_position.copy(), [r.copy() for r in self.all_rock_data],
                                   self.solver)

    shallow_copy = copy
    deep_copy = copy

    def update(self, other_belief):
        self.all_rock_data = other_belief.data.all_rock_data

    def any_good_rocks(self):
        return any(r.goodness_number > 0 for r in self.all_rock_data)

    def create_child(self, rock_action, rock_observation):
        """Statistics of the child reached by (action, observation) (rock_position_history.py:96-137)."""
        nxt = self.copy()
        nxt.grid_position, _ = self.model.make_next_position(self.grid_position.copy(), rock_action.bin_number)
        a = rock_action.bin_number
        if a == ActionType.SAMPLE:
            rd = nxt.all_rock_data[self.model.get_cell_type(self.grid_position)]
            rd.chance_good, rd.check_count, rd.goodness_number = 0.0, 10, -10
        elif a >= ActionType.CHECK:
            k = rock_action.rock_no
            p_ok = float(self.model.get_sensor_correctness_probability(
                self.grid_position.euclidean_distance(self.model.rock_positions[k])))
            rd = nxt.all_rock_data[k]
            rd.check_count += 1
            good, bad = rd.chance_good, 1 - rd.chance_good
            if rock_observation.is_good:
                rd.goodness_number += 1
                good *= p_ok
                bad *= 1 - p_ok
            else:
                rd.goodness_number -= 1
                good *= 1 - p_ok
                bad *= p_ok
            if not (math.fabs(good) < 0.01 and math.fabs(bad) < 0.01):
                rd.chance_good = good / (good + bad)
        return nxt

    def generate_legal_actions(self):
        return self.model.legal_actions_at(self.grid_position)

    def generate_smart_actions(self):
        """The preferred-action heuristic (rock_position_history.py:170-234)."""
        m, pos = self.model, self.grid_position
        code = m.get_cell_type(pos)
        if 0 <= code < m.n_rocks:
            rd = self.all_rock_data[code]
            if rd.chance_good == 1.0 or rd.goodness_number > 0:
                return [ActionType.SAMPLE]
        target = next((k for k, rd in enumerate(self.all_rock_data)
                       if rd.chance_good != 0.0 and rd.goodness_number >= 0), None)
        if target is None:
            return [ActionType.EAST]
        rp, smart = m.rock_positions[target], []
        if rp.i < pos.i:
            smart.append(ActionType.NORTH)
        if rp.i > pos.i:
            smart.append(ActionType.SOUTH)
        if rp.j > pos.j:
            smart.append(ActionType.EAST)
        if rp.j < pos.j:
            smart.append(ActionType.WEST)
        for k, rd in enumerate(self.all_rock_data):
            if rd.chance_good != 0.0 and rd.chance_good != 1.0 and abs(rd.goodness_number) < 2:
                smart.append(ActionType.CHECK + k)
        return smart
