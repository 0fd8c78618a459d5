"""RockSample value types (mirror reference examples/rock_sample/{grid_position,rock_state,rock_action,
rock_observation}.py): same constructors, attributes and string forms."""
import math

from ..discrete import DiscreteAction, DiscreteObservation, DiscreteState


class ActionType(object):
    NORTH, EAST, SOUTH, WEST, SAMPLE, CHECK = 0, 1, 2, 3, 4, 5

_NAMES = ("NORTH", "EAST", "SOUTH", "WEST", "SAMPLE")


class GridPosition(object):
    def __init__(self, i=0, j=0):
        self.i = 0 if i is None and j is None else i
        self.j = 0 if i is None and j is None else j

    def __eq__(self, other):
        return self.i == other.i and self.j == other.j

    def __hash__(self):
        return hash((self.i, self.j))

    def copy(self):
        return GridPosition(self.i, self.j)

    def as_list(self):
        return [self.i, self.j]

    def to_string(self):
        return "(%s,%s)" % (self.i, self.j)

    def print_position(self):
        print(self.to_string())

    def manhattan_distance(self, other):
        return float(abs(self.i - other.i) + abs(self.j - other.j))

    def euclidean_distance(self, other):
        return math.sqrt(float(other.j - self.j) ** 2 + float(other.i - self.i) ** 2)


class RockState(DiscreteState):
    """Robot position + one truthy/falsy flag per rock (good / bad)."""

    def __init__(self, grid_position, rock_states):
        self.position = grid_position
        self.rock_states = rock_states

    def copy(self):
        return RockState(self.position, self.rock_states)

    def distance_to(self, other):
        return sum(1 for a, b in zip(self.rock_states, other.rock_states) if bool(a) != bool(b))

    def __eq__(self, other):
        return self.position == other.position and \
            [bool(r) for r in self.rock_states] == [bool(r) for r in other.rock_states]

    def __hash__(self):
        return hash((self.position.i, self.position.j, self.rock_mask()))

    def rock_mask(self):
        return sum(1 << k for k, r in enumerate(self.rock_states) if r)

    def as_list(self):
        return [self.position.i, self.position.j] + [bool(r) for r in self.rock_states]

    def to_string(self):
        return self.position.to_string() + " - " + "".join("1 " if r else "0 " for r in self.rock_states)

    def separate_rocks(self):
        good = [k for k, r in enumerate(self.rock_states) if r]
        bad = [k for k, r in enumerate(self.rock_states) if not r]
        return good, bad


class RockAction(DiscreteAction):
    def __init__(self, bin_number):
        super(RockAction, self).__init__(bin_number)
        self.rock_no = bin_number - ActionType.CHECK if bin_number >= ActionType.CHECK else 0

    def copy(self):
        return RockAction(self.bin_number)

    def to_string(self):
        if self.bin_number >= ActionType.CHECK:
            return "CHECK"
        if 0 <= self.bin_number < len(_NAMES):
            return _NAMES[self.bin_number]
        return "UNDEFINED ACTION"


class RockObservation(DiscreteObservation):
    """is_good / is_empty; equality and hash use is_good only (rock_observation.py:21-25), so the belief tree
    has at most two children per action."""

    def __init__(self, is_good=False, is_empty=None):
        super(RockObservation, self).__init__(0 if is_empty else (1 if is_good else 2))
        self.is_empty = True if is_empty is None else is_empty
        self.is_good = is_good

    def copy(self):
        return RockObservation(self.is_good, self.is_empty)

    def distance_to(self, other):
        return abs(int(bool(self.is_good)) - int(bool(other.is_good)))

    def __eq__(self, other):
        return bool(self.is_good) == bool(other.is_good)

    def __hash__(self):
        return 1 if self.is_good else 0

    def to_string(self):
        if self.is_empty:                       # same strings as rock_observation.py:38-46
            return "EMPTY"
        return "Good" if self.is_good == 1 else ("Bad" if self.is_good == 2 else str(self.is_good))
