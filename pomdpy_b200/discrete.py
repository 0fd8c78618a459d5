"""Discrete action / observation / state types and pools (mirrors reference pomdpy/discrete_pomdp/*).

The reference keeps per-node Python dictionaries of DiscreteActionMappingEntry objects
(discrete_action_mapping.py:9-184); here the statistics live in the device tree arena and
``ActionMappingView`` / ``ActionEntryView`` are read-only snapshots with the same attribute names."""
import abc


class DiscreteAction(abc.ABC):
    def __init__(self, bin_number):
        self.bin_number = bin_number

    def __hash__(self):
        return self.bin_number

    def __eq__(self, other):
        return self.bin_number == other.bin_number

    @abc.abstractmethod
    def copy(self):
        pass

    @abc.abstractmethod
    def to_string(self):
        pass

    def print_action(self):
        print(self.to_string())

    def distance_to(self, other):
        return abs(self.bin_number - other.bin_number)


class DiscreteObservation(abc.ABC):
    def __init__(self, bin_number):
        self.bin_number = bin_number

    def __hash__(self):
        return self.bin_number

    def __eq__(self, other):
        return self.bin_number == other.bin_number

    @abc.abstractmethod
    def copy(self):
        pass

    @abc.abstractmethod
    def to_string(self):
        pass

    def print_observation(self):
        print(self.to_string())


class DiscreteState(abc.ABC):
    @abc.abstractmethod
    def copy(self):
        pass

    @abc.abstractmethod
    def as_list(self):
        pass

    @abc.abstractmethod
    def to_string(self):
        pass

    def print_state(self):
        print(self.to_string())


class DiscreteActionPool(object):
    """Factory of per-node action sets (discrete_action_pool.py:6-34)."""

    def __init__(self, model):
        self.all_actions = model.get_all_actions()

    def sample_an_action(self, bin_number):
        return self.all_actions[bin_number]

    def sample_random_action(self):
        import numpy as np
        return self.all_actions[int(np.random.randint(len(self.all_actions)))]

    @staticmethod
    def create_bin_sequence(belief_node):
        return belief_node.data.legal_actions()


class DiscreteObservationPool(object):
    """discrete_observation_pool.py:6-20; child lookup is the device child table, so this is a tag object."""

    def __init__(self, agent):
        self.agent = agent


class ActionEntryView(object):
    """Snapshot of one (belief, action) edge: same fields as DiscreteActionMappingEntry (:113-131)."""

    def __init__(self, mapping, bin_number, visit_count, total_q_value, is_legal):
        self.map = mapping
        self.bin_number = bin_number
        self.visit_count = int(visit_count)
        self.total_q_value = float(total_q_value)
        self.mean_q_value = float(total_q_value) / visit_count if visit_count else 0.0
        self.is_legal = bool(is_legal)
        self.preferred_action = False

    def get_action(self):
        return self.map.pool.sample_an_action(self.bin_number)


class ActionMappingView(object):
    """Snapshot of a node's DiscreteActionMapping (:9-110): entries, total_visit_count, legal bins."""

    def __init__(self, pool, n, total_q, legal_mask):
        self.pool = pool
        self.number_of_bins = len(n)
        self.entries = {a: ActionEntryView(self, a, n[a], total_q[a], (legal_mask >> a) & 1)
                        for a in range(len(n))}
        self.total_visit_count = int(sum(int(x) for x in n))
        self.bin_sequence = [a for a in range(len(n)) if (legal_mask >> a) & 1]

    def get_entry(self, action_bin_number):
        return self.entries.get(action_bin_number)

    def get_visited_entries(self):
        return [e for e in self.entries.values() if e.visit_count > 0]

    def get_all_entries(self):
        return list(self.entries.values())
