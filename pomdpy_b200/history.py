"""History of the real (s, a, o, r, s') steps of an episode (mirrors reference pomdpy/pomdp/history.py).

Host-side bookkeeping only: Agent.run_pomcp calls solver.history.add_entry() / .show() (agent.py:189,202)."""
from .util import print_divider


class HistoryEntry(object):
    def __init__(self, owning_sequence, id):
        self.owning_sequence = owning_sequence
        self.id = id
        self.state = None
        self.action = None
        self.observation = None
        self.reward = 0

    @staticmethod
    def update_history_entry(h, r, a, o, s):
        h.reward, h.action, h.observation, h.state = r, a, o, s


class HistorySequence(object):
    def __init__(self, id):
        self.id = id
        self.entry_sequence = []

    def get_states(self):
        return [e.state for e in self.entry_sequence]

    def get_length(self):
        return len(self.entry_sequence)

    def add_entry(self):
        e = HistoryEntry(self, len(self.entry_sequence))
        self.entry_sequence.append(e)
        return e

    def remove_entry(self, history_entry):
        del self.entry_sequence[history_entry.id]

    def show(self):
        print_divider("medium")
        print("\tDisplaying history sequence")
        for e in self.entry_sequence:
            print_divider("medium")
            print("id: ", e.id)
            print("action: ", e.action.to_string())
            print("observation: ", e.observation.to_string())
            print("next state: ", e.state.to_string())
            print("reward: ", e.reward)


class Histories(object):
    def __init__(self):
        self.sequences_by_id = []

    def get_number_of_sequences(self):
        return len(self.sequences_by_id)

    def create_sequence(self):
        seq = HistorySequence(len(self.sequences_by_id))
        self.sequences_by_id.append(seq)
        return seq

    def delete_sequence(self, sequence):
        last = self.sequences_by_id.pop()
        if last is not sequence:
            self.sequences_by_id[sequence.id] = last
            last.id = sequence.id
