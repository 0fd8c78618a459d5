"""Base class of all solvers (mirrors reference pomdpy/solvers/solver.py:6-20)."""
import abc


class Solver(abc.ABC):
    def __init__(self, agent):
        self.model = agent.model

    @staticmethod
    @abc.abstractmethod
    def reset(agent):
        """Return a new instance of the concrete solver (the Agent's per-epoch factory, agent.py:33,154)."""
