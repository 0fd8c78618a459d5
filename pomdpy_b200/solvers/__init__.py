from .solver import Solver
from .pomcp import POMCP, greedy_action

__all__ = ["Solver", "POMCP", "greedy_action"]
