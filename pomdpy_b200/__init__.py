"""pomdpy_b200 -- a B200-native (sm_100a CUDA) POMCP planner behind POMDPy's Agent / Solver / Model plugin API."""
from .agent import Agent, Results
from .model import Model, StepResult

__all__ = ["Agent", "Results", "Model", "StepResult"]
__version__ = "0.1.0"
