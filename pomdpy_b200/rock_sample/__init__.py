from .types import ActionType, GridPosition, RockAction, RockObservation, RockState
from .history_data import PositionAndRockData, RockData
from .rock_model import RockModel, RSCellType

__all__ = ["ActionType", "GridPosition", "RockAction", "RockObservation", "RockState", "PositionAndRockData",
           "RockData", "RockModel", "RSCellType"]
