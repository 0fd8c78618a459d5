from .tiger_model import TigerAction, TigerModel, TigerObservation, TigerState

__all__ = ["TigerAction", "TigerModel", "TigerObservation", "TigerState"]
