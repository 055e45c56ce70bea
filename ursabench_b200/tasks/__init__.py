from .prediction import Prediction
from .task_base import _Task

__all__ = ["Prediction", "_Task"]
