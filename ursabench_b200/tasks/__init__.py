from .decision_making import Decision
from .ood_detection import OODDetection
from .prediction import Prediction
from .task_base import _Task

__all__ = ["Prediction", "OODDetection", "Decision", "_Task"]
