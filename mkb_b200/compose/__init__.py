from .pipeline import Pipeline
from .trainer import DeviceTrainer

__all__ = ["Pipeline", "DeviceTrainer"]
