"""Training loops: Pipeline keeps the reference's learn() contract; DeviceTrainer is the device-resident
step it runs on (single GPU, column-parallel replicas, or a row-sharded entity table)."""
from .pipeline import Pipeline
from .trainer import DeviceTrainer

__all__ = ["Pipeline", "DeviceTrainer"]
