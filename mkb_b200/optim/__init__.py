from .adam import DenseAdam

__all__ = ["DenseAdam"]
