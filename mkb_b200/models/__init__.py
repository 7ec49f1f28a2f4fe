from .base import BaseModel
from .kge import ComplEx, DistMult, RotatE, TransE

__all__ = ["BaseModel", "ComplEx", "DistMult", "RotatE", "TransE"]
