from .base import BaseModel
from .kge import ComplEx, DistMult, RotatE, TransE, pRotatE

__all__ = ["BaseModel", "ComplEx", "DistMult", "RotatE", "TransE", "pRotatE"]
