"""Drop-in for ``set_transformer-master/models.py`` under the reference's own class names: ``DeepSet`` (:3-28) and the
generic ``SetTransformer`` with the SAB decoder (:30-44)."""
from .models import DeepSet, SetTransformerSAB as SetTransformer

__all__ = ["DeepSet", "SetTransformer"]
