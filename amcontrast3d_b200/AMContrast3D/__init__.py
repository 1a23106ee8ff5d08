"""Host-side mirror of openpoints/AMContrast3D (Tier 3): same module names, class
names, call signatures and return values as the reference, running on the sm_100a kernels."""
from .MarginContrast import AmbiguityHead, ContrastHead
from .MaskedRefine import RefinementMethod
from .metrics import ambiguity_metrics, ambiguity_summary, posmask_searching
