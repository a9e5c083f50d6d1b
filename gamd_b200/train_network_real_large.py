"""``ParticleNetLightning`` of code/water/train_network_real_large.py (DFT-water, dynamic box) bound to its module
constants (:24-26: CUTOFF_RADIUS = 3.4 in the shipped drivers is passed by the caller)."""
from .force_field import DynamicBoxForceFacade as ParticleNetLightning  # noqa: F401
from .force_field import create_water_bond  # noqa: F401
