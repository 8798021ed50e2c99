"""Drop-ins for the reference's ``src/frameworks`` drivers."""
from .lec_fixed_framework import lec_fixed
from .lec_moving_framework import lec_moving

__all__ = ["lec_fixed", "lec_moving"]
