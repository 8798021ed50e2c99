"""Drop-ins for the reference's ``src/analysis`` term classes (same names, constructor and
``calc_*`` methods); every number comes from the CUDA engine via ``BoxData``."""
from .energy_contents import EnergyContents
from .conversion_terms import ConversionTerms
from .boundary_terms import BoundaryTerms
from .generation_and_dissipation_terms import GenerationDissipationTerms

__all__ = ["EnergyContents", "ConversionTerms", "BoundaryTerms", "GenerationDissipationTerms"]
