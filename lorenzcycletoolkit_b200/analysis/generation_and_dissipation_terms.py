"""``GenerationDissipationTerms`` drop-in
(reference: ``src/analysis/generation_and_dissipation_terms.py:122-188``)."""
from ._base import TermBase


class GenerationDissipationTerms(TermBase):
    """Gz, Ge [W/m^2] from the diabatic-heating residual Q (thermodynamics.py:76-124), which the
    row kernel evaluates pointwise.  Dz/De need a "Friction Velocity" namelist row that no
    bundled namelist has and that the reference marks as not fully implemented (:154,172)."""

    def calc_gz(self):
        return self._volume_term("Gz")

    def calc_ge(self):
        return self._volume_term("Ge")

    def calc_dz(self):
        raise NotImplementedError("Dz needs friction-velocity inputs; run with -r (residuals) as the "
                                  "reference's bundled cases do")

    def calc_de(self):
        raise NotImplementedError("De needs friction-velocity inputs; run with -r (residuals) as the "
                                  "reference's bundled cases do")
