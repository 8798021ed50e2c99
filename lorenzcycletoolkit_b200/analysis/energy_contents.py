"""``EnergyContents`` drop-in (reference: ``src/analysis/energy_contents.py:99-165``)."""
from ._base import TermBase, G


class EnergyContents(TermBase):
    """Az, Ae, Kz, Ke [J/m^2].  The integrands (``AA(T_AE^2)/2 sigma`` ...) are evaluated by the
    CUDA engine; each ``calc_*`` appends the per-level rows and returns the pressure integral."""

    def calc_az(self):
        return self._volume_term("Az")

    def calc_ae(self):
        return self._volume_term("Ae")

    def calc_kz(self):
        return self._volume_term("Kz", 1.0 / (2 * G))

    def calc_ke(self):
        return self._volume_term("Ke", 1.0 / (2 * G))
