"""``ConversionTerms`` drop-in (reference: ``src/analysis/conversion_terms.py:103-245``)."""
from ._base import TermBase, G, RD


class ConversionTerms(TermBase):
    """Cz, Ca, Ck, Ce [W/m^2] and their 15 per-level CSV families."""

    def _term1(self):
        return RD / (self.PressureData * G)          # conversion_terms.py:151,177

    def calc_cz(self):
        self._save_vertical_levels(self._term1(), "Cz_1")
        self._save_vertical_levels(self._levels("Cz_2"), "Cz_2")
        return self._volume_term("Cz")

    def calc_ca(self):
        self._save_vertical_levels(self._levels("Ca_1"), "Ca_1")
        self._save_vertical_levels(self._levels("Ca_2"), "Ca_2")
        return self._volume_term("Ca")

    def calc_ck(self):
        for n in ("Ck_1", "Ck_2", "Ck_3", "Ck_4", "Ck_5"):
            self._save_vertical_levels(self._levels(n), n)
        return self._volume_term("Ck", 1.0 / G)

    def calc_ce(self):
        self._save_vertical_levels(self._term1(), "Ce_1")
        self._save_vertical_levels(self._levels("Ce_2"), "Ce_2")
        return self._volume_term("Ce")
