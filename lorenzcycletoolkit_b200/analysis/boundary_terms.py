"""``BoundaryTerms`` drop-in (reference: ``src/analysis/boundary_terms.py:122-418``)."""
import numpy as np

from ._base import TermBase


class BoundaryTerms(TermBase):
    """BAz, BAe, BKz, BKe, BΦZ, BΦE [W/m^2]: east-west, north-south and bottom-top flux
    differences, evaluated on the device.  No per-level files (as in the reference).  The
    reference applies ``_handle_nans`` to 3-D intermediates here; a box with missing values
    yields NaN boundary terms from the engine and a warning instead."""

    def _boundary(self, name):
        v = self.box_obj.term(name)
        if np.isnan(v).any() and self.app_logger is not None:
            self.app_logger.warning(f"⚠️ {name}: missing values inside the box; boundary term is NaN")
        return self._result(v)

    def calc_baz(self):
        return self._boundary("BAz")

    def calc_bae(self):
        return self._boundary("BAe")

    def calc_bkz(self):
        return self._boundary("BKz")

    def calc_bke(self):
        return self._boundary("BKe")

    def calc_boz(self):
        return self._boundary("BΦZ")

    def calc_boe(self):
        return self._boundary("BΦE")
