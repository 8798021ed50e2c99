"""``BoundaryTerms`` drop-in (reference: ``src/analysis/boundary_terms.py:122-438``)."""
import numpy as np

from ._base import TermBase, handle_nans, trapz_levels


class BoundaryTerms(TermBase):
    """BAz, BAe, BKz, BKe, BΦZ, BΦE [W/m^2]: east-west, north-south and bottom-top flux
    differences, evaluated on the device.  No per-level files (as in the reference).

    NaN path: the reference passes the three ``(time, level)`` pieces of every term through
    ``_handle_nans`` (:142,157,169 for BAz and the same places of the other five; :420-438) before
    ``.integrate(level)`` / ``isel(level=-1) - isel(level=0)``.  When the engine flags a non-finite
    integrand, the same rule -- linear interpolation along p, then drop the levels that still hold a NaN
    at any time -- is applied here to the per-level pieces the engine returns
    (``lec_set_boundary_levels``) and the term is re-assembled on the host."""

    def _boundary(self, name):
        box = self.box_obj
        v = box.term(name)
        if box.has_nonfinite and np.isnan(v).any():
            pieces = box.boundary_pieces(name)                    # [step][3][level]
            p = self.PressureData
            ew, pe = handle_nans(pieces[:, 0], p)
            ns, pn = handle_nans(pieces[:, 1], p)
            vf, _ = handle_nans(pieces[:, 2], p)
            c1, c2 = box.c12[:, 0], box.c12[:, 1]
            if vf.shape[-1] == 0:
                bt = np.full(len(pieces), np.nan)
            else:
                bt = vf[:, -1] - vf[:, 0]
            v = trapz_levels(ew, pe) * c1 + trapz_levels(ns, pn) * c2 - bt
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
