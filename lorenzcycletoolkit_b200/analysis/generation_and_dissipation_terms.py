"""``GenerationDissipationTerms`` drop-in
(reference: ``src/analysis/generation_and_dissipation_terms.py:122-188``)."""
import numpy as np

from ._base import TermBase, G


def _trapz(y, x, axis=-1):
    """``DataArray.integrate``: trapezoid along ``axis`` (xarray duck_array_ops.trapz)."""
    y = np.moveaxis(np.asarray(y, dtype=np.float64), axis, -1)
    x = np.asarray(x, dtype=np.float64)
    return np.sum((x[1:] - x[:-1]) * 0.5 * (y[..., 1:] + y[..., :-1]), axis=-1)


class GenerationDissipationTerms(TermBase):
    """Gz, Ge [W/m^2] from the diabatic-heating residual Q (thermodynamics.py:76-124), which the
    row kernel evaluates pointwise.

    Dz, De (:154-188; "still needs to be fully implemented and tested" in the reference, never evaluated by a
    bundled case: parity UNPINNED) need a "Friction Velocity" namelist row (box_data.py:185-205; the same
    variable serves as both stress components).  They touch ONE level of u, v and a surface field, so they are
    evaluated here on the host from the box slice -- there is nothing for the GPU to win:
      Dz = AA( [u]_k0 [u*] + [v]_k0 [v*] ) / g        with k0 = isel(level=0), the FIRST level of the sorted axis
      De = AA( u'_k0 u*' + v'_k0 v*' ) / g            (the reference omits the zonal mean of this product and
                                                      would return a (time, lon) array; the area mean of the
                                                      zonal mean is taken here, as for every other eddy term)."""

    def calc_gz(self):
        return self._volume_term("Gz")

    def calc_ge(self):
        return self._volume_term("Ge")

    # -- dissipation from a friction-velocity field -------------------------------------------- #
    def _friction_inputs(self):
        box = self.box_obj
        vl = getattr(box, "variable_list_df", None)
        if vl is None or "Friction Velocity" not in vl.index:
            raise ValueError("Dz / De need a 'Friction Velocity' row in the namelist (none of the reference's "
                             "namelists has one); run with -r to obtain the dissipation terms as residuals")
        from ..utils.box_data import unit_factor
        i0, i1, j0, j1 = box.idx
        data = box.data
        sl = (slice(j0, j1 + 1), slice(i0, i1 + 1))

        def field(row):
            a = np.asarray(data[vl.loc[row]["Variable"]], dtype=np.float64) * unit_factor(vl.loc[row]["Units"], row)
            return a[(Ellipsis,) + sl]
        u, v = field("Eastward Wind Component")[:, 0], field("Northward Wind Component")[:, 0]     # isel(level=0)
        ust = field("Friction Velocity")
        rl = np.asarray(data.rlons, dtype=np.float64)[sl[1]]
        rp = np.asarray(data.rlats, dtype=np.float64)[sl[0]]
        cos = np.asarray(data.coslats, dtype=np.float64)[sl[0]]
        xlen = rl[-1] - rl[0]
        ylen = np.sin(rp[-1]) - np.sin(rp[0])
        za = lambda f: _trapz(f, rl) / xlen
        aa = lambda f: _trapz(f * cos, rp) / ylen
        if ust.ndim == 4:                      # a level-dependent stress: broadcast the level-0 winds against it
            u, v = u[:, None], v[:, None]
        return u, v, ust, za, aa

    def calc_dz(self):
        u, v, ust, za, aa = self._friction_inputs()
        ust_za = za(ust)
        function = aa(za(u) * ust_za + za(v) * ust_za) / G
        if function.ndim == 2:
            self._save_vertical_levels(function, "Dz")
            raise NotImplementedError("a level-dependent friction velocity leaves Dz per level; the reference "
                                      "does not integrate it")
        return self._result(function)

    def calc_de(self):
        u, v, ust, za, aa = self._friction_inputs()
        ust_ze = ust - za(ust)[..., None]
        u_ze, v_ze = u - za(u)[..., None], v - za(v)[..., None]
        function = aa(za(u_ze * ust_ze + v_ze * ust_ze)) / G
        if function.ndim == 2:
            self._save_vertical_levels(function, "De")
            raise NotImplementedError("a level-dependent friction velocity leaves De per level; the reference "
                                      "does not integrate it")
        return self._result(function)
