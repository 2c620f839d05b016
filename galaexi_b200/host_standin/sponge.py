"""Sponge zone (host, init time): ramp function, damping matrix and base flow. Mirror of /root/reference/src/sponge/sponge.f90:
  * CalcSpongeRamp :259-457  x* = ((x - xStart) . dir) / distance (ramp) or (r - radius) / distance (cylinder), clipped to
                             [0,1]; sigma = min(1, sum over ramps of 6 x*^5 - 15 x*^4 + 10 x*^3); SpongeMat = damping sigma / J^-1
  * InitSponge     :116-244  base flow: constant refstate / exact function / Pruett (starts from the exact function)
The source itself (Sponge :529-588, step 13 of the RHS, dg.f90:419) and the Pruett temporal filter
(pruettdamping.f90:69-92, called once per time step from timedisc_func.f90:357) run on the device."""
from __future__ import annotations

import numpy as np

SPONGESHAPE_RAMP, SPONGESHAPE_CYLINDRICAL = 1, 2


def sponge_mat(case, ramps, damping: float) -> np.ndarray:
    """SpongeMat for ALL elements [e,k,j,i] (zero outside the sponge; the reference keeps a compact list of sponge
    elements, SpongeMap), already divided by sJ like sponge.f90:449-454.

    ramps: list of dicts(shape=1|2, xStart=(3,), distance=float, dir=(3,) [ramp] | radius=float, axis=(3,) [cylinder])."""
    x = case.geo["Elem_xGP"]
    sigma = np.zeros(x.shape[:-1])
    for r in ramps:
        x0 = np.asarray(r.get("xStart", (0.0, 0.0, 0.0)), dtype=np.float64)
        if int(r.get("shape", 1)) == SPONGESHAPE_RAMP:
            v = np.asarray(r.get("dir", (1.0, 0.0, 0.0)), dtype=np.float64)
            v = v / np.sqrt(np.dot(v, v))
            xs = np.sum((x - x0) * v, axis=-1) / float(r["distance"])
        else:
            ax = np.asarray(r.get("axis", (0.0, 0.0, 1.0)), dtype=np.float64)
            rv = x - x0
            rv = rv - np.sum((x - x0) * ax, axis=-1)[..., None] * ax
            xs = (np.sqrt(np.sum(rv * rv, axis=-1)) - float(r["radius"])) / float(r["distance"])
        xs = np.minimum(1.0, np.maximum(0.0, xs))
        sigma = np.minimum(1.0, sigma + 6.0 * xs ** 5 - 15.0 * xs ** 4 + 10.0 * xs ** 3)
    # elements without any node inside a ramp carry no sponge (applySponge, :340-365) -- sigma is zero there anyway
    return damping * sigma / case.geo["sJ"]
