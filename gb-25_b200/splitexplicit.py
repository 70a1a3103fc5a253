"""Split-explicit free-surface parameters (host side).

Mirrors Oceananigans' ``SplitExplicitFreeSurface(substeps=N)`` materialisation
(``FixedSubstepNumber`` averaging weights; SURVEY.md A.11, confidence M), selected in the
reference at /root/reference/src/baroclinic_instability_model.jl:22.  A Julia host would pass
``model.free_surface.substepping.averaging_weights`` straight through the C ABI instead.
"""
from __future__ import annotations

import numpy as np


def averaging_shape_function(tau, p=2, q=4, r=0.18927):
    tau0 = (p + 2) * (p + q + 2) / ((p + 1) * (p + q + 1))
    return (tau / tau0) ** p * (1 - (tau / tau0) ** q) - r * (tau / tau0)


def averaging_weights(substeps=30):
    """Returns (dtau_frac, weights): fractional barotropic step (2/substeps) and the kept,
    normalised averaging weights (21 of them for substeps=30, the leading ones negative)."""
    N = int(substeps)
    tau = 2.0 * np.arange(1, N + 1) / N
    A = averaging_shape_function(tau)
    lo, hi = 0, N + 1            # Julia searchsortedlast(A, 0; rev=true)
    while lo < hi - 1:
        m = (lo + hi) >> 1
        if A[m - 1] < 0:
            hi = m
        else:
            lo = m
    w = A[:lo] / A[:lo].sum()
    return 2.0 / N, w
