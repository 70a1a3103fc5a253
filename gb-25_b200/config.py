"""Physics/fidelity configuration shared by the host mirror, the C ABI and (in tests) the oracle.

Defaults are the GB-25 benchmark model (/root/reference/src/baroclinic_instability_model.jl:17-40):
SplitExplicitFreeSurface(substeps=30), SeawaterBuoyancy(TEOS10), closure=nothing,
HydrostaticSphericalCoriolis, WENOVectorInvariant(order=5), WENO(order=5), halo (8,8,8).
The integer flags are the named fidelity decisions of the uncertainty register (DESIGN.md,
SURVEY.md A.15): they exist because the Oceananigans source is not available here."""
from __future__ import annotations

import dataclasses

from .grids import G_EARTH


@dataclasses.dataclass
class PhysicsConfig:
    substeps: int = 30
    g: float = G_EARTH
    rho0: float = 1020.0            # TEOS-10 reference density
    chi: float = 0.1                # QuasiAdamsBashforth2 χ
    weno_eps: float = 1e-8
    coriolis_scheme: int = 1        # U2: 1 = ActiveCellEnstrophyConserving, 0 = EnstrophyConserving
    fold_variant: int = 0           # U1: 1 = also overwrite the redundant half of row Ny
    south_inactive: int = 1         # U4: tripolar: cells south of j=1 are outside the domain
    cond_diff: int = 1              # U15: immersed-aware differences in ζ and ∇p
    eos_r0: int = 0                 # U8: include r0(z) in ρ′
    # closure.  The reference default is `closure = nothing`; the commented alternative at
    # src/baroclinic_instability_model.jl:31 is VerticalScalarDiffusivity(VerticallyImplicitTimeDiscretization(),
    # κ=1e-5, ν=1e-4): closure = 2 with these coefficients.  closure = 1 is the explicit time discretization.
    closure: int = 0
    kappa: float = 1e-5
    nu: float = 1e-4
    # test-side only: which smoothness-indicator form the CPU checker evaluates (0 = expanded, the recalled
    # reference form; 1 = sum of squares, the form libgb25cuda always uses — DESIGN.md deviation D1)
    oracle_beta_form: int = 0
