"""Host-side grid products consumed by libgb25cuda (and, in tests, by the CPU oracle).

The library takes every grid product as an input array (metrics at the four horizontal
staggerings, Coriolis parameter at (F,F), vertical faces/centres/spacings, bottom height),
so latitude-longitude and tripolar grids share one device code path.  In the reference these
arrays come from Oceananigans on the host:

* ``simple_latitude_longitude_grid``  <- /root/reference/src/model_utils.jl:56-65
* ``gaussian_islands_tripolar_grid``  <- /root/reference/src/model_utils.jl:134-146
* ``exponential_z_faces``             <- ClimaOcean (un-vendored), SURVEY.md A.2
* ``resolution_to_points``            <- /root/reference/src/model_utils.jl:45-49

Array convention (shared by the C ABI, include/gb25cuda.h): 2-D arrays have shape
``(PY, PX) = (Ny+2Hy+1, Nx+2Hx)`` in NumPy (C order), i.e. x fastest in memory exactly like
a Julia column-major ``(PX, PY)`` array; interior index ``(i, j)`` (1-based) sits at
``[j+Hy-1, i+Hx-1]``.  The extra row holds the ``Ny+1`` face of Bounded-y fields.  Vertical
arrays have length ``PZ = Nz+2Hz+1`` with ``k`` at ``[k+Hz-1]``.
"""
from __future__ import annotations

import dataclasses
import numpy as np

R_EARTH = 6371.0e3          # Oceananigans.defaults / SURVEY.md A.1
OMEGA_EARTH = 7.292115e-5
G_EARTH = 9.80665

TOPO_BOUNDED = 0            # LatitudeLongitudeGrid y topology
TOPO_FOLD = 1               # TripolarGrid y topology (RightConnected, zipper north fold)

METRIC_NAMES = ("dx_cc", "dx_fc", "dx_cf", "dx_ff", "dy_cc", "dy_fc", "dy_cf", "dy_ff",
                "az_cc", "az_fc", "az_cf", "az_ff", "f_ff")
Z_NAMES = ("z_f", "z_c", "dz_c", "dz_f")


def resolution_to_points(resolution):
    """/root/reference/src/model_utils.jl:45-49"""
    nx, ny = 384 / resolution, 192 / resolution
    if nx != int(nx) or ny != int(ny):
        raise ValueError(f"resolution={resolution} does not divide 384 and 192")
    return int(nx), int(ny)


def exponential_z_faces(Nz, depth=4000.0, h=30.0):
    """ClimaOcean.exponential_z_faces(; Nz, depth, h) (confidence M)."""
    k = np.arange(1, Nz + 2, dtype=np.float64)
    L = Nz + 1
    e = (np.exp(k / h) - np.exp(-L / h)) / (1 - np.exp(-L / h))
    e = e - e[0]
    e = e * (-depth / e[-1])
    e[0] = 0.0
    return e[::-1].copy()


@dataclasses.dataclass
class Grid:
    Nx: int
    Ny: int
    Nz: int
    Hx: int
    Hy: int
    Hz: int
    topo_y: int
    metrics: dict            # name -> float64 (PY, PX)
    z: dict                  # name -> float64 (PZ,)
    bottom_height: np.ndarray | None   # float64 (PY, PX) or None (flat, not immersed)
    lam_cc: np.ndarray       # physical longitude / latitude of (C,C) nodes, degrees, (PY, PX)
    phi_cc: np.ndarray
    kind: str = "latlon"
    # tiles of a partitioned grid: does THIS tile own the north wall (Bounded y)?  None = whole domain
    wall_n: bool | None = None

    @property
    def PX(self): return self.Nx + 2 * self.Hx
    @property
    def PY(self): return self.Ny + 2 * self.Hy + 1
    @property
    def PZ(self): return self.Nz + 2 * self.Hz + 1
    @property
    def immersed(self): return self.bottom_height is not None

    def zc_interior(self):
        return self.z["z_c"][self.Hz:self.Hz + self.Nz]

    def field_shape(self, loc):
        """Oceananigans parent shape (x, y, z) for a field at ``loc`` = (lx, ly, lz), 1 = Face.
        Bounded-Face has N+1 points; Periodic / RightConnected Face has N."""
        lx, ly, lz = loc
        ny = self.Ny + (1 if (ly == 1 and self.owns_north_wall) else 0)
        nz = self.Nz + (1 if lz == 1 else 0)
        return (self.Nx + 2 * self.Hx, ny + 2 * self.Hy, nz + 2 * self.Hz)

    @property
    def owns_north_wall(self):
        return (self.topo_y == TOPO_BOUNDED) if self.wall_n is None else bool(self.wall_n)


def _vertical(Nz, Hz, z_faces):
    """Static vertical coordinate with halos by linear extrapolation of the end spacings."""
    zf = np.zeros(Nz + 2 * Hz + 2)
    zf[Hz:Hz + Nz + 1] = z_faces
    dlo, dhi = z_faces[1] - z_faces[0], z_faces[-1] - z_faces[-2]
    for m in range(1, Hz + 1):
        zf[Hz - m] = z_faces[0] - m * dlo
    for m in range(1, Hz + 2):
        zf[Hz + Nz + m] = z_faces[-1] + m * dhi
    zc_full = 0.5 * (zf[:-1] + zf[1:])                    # length PZ
    PZ = Nz + 2 * Hz + 1
    z_f = zf[:PZ].copy()
    z_c = zc_full[:PZ].copy()
    dz_c = (zf[1:] - zf[:-1])[:PZ].copy()
    dz_f = np.empty(PZ)
    dz_f[1:] = z_c[1:] - z_c[:-1]
    dz_f[0] = dz_f[1]
    return {"z_f": z_f, "z_c": z_c, "dz_c": dz_c, "dz_f": dz_f}


def simple_latitude_longitude_grid(Nx, Ny, Nz, halo=(8, 8, 8), latitude=(-80.0, 80.0),
                                   longitude=(0.0, 360.0), depth=4000.0, h=30.0, radius=R_EARTH):
    """LatitudeLongitudeGrid(size, halo, z=exponential_z_faces, latitude=(-80,80), longitude=(0,360)).
    Mirrors /root/reference/src/model_utils.jl:56-65; metric formulas SURVEY.md A.2 (confidence H)."""
    Hx, Hy, Hz = halo
    PX, PY = Nx + 2 * Hx, Ny + 2 * Hy + 1
    dlam = (longitude[1] - longitude[0]) / Nx
    dphi = (latitude[1] - latitude[0]) / Ny
    j = np.arange(1 - Hy, Ny + Hy + 2, dtype=np.float64)      # PY rows
    i = np.arange(1 - Hx, Nx + Hx + 1, dtype=np.float64)
    phif = latitude[0] + (j - 1) * dphi
    phic = phif + dphi / 2
    phic_m1 = phic - dphi
    lamc = longitude[0] + (i - 0.5) * dlam
    dl, dp = np.deg2rad(dlam), np.deg2rad(dphi)
    sind, cosd = (lambda a: np.sin(np.deg2rad(a))), (lambda a: np.cos(np.deg2rad(a)))
    col = lambda v: np.repeat(v[:, None], PX, axis=1)
    dx_c = radius * cosd(phic) * dl
    dx_f = radius * cosd(phif) * dl
    dy = np.full(PY, radius * dp)
    az_c = radius ** 2 * dl * (sind(phif + dphi) - sind(phif))
    az_f = radius ** 2 * dl * (sind(phic) - sind(phic_m1))
    metrics = {
        "dx_cc": col(dx_c), "dx_fc": col(dx_c), "dx_cf": col(dx_f), "dx_ff": col(dx_f),
        "dy_cc": col(dy), "dy_fc": col(dy), "dy_cf": col(dy), "dy_ff": col(dy),
        "az_cc": col(az_c), "az_fc": col(az_c), "az_cf": col(az_f), "az_ff": col(az_f),
        "f_ff": col(2 * OMEGA_EARTH * sind(phif)),
    }
    z = _vertical(Nz, Hz, exponential_z_faces(Nz, depth, h))
    lam_cc = np.repeat(lamc[None, :], PY, axis=0)
    phi_cc = col(phic)
    return Grid(Nx, Ny, Nz, Hx, Hy, Hz, TOPO_BOUNDED, metrics, z, None, lam_cc, phi_cc, "latlon")


# ----------------------------------------------------------------------------------------------
# Tripolar grid.  Oceananigans builds its TripolarGrid on the host (OrthogonalSphericalShellGrids,
# un-vendored); the library only consumes the resulting arrays.  With no Julia here, the benchmark
# and tests use this analytic generator: an orthogonal tripolar grid obtained by composing the
# polar stereographic projection with the Joukowski map  z = Z + a^2/Z,  Z = R exp(i*Lambda),
# R = tan((90-Phi)/2).  Circles |Z| = const > a become confocal ellipses with foci (the two grid
# poles) at z = +-2a, i.e. at latitude `north_poles_latitude`, longitudes `first_pole_longitude`
# and +180; the circle |Z| = a collapses onto the segment between the poles and is the fold line
# (Lambda <-> -Lambda), which Centre row Ny lies on, as in Oceananigans' zipper convention
# (SURVEY.md A.5).  The map is conformal, so the grid is orthogonal and the scale factors are
# analytic.  Halos: periodic in x analytically, fold-copied in the north.
# ----------------------------------------------------------------------------------------------
def fold_index_maps(Nx, Ny, Hx, Hy, loc_x, loc_y):
    """Source indices (1-based i', j') of the north-fold halo rows for a field at (loc_x, loc_y).
    Returns (isrc[i-1], quirk[i-1], jsrc[m-1]) for i=1..Nx, m=1..Hy; ``quirk`` marks the
    Face-x element whose i' wraps past Nx (sign becomes |sign|), SURVEY.md A.5."""
    i = np.arange(1, Nx + 1)
    if loc_x == 0:
        isrc = Nx - i + 1
        quirk = np.zeros(Nx, dtype=bool)
    else:
        isrc = Nx - i + 2
        quirk = isrc > Nx
        isrc = np.where(quirk, isrc - Nx, isrc)
    m = np.arange(1, Hy + 1)
    jsrc = Ny - m if loc_y == 0 else Ny - m + 1
    return isrc, quirk, jsrc


def _fold_fill_2d(a, Nx, Ny, Hx, Hy, loc_x, loc_y):
    """Fill rows j > Ny of a sign-free 2-D array by the zipper map, then re-wrap x halos."""
    isrc, _, jsrc = fold_index_maps(Nx, Ny, Hx, Hy, loc_x, loc_y)
    for m in range(1, Hy + 2):       # one extra row (the PY padding row) gets the same treatment
        js = (Ny - m) if loc_y == 0 else (Ny - m + 1)
        a[Ny + m + Hy - 1, Hx:Hx + Nx] = a[js + Hy - 1, isrc + Hx - 1]
    a[:, :Hx] = a[:, Nx:Nx + Hx]
    a[:, Nx + Hx:] = a[:, Hx:2 * Hx]
    return a


def mtn1(lam, phi):
    """/root/reference/src/model_utils.jl:67-72"""
    return np.exp(-((lam - 70.0) ** 2 + (phi - 55.0) ** 2) / (2 * 5.0 ** 2))


def mtn2(lam, phi):
    """/root/reference/src/model_utils.jl:74-80"""
    return np.exp(-((lam - 250.0) ** 2 + (phi - 55.0) ** 2) / (2 * 5.0 ** 2))


def tripolar_grid(Nx, Ny, Nz, halo=(8, 8, 8), southernmost_latitude=-80.0, north_poles_latitude=55.0,
                  first_pole_longitude=70.0, depth=4000.0, h=30.0, radius=R_EARTH,
                  min_scale=2.0e-2):
    """Analytic orthogonal tripolar grid (see block comment above).  Topology
    (Periodic, RightConnected, Bounded) like TripolarGrid(arch; size, halo, z)."""
    if Nx % 2:
        raise ValueError("tripolar grid needs even Nx")
    Hx, Hy, Hz = halo
    PX, PY = Nx + 2 * Hx, Ny + 2 * Hy + 1
    a = 0.5 * np.tan(np.deg2rad(90.0 - north_poles_latitude) / 2)      # foci at |z| = 2a
    Phi_N = 90.0 - 2 * np.rad2deg(np.arctan(a))                         # logical latitude of the fold
    Phi_S = southernmost_latitude
    dPhi = (Phi_N - Phi_S) / (Ny - 0.5)
    dLam = 360.0 / Nx
    j = np.arange(1 - Hy, Ny + Hy + 2, dtype=np.float64)
    i = np.arange(1 - Hx, Nx + Hx + 1, dtype=np.float64)
    Phif = Phi_S + (j - 1) * dPhi
    Phic = Phif + dPhi / 2
    Lamf = (i - 1) * dLam
    Lamc = Lamf + dLam / 2
    dl, dp = np.deg2rad(dLam), np.deg2rad(dPhi)

    def eval_at(Lam, Phi):
        L, P = np.meshgrid(np.deg2rad(Lam), Phi)                       # (PY, PX)
        Rr = np.tan(np.deg2rad(90.0 - P) / 2)
        Rr = np.maximum(Rr, 1e-6)
        Z = Rr * np.exp(1j * L)
        zz = Z + a * a / Z
        m = np.abs(1 - a * a / (Z * Z))
        m = np.maximum(m, min_scale)                                    # keep the pole cells finite (they are land)
        den = 1 + np.abs(zz) ** 2
        h_lam = radius * 2 * m * Rr / den
        h_phi = radius * m * (1 + Rr ** 2) / den
        phi = 90.0 - 2 * np.rad2deg(np.arctan(np.abs(zz)))
        lam = np.mod(np.rad2deg(np.angle(zz)) + first_pole_longitude, 360.0)
        return h_lam * dl, h_phi * dp, lam, phi

    out = {}
    locs = {"cc": (Lamc, Phic, 0, 0), "fc": (Lamf, Phic, 1, 0), "cf": (Lamc, Phif, 0, 1), "ff": (Lamf, Phif, 1, 1)}
    coords = {}
    for tag, (Lam, Phi, lx, ly) in locs.items():
        dx, dy, lam, phi = eval_at(Lam, Phi)
        for nm, arr in (("dx", dx), ("dy", dy), ("az", dx * dy)):
            out[f"{nm}_{tag}"] = _fold_fill_2d(arr.copy(), Nx, Ny, Hx, Hy, lx, ly)
        coords[tag] = (_fold_fill_2d(lam.copy(), Nx, Ny, Hx, Hy, lx, ly),
                       _fold_fill_2d(phi.copy(), Nx, Ny, Hx, Hy, lx, ly))
    out["f_ff"] = 2 * OMEGA_EARTH * np.sin(np.deg2rad(coords["ff"][1]))
    z = _vertical(Nz, Hz, exponential_z_faces(Nz, depth, h))
    g = Grid(Nx, Ny, Nz, Hx, Hy, Hz, TOPO_FOLD, {k: out[k] for k in METRIC_NAMES}, z, None,
             coords["cc"][0], coords["cc"][1], "tripolar")
    return g


def gaussian_islands_tripolar_grid(Nx, Ny, Nz, halo=(8, 8, 8), **kw):
    """ImmersedBoundaryGrid(TripolarGrid, GridFittedBottom(zb + h (mtn1 + mtn2))).
    Mirrors /root/reference/src/model_utils.jl:134-146 (zb = z[1], h = -zb + 100)."""
    g = tripolar_grid(Nx, Ny, Nz, halo, **kw)
    zb = g.z["z_f"][g.Hz]
    hh = -zb + 100.0
    g.bottom_height = zb + hh * (mtn1(g.lam_cc, g.phi_cc) + mtn2(g.lam_cc, g.phi_cc))
    g.kind = "gaussian_islands"
    return g


def smooth_step(phi):
    """/root/reference/src/model_utils.jl:83-87"""
    return (1 - np.tanh((np.abs(phi) - 40.0) / 5.0)) / 2


def baroclinic_instability_state(grid: Grid):
    """Interior T, S of set_baroclinic_instability_kernel! (/root/reference/src/model_utils.jl:99-110).
    Returns arrays of shape (Nz, Ny, Nx)."""
    zc = grid.zc_interior()[:, None, None]
    phi = grid.phi_cc[grid.Hy:grid.Hy + grid.Ny, grid.Hx:grid.Hx + grid.Nx][None]
    T = (30 + 1e-3 * zc) * smooth_step(phi)
    S = -5e-3 * zc + 0 * phi
    return T, S
