"""Partition rule of the reference: ``factors(N) -> (Rx, Ry)`` with Rx = 2 Ry and the
reference's special cases (/root/reference/src/sharding_utils.jl:39-62), plus the tile
geometry of /root/reference/sharding/sharded_baroclinic_instability_simulation_run.jl:82-88."""
from __future__ import annotations

import math

_SPECIAL = {4: (2, 2), 16: (4, 4), 512: (32, 32), 6136: (104, 59), 9152: (143, 64),
            9180: (135, 68), 16384: (128, 128)}


def factors(N: int):
    if N == 1:
        return 1, 1          # single device: Partition(1, 1, 1) (not covered by the reference rule)
    if N in _SPECIAL:
        return _SPECIAL[N]
    if N % 2:
        raise ValueError(f"N must be even; got N = {N}")
    half = N // 2
    D = math.isqrt(half)
    if D * D != half:
        raise ValueError(f"N ÷ 2 = {half} is not a perfect square")
    return 2 * D, D


def rank_coords(rank: int, Rx: int, Ry: int):
    """x-fastest rank layout: rank = rx + Rx * ry."""
    return rank % Rx, rank // Rx


def global_size_from_tile(tile_x, tile_y, Rx, Ry, H=8):
    """Reference sizing: the CLI tile includes halos; Nx = tile_x*Rx - 2H (…_run.jl:82-88)."""
    return tile_x * Rx - 2 * H, tile_y * Ry - 2 * H
