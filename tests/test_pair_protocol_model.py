"""Model check (CPU, pure NumPy) of the synchronisation protocol of the persistent split-explicit kernel
(gb-25_b200/csrc/gb25_baro.cu): bands of rows per tile, band edges and tile edges exchanged as (value, sequence)
pairs that the consumer polls for an EXACT sequence number, one buffer per direction, a ring of four row buffers for
the tripolar fold partner, initial values published by the prologue under the launch's base number.

Every (tile, band) is a cooperative task whose work items (one column of one row of one half-substep) run as soon
as their inputs carry the expected number; a randomised, deliberately unfair scheduler interleaves the tasks.  A
buffer that is overwritten before it was consumed, a wrong index or a wrong number shows up as a deadlock (no task can
make progress) or as a result that differs from the plain whole-domain substep loop, which the test also runs.  This is
host-side logic: it needs no GPU.  (The arithmetic itself is checked on the device against k_baro_eta / k_baro_uv.)"""
import numpy as np
import pytest


def _fields(rng, GNx, GNy):
    """Global arrays with a one-cell frame (index [j, i], 1-based interior), random positive metrics."""
    shp = (GNy + 2, GNx + 2)
    f = {k: rng.uniform(0.5, 2.0, shp) for k in ("dyfc", "dxcf", "az", "dxfc", "dycf", "Hfc", "Hcf")}
    for k in ("GU", "GV"):
        f[k] = rng.uniform(-1e-3, 1e-3, shp)
    for k in ("eta", "U", "V"):
        f[k] = rng.uniform(-1.0, 1.0, shp)
    for k in ("dyfc", "dxcf"):               # periodic frame in x for the metrics that are read at i+1
        f[k][:, GNx + 1] = f[k][:, 1]
    return f


def reference(f, GNx, GNy, fold, nsub, weights, dtau, g=9.8):
    """The substep loop on the whole domain: periodic in x, wall at the south, wall or tripolar fold at the north."""
    eta, U, V = f["eta"].copy(), f["U"].copy(), f["V"].copy()
    ae, au, av = (np.zeros_like(eta) for _ in range(3))
    I = np.arange(1, GNx + 1)
    for m in range(nsub):
        en = eta.copy()
        for j in range(1, GNy + 1):
            ie = np.where(I == GNx, 1, I + 1)
            dU = f["dyfc"][j, ie] * U[j, ie] - f["dyfc"][j, I] * U[j, I]
            if j == 1:
                dV = f["dxcf"][j + 1, I] * V[j + 1, I]
            elif j == GNy:
                if fold:
                    dV = f["dxcf"][j + 1, I] * (-V[GNy, GNx - I + 1]) - f["dxcf"][j, I] * V[j, I]
                else:
                    dV = -(f["dxcf"][j, I] * V[j, I])
            else:
                dV = f["dxcf"][j + 1, I] * V[j + 1, I] - f["dxcf"][j, I] * V[j, I]
            en[j, I] = eta[j, I] - dtau * (dU + dV) / f["az"][j, I]
        eta = en
        Un, Vn = U.copy(), V.copy()
        for j in range(1, GNy + 1):
            iw = np.where(I == 1, GNx, I - 1)
            dxe = (eta[j, I] - eta[j, iw]) / f["dxfc"][j, I]
            dye = 0.0 if j == 1 else (eta[j, I] - eta[j - 1, I]) / f["dycf"][j, I]
            Un[j, I] = U[j, I] + dtau * (-g * f["Hfc"][j, I] * dxe + f["GU"][j, I])
            Vn[j, I] = V[j, I] + dtau * (-g * f["Hcf"][j, I] * dye + f["GV"][j, I])
        U, V = Un, Vn
        ae += weights[m] * eta; au += weights[m] * U; av += weights[m] * V
    return ae, au, av


class Pairs:
    """A buffer of (value, sequence) pairs, zero-initialised like the device buffers."""
    def __init__(self, *shape):
        self.v = np.zeros(shape); self.s = np.zeros(shape, dtype=np.int64)

    def put(self, idx, val, seq):
        self.v[idx] = val; self.s[idx] = seq

    def get(self, idx, seq):          # None: not there yet (the consumer keeps polling)
        return self.v[idx] if self.s[idx] == seq else None


class Tile:
    def __init__(self, rx, ry, Rx, Ry, Nx, Ny, nb, f, fold):
        self.rx, self.ry, self.Rx, self.Ry, self.Nx, self.Ny, self.nb = rx, ry, Rx, Ry, Nx, Ny, nb
        self.fold_top = fold and ry == Ry - 1
        i0, j0 = rx * Nx, ry * Ny
        self.c = {k: f[k][j0:j0 + Ny + 2, i0:i0 + Nx + 2].copy() for k in f}      # local arrays with their one-cell frame
        base, extra = divmod(Ny, nb)
        self.rows = []                                                          # band b owns rows rows[b] (1-based, local)
        j = 1
        for b in range(nb):
            r = base + (1 if b < extra else 0)
            self.rows.append(list(range(j, j + r))); j += r
        self.ll_eta, self.ll_v = Pairs(nb, Nx + 1), Pairs(nb, Nx + 1)
        self.inW, self.inE = Pairs(Ny + 1), Pairs(Ny + 1)
        self.inS, self.inN, self.inF = Pairs(Nx + 1), Pairs(Nx + 1), Pairs(4, Nx + 1)
        self.acc = {k: np.zeros((Ny + 2, Nx + 2)) for k in ("eta", "U", "V")}
        self.E = self.W = self.N = self.S = self.F = None                       # neighbouring tiles


def band_task(t, b, seq0, nsub, weights, dtau, rng, g=9.8):
    """Generator: one CTA of the persistent kernel.  Yields True after progress, False when every pending item polls."""
    c, Nx, Ny = t.c, t.Nx, t.Ny
    rows, top, bot = t.rows[b], b == t.nb - 1, b == 0
    xp = t.Rx > 1
    south_wall = t.ry == 0
    north_wall = t.ry == t.Ry - 1 and not t.fold_top
    local_fold = t.fold_top and t.Rx == 1
    # ---- prologue: publish what the neighbours read in their first eta phase, numbered seq0
    for i in range(1, Nx + 1):
        if not bot:
            t.ll_v.put((b, i), c["V"][rows[0], i], seq0)
        if bot and t.S is not None:
            t.S.inN.put(i, c["V"][1, i], seq0)
        if top and t.fold_top and xp:
            t.F.inF.put((seq0 & 3, Nx - i + 1), -c["V"][Ny, i], seq0)
    if xp:
        for j in rows:
            t.W.inE.put(j, c["U"][j, 1], seq0)
    yield True
    for m in range(nsub):
        seq = seq0 + m + 1
        # ================= eta phase
        new = {}
        pending = [(j, i) for j in rows for i in range(1, Nx + 1)]
        while pending:
            rng.shuffle(pending)
            rest, progressed = [], False
            for (j, i) in pending:
                if i < Nx:
                    UE, dyE = c["U"][j, i + 1], c["dyfc"][j, i + 1]
                elif not xp:
                    UE, dyE = c["U"][j, 1], c["dyfc"][j, 1]
                else:
                    UE, dyE = t.inE.get(j, seq - 1), c["dyfc"][j, Nx + 1]
                need_n = not (j == Ny and (north_wall or local_fold))
                VN = 0.0
                if j == Ny and local_fold:
                    VN = -c["V"][Ny, Nx - i + 1]
                elif need_n:
                    if j < rows[-1]:
                        VN = c["V"][j + 1, i]
                    elif not top:
                        VN = t.ll_v.get((b + 1, i), seq - 1)
                    elif t.fold_top:
                        VN = t.inF.get(((seq - 1) & 3, i), seq - 1)
                    else:
                        VN = t.inN.get(i, seq - 1)
                if UE is None or VN is None:
                    rest.append((j, i)); continue
                dU = dyE * UE - c["dyfc"][j, i] * c["U"][j, i]
                if j == 1 and south_wall:
                    dV = c["dxcf"][j + 1, i] * VN
                elif j == Ny and north_wall:
                    dV = -(c["dxcf"][j, i] * c["V"][j, i])
                else:
                    dV = c["dxcf"][j + 1, i] * VN - c["dxcf"][j, i] * c["V"][j, i]
                en = c["eta"][j, i] - dtau * (dU + dV) / c["az"][j, i]
                new[(j, i)] = en
                if j == rows[-1] and not top:
                    t.ll_eta.put((b, i), en, seq)
                if xp and i == Nx:
                    t.E.inW.put(j, en, seq)
                if j == Ny and t.N is not None:
                    t.N.inS.put(i, en, seq)
                progressed = True
            pending = rest
            yield progressed
        for (j, i), v in new.items():            # __syncthreads(): the band's eta is complete
            c["eta"][j, i] = v
        # ================= U, V phase
        newU, newV = {}, {}
        pending = [(j, i) for j in rows for i in range(1, Nx + 1)]
        while pending:
            rng.shuffle(pending)
            rest, progressed = [], False
            for (j, i) in pending:
                e0 = c["eta"][j, i]
                eW = c["eta"][j, i - 1] if i > 1 else (c["eta"][j, Nx] if not xp else t.inW.get(j, seq))
                if j == 1 and south_wall:
                    eS = e0
                elif j > rows[0]:
                    eS = c["eta"][j - 1, i]
                elif not bot:
                    eS = t.ll_eta.get((b - 1, i), seq)
                else:
                    eS = t.inS.get(i, seq)
                if eW is None or eS is None:
                    rest.append((j, i)); continue
                dxe = (e0 - eW) / c["dxfc"][j, i]
                dye = 0.0 if (j == 1 and south_wall) else (e0 - eS) / c["dycf"][j, i]
                Un = c["U"][j, i] + dtau * (-g * c["Hfc"][j, i] * dxe + c["GU"][j, i])
                Vn = c["V"][j, i] + dtau * (-g * c["Hcf"][j, i] * dye + c["GV"][j, i])
                newU[(j, i)], newV[(j, i)] = Un, Vn
                if j == rows[0] and not bot:
                    t.ll_v.put((b, i), Vn, seq)
                if xp and i == 1:
                    t.W.inE.put(j, Un, seq)
                if j == 1 and t.S is not None:
                    t.S.inN.put(i, Vn, seq)
                if j == Ny and t.fold_top and xp:
                    t.F.inF.put((seq & 3, Nx - i + 1), -Vn, seq)
                t.acc["eta"][j, i] += weights[m] * e0
                t.acc["U"][j, i] += weights[m] * Un
                t.acc["V"][j, i] += weights[m] * Vn
                progressed = True
            pending = rest
            yield progressed
        for (j, i) in newU:
            c["U"][j, i] = newU[(j, i)]; c["V"][j, i] = newV[(j, i)]
    # ---- epilogue: the averages become the state
    for j in rows:
        for k in ("eta", "U", "V"):
            c[k][j, 1:Nx + 1] = t.acc[k][j, 1:Nx + 1]


def run_partitioned(f, Rx, Ry, Nx, Ny, nb, fold, nsub, weights, dtau, rng, seq0, starve=None):
    tiles = {(rx, ry): Tile(rx, ry, Rx, Ry, Nx, Ny, nb, f, fold) for rx in range(Rx) for ry in range(Ry)}
    for (rx, ry), t in tiles.items():
        if Rx > 1:
            t.E, t.W = tiles[((rx + 1) % Rx, ry)], tiles[((rx - 1) % Rx, ry)]
        t.N = tiles.get((rx, ry + 1)); t.S = tiles.get((rx, ry - 1))
        if fold and ry == Ry - 1 and Rx > 1:
            t.F = tiles[(Rx - 1 - rx, ry)]
    return tiles


def drive(tiles, launches, nsub, weights, dtau, rng, starve=None):
    """Two launches in a row (sequence numbers continue, buffers are NOT cleared), random unfair scheduling."""
    seq0 = 1
    for _ in range(launches):
        for t in tiles.values():
            for k in t.acc:
                t.acc[k][:] = 0.0
        tasks = {(key, b): band_task(t, b, seq0, nsub, weights, dtau, rng) for key, t in tiles.items() for b in range(t.nb)}
        idle_sweeps = 0
        while tasks:
            keys = list(tasks)
            rng.shuffle(keys)
            if starve is not None and len(keys) > 1 and rng.random() < 0.9:     # one tile runs only now and then
                keys = [k for k in keys if k[0] != starve] or keys
            any_progress = False
            for k in keys[:max(1, len(keys) // 2)]:
                try:
                    any_progress |= bool(next(tasks[k]))
                except StopIteration:
                    del tasks[k]; any_progress = True
            idle_sweeps = 0 if any_progress else idle_sweeps + 1
            assert idle_sweeps < 400, "deadlock: every pending item polls for a number that never arrives"
        seq0 += nsub + 1


@pytest.mark.parametrize("Rx,Ry,fold", [(1, 1, False), (1, 1, True), (2, 1, True), (2, 2, False), (2, 2, True), (4, 2, True), (1, 2, True)])
@pytest.mark.parametrize("starve", [None, "last"])
def test_pair_protocol_matches_the_whole_domain_loop_under_random_scheduling(Rx, Ry, fold, starve):
    Nx, Ny, nb, nsub, dtau = 4, 5, 3, 6, 0.05
    rng = np.random.default_rng(1000 * Rx + 10 * Ry + int(fold))
    weights = rng.uniform(-0.1, 0.3, nsub)
    GNx, GNy = Nx * Rx, Ny * Ry
    f = _fields(rng, GNx, GNy)
    tiles = run_partitioned(f, Rx, Ry, Nx, Ny, nb, fold, nsub, weights, dtau, rng, 1)
    drive(tiles, 2, nsub, weights, dtau, rng, starve=(Rx - 1, Ry - 1) if starve else None)
    # reference: the same two solves on the whole domain (the second starts from the averages of the first)
    g = {k: v.copy() for k, v in f.items()}
    for _ in range(2):
        ae, au, av = reference(g, GNx, GNy, fold, nsub, weights, dtau)
        g["eta"], g["U"], g["V"] = ae, au, av
    for (rx, ry), t in tiles.items():
        sl = (slice(ry * Ny + 1, ry * Ny + Ny + 1), slice(rx * Nx + 1, rx * Nx + Nx + 1))
        for k in ("eta", "U", "V"):
            assert np.array_equal(t.c[k][1:Ny + 1, 1:Nx + 1], g[k][sl]), (k, rx, ry)


def test_a_zero_base_number_would_accept_the_zeroed_buffers():
    """Why the first launch starts at 1: a consumer polling for 0 is satisfied by a buffer that was only memset."""
    p = Pairs(4)
    assert p.get(2, 0) is not None and p.get(2, 1) is None
