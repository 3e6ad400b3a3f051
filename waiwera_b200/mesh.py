"""Synthetic mesh generator and partitioner producing the arrays the Newton-step path consumes.

The reference builds these arrays with DMPlex (src/mesh.F90:438-664 geometry, :727-802 ghost /
flux-face arrays, :143-171 distribution with overlap 1); that machinery is out of scope
(SURVEY.md section 8), so structured grids are generated directly in the same layout:

  face_cells [nf,2] int32   support cells of each flux face, normal from cell 1 to cell 2
  face_geom  [nf,12]        area, distance(2), distance12, normal(3), gravity_normal,
                            centroid(3), permeability_direction       (src/face.F90:127-133)
  cell_geom  [nc,4]         centroid(3), volume                       (src/cell.F90:93-94)
  rock       [nc,8]         permeability(3), wet/dry conductivity, porosity, density,
                            specific heat                             (src/rock.F90:97-112)

Local cell numbering: owned cells, then partition ghost cells (grouped by owner rank), then
Dirichlet boundary ghost cells.  Setup-time host code (numpy); nothing here is on the hot path.
"""
from dataclasses import dataclass, field

import numpy as np

SEED = 20240917


@dataclass
class Mesh:
    ncell: int            # local cells incl. ghosts
    ninterior: int        # owned + partition ghosts
    nowned: int
    face_cells: np.ndarray
    face_geom: np.ndarray
    cell_geom: np.ndarray
    rock: np.ndarray
    dims: tuple = (0, 0, 0)
    natural: np.ndarray = None        # natural (global) index of each interior local cell
    boundary: dict = field(default_factory=dict)  # ghost_cells, interior_cells (local indices)
    # partition info (None on a serial mesh)
    rank: int = 0
    nranks: int = 1
    first_cell: int = 0               # global (rank-contiguous) index of the first owned cell
    ncell_global: int = 0
    neigh_rank: np.ndarray = None
    send_ptr: np.ndarray = None
    send_idx: np.ndarray = None
    recv_ptr: np.ndarray = None
    recv_idx: np.ndarray = None
    # MINC (add_minc): number of matrix levels and number of fracture (original) cells of the serial mesh
    minc_levels: int = 0
    minc_base: int = 0

    @property
    def nface(self):
        return len(self.face_cells)


def default_rock(n, rng=None, heterogeneous=True):
    """SURVEY 8(d) config 2 rock: k=(1e-13,1e-13,1e-14)*10^U(-0.5,0.5), defaults of src/rock.F90:69-76."""
    rock = np.zeros((n, 8))
    fac = 10.0 ** rng.uniform(-0.5, 0.5, n) if (heterogeneous and rng is not None) else np.ones(n)
    rock[:, 0] = 1e-13 * fac
    rock[:, 1] = 1e-13 * fac
    rock[:, 2] = 1e-14 * fac
    rock[:, 3] = 2.5
    rock[:, 4] = 2.5
    rock[:, 5] = 0.1
    rock[:, 6] = 2200.0
    rock[:, 7] = 1000.0
    return rock


def structured(nx, ny, nz, dx=10.0, dy=None, dz=None, gravity=(0.0, 0.0, -9.8), seed=SEED, heterogeneous=True,
               top_boundary=False):
    """nx*ny*nz box mesh; cell index i + nx*(j + ny*k); k = 0 is the top layer (z = -dz/2)."""
    dy = dx if dy is None else dy
    dz = dx if dz is None else dz
    g = np.asarray(gravity, float)
    rng = np.random.default_rng(seed)
    n = nx * ny * nz
    idx = np.arange(n, dtype=np.int64)
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    cell_geom = np.zeros((n, 4))
    cell_geom[:, 0] = (i + 0.5) * dx
    cell_geom[:, 1] = (j + 0.5) * dy
    cell_geom[:, 2] = -(k + 0.5) * dz
    cell_geom[:, 3] = dx * dy * dz
    fcs, fgs = [], []
    for d, (cond, step, area, dist, normal) in enumerate([
            (i < nx - 1, 1, dy * dz, dx, (1.0, 0.0, 0.0)),
            (j < ny - 1, nx, dx * dz, dy, (0.0, 1.0, 0.0)),
            (k < nz - 1, nx * ny, dx * dy, dz, (0.0, 0.0, -1.0))]):
        c1 = idx[cond]
        c2 = c1 + step
        fg = np.zeros((len(c1), 12))
        fg[:, 0] = area
        fg[:, 1] = 0.5 * dist
        fg[:, 2] = 0.5 * dist
        fg[:, 3] = dist
        fg[:, 4:7] = normal
        fg[:, 7] = float(np.dot(g, normal))
        fg[:, 8:11] = 0.5 * (cell_geom[c1, :3] + cell_geom[c2, :3])
        fg[:, 11] = d + 1
        fcs.append(np.stack([c1, c2], 1))
        fgs.append(fg)
    rock = default_rock(n, rng, heterogeneous)
    ncell = n
    boundary = {}
    if top_boundary:
        # Dirichlet ghost cell above every top-layer cell (src/mesh.F90:583-664, 1069-1264):
        # cell 2 is the ghost, distance = (d1, 0), distance12 = d1, volume 0
        top = idx[k == 0]
        ghosts = n + np.arange(len(top))
        fg = np.zeros((len(top), 12))
        fg[:, 0] = dx * dy
        fg[:, 1] = 0.5 * dz
        fg[:, 2] = 0.0
        fg[:, 3] = 0.5 * dz
        fg[:, 4:7] = (0.0, 0.0, 1.0)
        fg[:, 7] = float(np.dot(g, (0.0, 0.0, 1.0)))
        fg[:, 8:11] = cell_geom[top, :3] + np.array([0.0, 0.0, 0.5 * dz])
        fg[:, 11] = 3
        fcs.append(np.stack([top, ghosts], 1))
        fgs.append(fg)
        gg = np.zeros((len(top), 4))
        gg[:, :3] = fg[:, 8:11]
        cell_geom = np.vstack([cell_geom, gg])
        rock = np.vstack([rock, rock[top]])
        ncell = n + len(top)
        boundary = {"ghost_cells": ghosts.astype(np.int32), "interior_cells": top.astype(np.int32)}
    return Mesh(ncell=ncell, ninterior=n, nowned=n,
                face_cells=np.ascontiguousarray(np.vstack(fcs), dtype=np.int32),
                face_geom=np.ascontiguousarray(np.vstack(fgs)),
                cell_geom=np.ascontiguousarray(cell_geom), rock=np.ascontiguousarray(rock),
                dims=(nx, ny, nz), natural=idx.copy(), boundary=boundary, ncell_global=n)


def add_boundary(mesh, cells, normal, distance, area, direction, gravity=(0.0, 0.0, -9.8)):
    """Dirichlet boundary ghost cells on the outward faces (unit `normal`) of the given interior cells of a serial
    mesh (src/mesh.F90:583-664, 1069-1264): the ghost is cell 2 of the new face, distances (distance, 0),
    distance12 = distance, volume 0, rock copied from the interior cell; `direction` = permeability direction 1..3."""
    assert mesh.nranks == 1 and mesh.nowned == mesh.ninterior
    cells = np.asarray(cells, np.int64)
    nb = len(cells)
    normal = np.asarray(normal, float)
    ghosts = mesh.ncell + np.arange(nb)
    fg = np.zeros((nb, 12))
    fg[:, 0] = area
    fg[:, 1] = distance
    fg[:, 2] = 0.0
    fg[:, 3] = distance
    fg[:, 4:7] = normal
    fg[:, 7] = float(np.dot(np.asarray(gravity, float), normal))
    fg[:, 8:11] = mesh.cell_geom[cells, :3] + distance * normal
    fg[:, 11] = direction
    gg = np.zeros((nb, 4))
    gg[:, :3] = fg[:, 8:11]
    old = mesh.boundary if mesh.boundary else {"ghost_cells": np.zeros(0, np.int32), "interior_cells": np.zeros(0, np.int32)}
    boundary = {"ghost_cells": np.concatenate([old["ghost_cells"], ghosts]).astype(np.int32),
                "interior_cells": np.concatenate([old["interior_cells"], cells]).astype(np.int32)}
    return Mesh(ncell=mesh.ncell + nb, ninterior=mesh.ninterior, nowned=mesh.nowned,
                face_cells=np.ascontiguousarray(np.vstack([mesh.face_cells, np.stack([cells, ghosts], 1)]), dtype=np.int32),
                face_geom=np.ascontiguousarray(np.vstack([mesh.face_geom, fg])),
                cell_geom=np.ascontiguousarray(np.vstack([mesh.cell_geom, gg])),
                rock=np.ascontiguousarray(np.vstack([mesh.rock, mesh.rock[cells]])),
                dims=mesh.dims, natural=mesh.natural, boundary=boundary, ncell_global=mesh.ncell_global)


def box_owner(mesh, parts):
    """owner rank of each interior cell for a px*py*pz box decomposition (rank = a + px*(b + py*c))."""
    nx, ny, nz = mesh.dims
    px, py, pz = parts
    idx = mesh.natural
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    a = np.minimum(i * px // nx, px - 1)
    b = np.minimum(j * py // ny, py - 1)
    c = np.minimum(k * pz // nz, pz - 1)
    return (a + px * (b + py * c)).astype(np.int32)


def coordinate_owner(mesh, nranks):
    """owner rank of each interior cell of ANY serial mesh (unstructured, from ingest): recursive coordinate bisection
    of the cell centroids along the longest extent, sizes balanced by cell counts, ranks split as evenly as nranks allows.
    The cells of a MINC mesh follow their fracture cell (src/mesh.F90:2201-2282: matrix cells are never separated from
    it), and count towards its weight."""
    n = mesh.ninterior
    parent = getattr(mesh, "minc_parent", None)
    if parent is None and mesh.minc_levels > 0 and mesh.minc_base > 0:
        parent = np.tile(np.arange(mesh.minc_base), mesh.minc_levels + 1)
    if parent is None:
        parent = np.arange(n)
    parent = np.asarray(parent)[:n]
    base = np.flatnonzero(parent == np.arange(n))
    weight = np.bincount(parent, minlength=n)[base].astype(float)
    xyz = mesh.cell_geom[base, :3]
    own_base = np.zeros(len(base), np.int32)

    def split(idx, r0, nr):
        if nr == 1 or len(idx) == 0:
            own_base[idx] = r0
            return
        left = nr // 2
        ext = xyz[idx].max(0) - xyz[idx].min(0)
        ax = int(np.argmax(ext))
        order = idx[np.lexsort((idx, xyz[idx, ax]))]
        cw = np.cumsum(weight[order])
        k = int(np.searchsorted(cw, cw[-1] * left / nr, side="left")) + 1
        k = min(max(k, 1), len(order) - 1) if len(order) > 1 else len(order)
        split(order[:k], r0, left)
        split(order[k:], r0 + left, nr - left)

    split(np.arange(len(base)), 0, int(nranks))
    where = np.zeros(n, np.int64)
    where[base] = np.arange(len(base))
    return own_base[where[parent]].astype(np.int32)


def default_parts(nranks):
    return {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(nranks, (1, 1, nranks))


def partition(mesh, owner, rank, nranks):
    """Rank-local mesh with one layer of ghost cells (DMPlexDistribute overlap 1, src/mesh.F90:143-171).

    Owned cells keep ascending natural order; ghost cells are grouped by owner rank, ascending
    natural order inside a group (both sides of a halo pair therefore agree on the message order).
    Faces: every face with at least one owned support cell, in global face order.
    """
    n = mesh.ninterior
    fc = mesh.face_cells
    mine = owner == rank
    owned = np.flatnonzero(mine)
    isb = fc[:, 1] >= n  # boundary faces (cell 2 is a Dirichlet ghost)
    o1 = np.where(fc[:, 0] < n, owner[np.minimum(fc[:, 0], n - 1)], -1)
    o2 = np.where(~isb, owner[np.minimum(fc[:, 1], n - 1)], -1)
    keep = (o1 == rank) | (o2 == rank)
    lf = np.flatnonzero(keep)
    cells = fc[lf]
    # partition ghosts: interior cells on kept faces that are not mine
    cand = cells[cells < n]
    gh = np.unique(cand[owner[cand] != rank])
    gh = gh[np.lexsort((gh, owner[gh]))]
    bnd = np.unique(cells[:, 1][cells[:, 1] >= n])
    order = np.concatenate([owned, gh, bnd])
    g2l = -np.ones(mesh.ncell, np.int64)
    g2l[order] = np.arange(len(order))
    local_fc = g2l[cells].astype(np.int32)
    assert (local_fc >= 0).all()
    counts = np.bincount(owner, minlength=nranks)
    first = int(counts[:rank].sum())
    # halo plan
    neigh = np.unique(owner[gh]) if len(gh) else np.zeros(0, np.int32)
    recv_ptr, recv_idx, send_ptr, send_idx = [0], [], [0], []
    # cells I must send to rank r: my owned cells adjacent (through a face) to a cell owned by r
    a, b = fc[~isb, 0], fc[~isb, 1]
    oa, ob = owner[a], owner[b]
    for r in neigh:
        ridx = gh[owner[gh] == r]
        recv_idx.extend(g2l[ridx])
        recv_ptr.append(len(recv_idx))
        s = np.unique(np.concatenate([a[(oa == rank) & (ob == r)], b[(ob == rank) & (oa == r)]]))
        send_idx.extend(g2l[s])
        send_ptr.append(len(send_idx))
    boundary = {}
    if len(bnd):
        bfaces = cells[cells[:, 1] >= n]
        boundary = {"ghost_cells": g2l[bfaces[:, 1]].astype(np.int32), "interior_cells": g2l[bfaces[:, 0]].astype(np.int32),
                    "global_ghost": bfaces[:, 1].astype(np.int64)}
    return Mesh(ncell=len(order), ninterior=len(owned) + len(gh), nowned=len(owned),
                face_cells=np.ascontiguousarray(local_fc), face_geom=np.ascontiguousarray(mesh.face_geom[lf]),
                cell_geom=np.ascontiguousarray(mesh.cell_geom[order]), rock=np.ascontiguousarray(mesh.rock[order]),
                dims=mesh.dims, natural=order[:len(owned) + len(gh)].copy(), boundary=boundary,
                rank=rank, nranks=nranks, first_cell=first, ncell_global=n,
                neigh_rank=np.asarray(neigh, np.int32), send_ptr=np.asarray(send_ptr, np.int32),
                send_idx=np.asarray(send_idx, np.int32), recv_ptr=np.asarray(recv_ptr, np.int32),
                recv_idx=np.asarray(recv_idx, np.int32), minc_levels=mesh.minc_levels, minc_base=mesh.minc_base)


def hydrostatic_state(mesh, seed=SEED, two_phase_layers=0, thermo_psat=None):
    """SURVEY 8(d) config 2 initial state on the interior cells of a (serial) structured mesh.

    Single-phase liquid: P = 1e5 + 9.8*997*depth (+U(-1e3,1e3)), T = 20 + 0.2*depth (+U(-0.5,0.5)), region 1.
    With two_phase_layers > 0 the top layers are region 4 with P = Psat(T) (thermo_psat: callable T -> P)
    and S_v = U(0.05, 0.4).  Returns unscaled primaries [n,2] and regions [n].
    """
    rng = np.random.default_rng(seed + 1)
    n = mesh.ninterior
    depth = -mesh.cell_geom[:n, 2]
    P = 1.0e5 + 9.8 * 997.0 * depth + rng.uniform(-1e3, 1e3, n)
    T = 20.0 + 0.2 * depth + rng.uniform(-0.5, 0.5, n)
    sv = rng.uniform(0.05, 0.4, n)
    primary = np.stack([P, T], 1)
    region = np.ones(n, np.int32)
    if two_phase_layers > 0:
        nx, ny, nz = mesh.dims
        k = mesh.natural[:n] // (nx * ny)
        tp = k < two_phase_layers
        # hot shallow zone so that Psat(T) is a sensible pressure
        Ttp = 150.0 + 2.0 * depth[tp] / max(depth.max(), 1.0) + rng.uniform(-0.5, 0.5, tp.sum())
        primary[tp, 0] = np.array([thermo_psat(t) for t in Ttp]) if tp.sum() < 50000 else thermo_psat(Ttp)
        primary[tp, 1] = sv[tp]
        region[tp] = 4
    return primary, region


def scale_primaries(primary, region, pressure_scale=1e6, temperature_scale=1e2, partial_pressure_scale=0.0):
    """eos%scale for eos_we / eos_w / eos_wce (src/eos.F90:186-196, src/eos_we.F90:104-109, src/eos_w.F90: pressure
    only; third primary of eos_wce: gas partial pressure, adaptive Pg/P scaling unless a fixed scale is given,
    src/eos_wge.F90:96-110, 639-655)."""
    y = np.array(primary, float, copy=True)
    if y.shape[1] > 2:
        y[:, 2] = y[:, 2] / (partial_pressure_scale if partial_pressure_scale > 0 else y[:, 0])
    y[:, 0] /= pressure_scale
    if y.shape[1] > 1:
        y[:, 1] = np.where(region == 4, y[:, 1], y[:, 1] / temperature_scale)
    return y


def wce_state(mesh, seed=SEED, two_phase_layers=0, thermo_psat=None, pco2=(1.0e4, 5.0e5)):
    """SURVEY 8(d) config 4 state for eos_wce: the eos_we hydrostatic state for the water partial pressure plus a
    CO2 partial pressure U(pco2) per cell; P = P_water + P_CO2.  Returns unscaled primaries [n,3], regions [n]."""
    primary, region = hydrostatic_state(mesh, seed=seed, two_phase_layers=two_phase_layers, thermo_psat=thermo_psat)
    rng = np.random.default_rng(seed + 2)
    pg = rng.uniform(pco2[0], pco2[1], len(primary))
    return np.stack([primary[:, 0] + pg, primary[:, 1], pg], 1), region


# IAPWS-IF97 saturation line (region 4 basic equation, the published closed form): used only to GENERATE
# synthetic initial states near the saturation line (config 4); the EOS kernels carry their own implementation.
_IF97_N = (0.11670521452767e4, -0.72421316703206e6, -0.17073846940092e2, 0.12020824702470e5, -0.32325550322333e7,
           0.14915108613530e2, -0.48232657361591e4, 0.40511340542057e6, -0.23855557567849, 0.65017534844798e3)


def if97_saturation_pressure(t_celsius):
    """saturation pressure (Pa) at temperature (degC), 0..373.9 degC"""
    n = _IF97_N
    T = np.asarray(t_celsius, float) + 273.15
    th = T + n[8] / (T - n[9])
    A = th * th + n[0] * th + n[1]
    B = n[2] * th * th + n[3] * th + n[4]
    Cc = n[5] * th * th + n[6] * th + n[7]
    return (2.0 * Cc / (-B + np.sqrt(B * B - 4.0 * A * Cc))) ** 4 * 1.0e6


def if97_saturation_temperature(p_pa):
    """saturation temperature (degC) at pressure (Pa)"""
    n = _IF97_N
    beta = (np.asarray(p_pa, float) * 1.0e-6) ** 0.25
    E = beta * beta + n[2] * beta + n[5]
    F = n[0] * beta * beta + n[3] * beta + n[6]
    G = n[1] * beta * beta + n[4] * beta + n[7]
    D = 2.0 * G / (-F - np.sqrt(F * F - 4.0 * E * G))
    return 0.5 * (n[9] + D - np.sqrt((n[9] + D) ** 2 - 4.0 * (n[8] + n[9] * D))) - 273.15


def wce_band_state(mesh, seed=SEED, band=(20, 30), pco2=(1.0e4, 5.0e5)):
    """SURVEY 8(d) config 4 state for eos_wce: hydrostatic liquid with a CO2 partial pressure U(pco2) per cell, and a
    band of layers straddling the saturation line (T within +-2 degC of T_sat(P - P_CO2)): in the band, cells are
    alternately single-phase liquid just below the saturation temperature (T = T_sat - U(0, 2), region 1) and
    two-phase with a little vapour (P = P_sat(T) + P_CO2, S_v = U(0.01, 0.1), region 4, T = T_sat + U(0, 2) of the
    hydrostatic water pressure), so that the first Newton updates push cells across the line in both directions.
    Returns unscaled primaries [n,3] = (P, T or S_v, P_CO2) and regions [n]."""
    primary, region = hydrostatic_state(mesh, seed=seed)
    n = len(primary)
    rng = np.random.default_rng(seed + 2)
    pg = rng.uniform(pco2[0], pco2[1], n)
    nx, ny, nz = mesh.dims
    nat = mesh.natural[:n] % (nx * ny * nz)
    k = nat // (nx * ny)
    inb = (k >= band[0]) & (k < band[1])
    pw = primary[:, 0].copy()                     # hydrostatic water pressure
    tsat = if97_saturation_temperature(pw)
    du = rng.uniform(0.0, 2.0, n)
    sv = rng.uniform(0.01, 0.1, n)
    two = inb & (((nat % nx) + (nat // nx) % ny + k) % 2 == 1)
    liq = inb & ~two
    P = pw + pg
    second = primary[:, 1].copy()
    second[liq] = tsat[liq] - du[liq]
    t2 = tsat[two] + du[two]
    P[two] = if97_saturation_pressure(t2) + pg[two]
    second[two] = sv[two]
    region = region.copy()
    region[two] = 4
    return np.stack([P, second, pg], 1), region


def cube_blocks(mesh, size):
    """block-Jacobi sub-domain of every owned cell: size^3 boxes of the structured grid (numbered per rank)"""
    nx, ny, nz = mesh.dims
    idx = mesh.natural[:mesh.nowned]
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    bx, by = -(-nx // size), -(-ny // size)
    key = (i // size) + bx * ((j // size) + by * (k // size))
    _, inv = np.unique(key, return_inverse=True)
    return inv.astype(np.int32)


# ---------------------------------------------------------------- MINC (dual porosity)

def minc_geometry(volumes, spacing, fracture_connection_distance=0.0):
    """MINC 'nested cube' geometry (src/minc.F90:393-544): volumes = (fracture, matrix level 1, ...) fractions summing
    to 1 (normalised if not), spacing = fracture spacing per set of planes (1-3 values).  Returns
    (volume fractions, connection_area[num_levels], connection_distance[num_levels + 1])."""
    vol = np.asarray(volumes, float)
    vol = vol / vol.sum()
    sp = np.atleast_1d(np.asarray(spacing, float))
    nlev = len(vol) - 1

    def proximity(d):  # :393-411
        fout = 1.0 - 2.0 * d / sp
        return 1.0 if (fout < 0).any() else 1.0 - np.prod(fout)

    def proximity_derivative(d):  # :415-433
        fout = 1.0 - 2.0 * d / sp
        if (fout < 0).any():
            return 0.0
        excl = np.array([np.prod(np.delete(fout, i)) for i in range(len(fout))])
        return 2.0 * np.sum(excl / sp)

    def inner_connection_distance(x):  # :437-460 (Pruess 1983)
        u = sp - 2.0 * x
        if len(u) == 1:
            return u[0] / 6.0
        if len(u) == 2:
            return 0.25 * np.prod(u) / np.sum(u)
        pair = sum(u[i] * u[(i + 1) % 3] for i in range(3))
        return 0.3 * np.prod(u) / pair

    vmatrix = 1.0 - vol[0]
    volsum = np.cumsum(vol[1:]) / vmatrix
    dist = np.zeros(nlev + 1)
    area = np.zeros(nlev)
    x = 0.0
    dist[0] = fracture_connection_distance
    area[0] = vmatrix * proximity_derivative(x)
    xr = vol[1] / area[0]
    for i in range(nlev - 1):  # :495-517: bracket, then root of proximity(x) - volsum(i) (Brent, tolerance 1e-8)
        xl = x
        while proximity(xr) - volsum[i] < 0.0:
            xr *= 2.0
        lo, hi = xl, xr
        for _ in range(200):  # bisection to machine precision: same root as the reference's Brent to its 1e-8 tolerance
            mid = 0.5 * (lo + hi)
            if proximity(mid) - volsum[i] < 0.0:
                lo = mid
            else:
                hi = mid
        x = 0.5 * (lo + hi)
        dist[i + 1] = 0.5 * (x - xl)
        area[i + 1] = vmatrix * proximity_derivative(x)
    dist[nlev] = inner_connection_distance(x)
    return vol, area, dist


def add_minc(mesh, volumes=(0.1, 0.9), spacing=(50.0, 50.0, 50.0), matrix_permeability_factor=1.0, cells=None,
             matrix_rock=None, fracture_connection_distance=0.0):
    """MINC mesh on top of a serial mesh, in the reference's numbering (src/mesh.F90:2286-2380): original cells keep
    their index and become the fracture cells, then all level-1 matrix cells (in cell order), then all level-2
    cells, ...; one new flux face per MINC cell with support (level m-1 cell, level m cell), appended after the
    original faces.  Geometry per src/mesh.F90:3120-3160: fracture volume = V*volume(1), level-m volume =
    V*volume(m+1), face area = V*connection_area(m), distance = connection_distance(m:m+1), normal = 0,
    gravity_normal = 0, permeability_direction = 1.
    cells: the MINC zone (default: every cell; the partition helpers minc_owner / minc_cube_blocks need that);
    matrix_rock: 8-double rock record of the matrix cells, or one per zone cell [len(cells), 8] (default: the
    fracture cell's rock with its permeability times matrix_permeability_factor)."""
    assert mesh.nranks == 1 and not mesh.boundary, "add_minc works on a serial mesh without boundary ghosts"
    vol, area, dist = minc_geometry(volumes, spacing, fracture_connection_distance)
    nlev = len(vol) - 1
    n = mesh.ninterior
    zone = np.arange(n, dtype=np.int64) if cells is None else np.asarray(cells, np.int64)
    nz = len(zone)
    V = mesh.cell_geom[zone, 3].copy()
    cg = [mesh.cell_geom[:n].copy()]
    cg[0][zone, 3] = V * vol[0]
    rock = [mesh.rock[:n].copy()]
    fcs, fgs = [mesh.face_cells], [mesh.face_geom]
    for m in range(1, nlev + 1):
        g = mesh.cell_geom[zone].copy()
        g[:, 3] = V * vol[m]
        cg.append(g)
        r = mesh.rock[zone].copy()
        if matrix_rock is not None:
            r[:] = np.asarray(matrix_rock, float)
        else:
            r[:, 0:3] *= matrix_permeability_factor
        rock.append(r)
        fg = np.zeros((nz, 12))
        fg[:, 0] = V * area[m - 1]
        fg[:, 1] = dist[m - 1]
        fg[:, 2] = dist[m]
        fg[:, 3] = dist[m - 1] + dist[m]
        fg[:, 8:11] = mesh.cell_geom[zone, :3]
        fg[:, 11] = 1.0
        inner = zone if m == 1 else n + (m - 2) * nz + np.arange(nz, dtype=np.int64)
        outer = n + (m - 1) * nz + np.arange(nz, dtype=np.int64)
        fcs.append(np.stack([inner, outer], 1).astype(np.int32))
        fgs.append(fg)
    ntot = n + nz * nlev
    natural = np.arange(ntot, dtype=np.int64)
    out = Mesh(ncell=ntot, ninterior=ntot, nowned=ntot, face_cells=np.ascontiguousarray(np.concatenate(fcs).astype(np.int32)),
               face_geom=np.ascontiguousarray(np.concatenate(fgs)), cell_geom=np.ascontiguousarray(np.concatenate(cg)),
               rock=np.ascontiguousarray(np.concatenate(rock)), dims=mesh.dims, natural=natural, ncell_global=ntot,
               minc_levels=nlev, minc_base=n if cells is None else -n)
    return out


def add_minc_zones(mesh, zones):
    """add_minc for several MINC zones with their own geometries (a list of "mesh.minc" entries): zones = [dict(cells,
    volumes, spacing, matrix_rock=None, fracture_connection_distance=0.0)], disjoint cell sets (a cell named twice
    belongs to the later zone).  Numbering (src/mesh.F90:2286-2380, pinned by mesh_test.F90:1505-1612): the original
    cells, then the level-1 cells of ALL zones in natural cell order, then the level-2 cells of the zones that have a
    second level, ...  One zone gives exactly add_minc's arrays.  The result carries minc_parent[ninterior] (fracture
    cell of every cell, itself for the original cells) and minc_level[ninterior]."""
    assert mesh.nranks == 1 and not mesh.boundary, "add_minc_zones works on a serial mesh without boundary ghosts"
    n = mesh.ninterior
    zone_of = np.full(n, -1)
    for k, z in enumerate(zones):
        zone_of[np.asarray(z["cells"], np.int64)] = k
    geo = [minc_geometry(z["volumes"], z["spacing"], z.get("fracture_connection_distance", 0.0)) for z in zones]
    nlev = [len(g[0]) - 1 for g in geo]
    V = mesh.cell_geom[:n, 3].copy()
    cg = [mesh.cell_geom[:n].copy()]
    for k, g in enumerate(geo):
        sel = zone_of == k
        cg[0][sel, 3] = V[sel] * g[0][0]
    rock = [mesh.rock[:n].copy()]
    fcs, fgs = [mesh.face_cells], [mesh.face_geom]
    parent, level = [np.arange(n, dtype=np.int64)], [np.zeros(n, np.int32)]
    index = {0: np.arange(n, dtype=np.int64)}           # index[m][c]: cell number of level m of original cell c
    ntot = n
    for m in range(1, max(nlev + [0]) + 1):
        cells = np.nonzero((zone_of >= 0) & (np.array([nlev[k] if k >= 0 else 0 for k in zone_of]) >= m))[0]
        nz = len(cells)
        g = mesh.cell_geom[cells].copy()
        r = mesh.rock[cells].copy()
        fg = np.zeros((nz, 12))
        for k, (vol, area, dist) in enumerate(geo):
            rows = np.nonzero(zone_of[cells] == k)[0]
            if len(rows) == 0:
                continue
            c = cells[rows]
            g[rows, 3] = V[c] * vol[m]
            mr = zones[k].get("matrix_rock")
            if mr is not None:
                mr = np.asarray(mr, float)
                if mr.ndim == 2:                          # one record per zone cell, in the order of zones[k]["cells"]
                    pos = {int(cc): i for i, cc in enumerate(np.asarray(zones[k]["cells"], np.int64))}
                    r[rows] = mr[[pos[int(cc)] for cc in c]]
                else:
                    r[rows] = mr
            fg[rows, 0] = V[c] * area[m - 1]
            fg[rows, 1] = dist[m - 1]
            fg[rows, 2] = dist[m]
            fg[rows, 3] = dist[m - 1] + dist[m]
        fg[:, 8:11] = mesh.cell_geom[cells, :3]
        fg[:, 11] = 1.0
        cg.append(g)
        rock.append(r)
        idx = np.full(n, -1, np.int64)
        idx[cells] = ntot + np.arange(nz)
        index[m] = idx
        fcs.append(np.stack([index[m - 1][cells], idx[cells]], 1).astype(np.int32))
        fgs.append(fg)
        parent.append(cells)
        level.append(np.full(nz, m, np.int32))
        ntot += nz
    whole = len(zones) == 1 and len(np.unique(np.asarray(zones[0]["cells"]))) == n
    out = Mesh(ncell=ntot, ninterior=ntot, nowned=ntot, face_cells=np.ascontiguousarray(np.concatenate(fcs).astype(np.int32)),
               face_geom=np.ascontiguousarray(np.concatenate(fgs)), cell_geom=np.ascontiguousarray(np.concatenate(cg)),
               rock=np.ascontiguousarray(np.concatenate(rock)), dims=mesh.dims, natural=np.arange(ntot, dtype=np.int64),
               ncell_global=ntot, minc_levels=max(nlev + [0]), minc_base=n if whole else -n)
    out.minc_parent, out.minc_level = np.concatenate(parent), np.concatenate(level)
    return out


def minc_owner(mesh, parts):
    """owner rank of every cell of a MINC mesh: a matrix cell stays with its fracture cell (src/mesh.F90:2201-2282)"""
    n = mesh.minc_base
    assert n > 0, "minc_owner needs a MINC zone covering every cell"
    base = Mesh(ncell=n, ninterior=n, nowned=n, face_cells=None, face_geom=None, cell_geom=None, rock=None,
                dims=mesh.dims, natural=np.arange(n, dtype=np.int64))
    own = box_owner(base, parts)
    return np.tile(own, mesh.minc_levels + 1)


def minc_cube_blocks(mesh, size):
    """block-Jacobi sub-domains of a (possibly partitioned) MINC mesh: size^3 boxes of fracture cells, every
    matrix cell in the sub-domain of its fracture cell"""
    assert mesh.minc_base > 0, "minc_cube_blocks needs a MINC zone covering every cell"
    nx, ny, nz = mesh.dims
    n = nx * ny * nz
    idx = mesh.natural[:mesh.nowned] % n
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    bx, by = -(-nx // size), -(-ny // size)
    key = (i // size) + bx * ((j // size) + by * (k // size))
    _, inv = np.unique(key, return_inverse=True)
    return inv.astype(np.int32)


# ---------------------------------------------------------------- 1-D radial meshes (config 1)

def radial_1d(nr, dr, thickness, outer_boundary=True):
    """1-D radial mesh as the reference builds it from a 2-D (r, z) gmsh strip with "radial": true
    (src/mesh.F90:340-432): cell volume = dr*thickness*2*pi*r_centroid and face area = thickness*2*pi*r_face
    (Pappus), horizontal connections only (gravity_normal = 0).  With outer_boundary a Dirichlet ghost cell
    sits beyond the last cell (distance (d1, 0), src/mesh.F90:583-664).  Cell i spans [i*dr, (i+1)*dr]."""
    i = np.arange(nr)
    rc = (i + 0.5) * dr
    cell_geom = np.zeros((nr + (1 if outer_boundary else 0), 4))
    cell_geom[:nr, 0] = rc
    cell_geom[:nr, 1] = -0.5 * thickness
    cell_geom[:nr, 3] = dr * thickness * 2.0 * np.pi * rc
    rf = (i[:-1] + 1.0) * dr
    fg = np.zeros((nr - 1, 12))
    fg[:, 0] = thickness * 2.0 * np.pi * rf
    fg[:, 1] = 0.5 * dr
    fg[:, 2] = 0.5 * dr
    fg[:, 3] = dr
    fg[:, 4] = 1.0
    fg[:, 8] = rf
    fg[:, 9] = -0.5 * thickness
    fg[:, 11] = 1.0
    fc = np.stack([i[:-1], i[:-1] + 1], 1)
    boundary = {}
    if outer_boundary:
        rb = nr * dr
        bg = np.zeros((1, 12))
        bg[0, 0] = thickness * 2.0 * np.pi * rb
        bg[0, 1] = 0.5 * dr
        bg[0, 2] = 0.0
        bg[0, 3] = 0.5 * dr
        bg[0, 4] = 1.0
        bg[0, 8] = rb
        bg[0, 9] = -0.5 * thickness
        bg[0, 11] = 1.0
        fg = np.concatenate([fg, bg])
        fc = np.concatenate([fc, [[nr - 1, nr]]])
        cell_geom[nr, 0] = rb
        cell_geom[nr, 1] = -0.5 * thickness
        boundary = {"ghost_cells": np.array([nr], np.int32), "interior_cells": np.array([nr - 1], np.int32)}
    rock = default_rock(len(cell_geom), None, heterogeneous=False)
    return Mesh(ncell=len(cell_geom), ninterior=nr, nowned=nr, face_cells=np.ascontiguousarray(fc.astype(np.int32)),
                face_geom=np.ascontiguousarray(fg), cell_geom=np.ascontiguousarray(cell_geom), rock=rock,
                dims=(nr, 1, 1), natural=np.arange(nr, dtype=np.int64), boundary=boundary, ncell_global=nr)
