"""Reads a Waiwera input (JSON + gmsh mesh) into the array contract of the C ABI (SURVEY.md section 8 row f-2,
Appendix A): cell and face geometry as src/mesh.F90:438-664 computes it (finite-volume centroids and areas, 2-D
meshes with a thickness or as solids of revolution, normal distances with the non-orthogonality correction,
gravity normal, permeability direction), Dirichlet boundary ghost cells from the "boundaries" face specifications
(src/mesh.F90:1631-1813, 583-664), rock records from the rock types (src/rock_setup.F90), initial primaries /
regions, fixed-rate sources, tracers and the EOS / curve parameters.

Setup-time host code (numpy), nothing here is on the hot path.  Also read: ExodusII meshes (netCDF classic, and
netCDF-4 through h5lite) and MULgraph geometry files, zones, MINC ("mesh.minc"), source controls (deliverability,
recharge, limiters, separators, tables in time), restarts from HDF5 output files.  Not covered: source networks.
Cell order = element order of the mesh file (DMPlex numbering of a serial mesh);
faces are ordered by (cell 1, cell 2), which is not DMPlex's face numbering -- only the rounding of the inflow
sums depends on it."""
import json
import os
import struct

import numpy as np

from . import mesh as wmesh

# gmsh element type -> (dimension, number of nodes)
_GMSH = {1: (1, 2), 2: (2, 3), 3: (2, 4), 4: (3, 4), 5: (3, 8), 6: (3, 6), 7: (3, 5), 15: (0, 1)}
# local faces of the 3-D elements (gmsh node ordering)
_FACES3 = {4: [(0, 2, 1), (0, 1, 3), (0, 3, 2), (1, 2, 3)],
           5: [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)],
           6: [(0, 2, 1), (3, 4, 5), (0, 1, 4, 3), (1, 2, 5, 4), (2, 0, 3, 5)],
           7: [(0, 3, 2, 1), (0, 1, 4), (1, 2, 4), (2, 3, 4), (3, 0, 4)]}


def _dmplex_cell_order(elems):
    """DMPlex numbers the prism (wedge) cells of a hybrid mesh after all the others, each group in file order (pinned by
    the reference's initial_test.F90:619-835 on hybrid10.msh, wedges first in the file, and by the cell indices of its
    minc_3d_refined deck): stable reordering of the elements of the highest dimension; the others keep their places"""
    dim = max(_GMSH[t][0] for t, _ in elems)
    top = [k for k, (t, _) in enumerate(elems) if _GMSH[t][0] == dim]
    new = [k for k in top if elems[k][0] != 6] + [k for k in top if elems[k][0] == 6]
    out = list(elems)
    for slot, k in zip(top, new):
        out[slot] = elems[k]
    return out


def read_gmsh(path, dmplex_order=True):
    """gmsh MSH 2.2 file, ASCII or binary -> (nodes [n,3], list of (element type, node indices 0-based)), cells in the
    order DMPlex numbers them (file order, wedges last: _dmplex_cell_order) or, dmplex_order=False, in the file's own"""
    data = open(path, "rb").read()
    pos = data.index(b"$MeshFormat") + len(b"$MeshFormat")
    end = data.index(b"\n", pos + 1)
    version, ftype, dsize = data[pos:end].split()
    assert version.startswith(b"2"), "only gmsh format 2.x is supported"
    binary = int(ftype) == 1
    pos = end + 1
    endian = "<"
    if binary:
        if struct.unpack("<i", data[pos:pos + 4])[0] != 1:
            endian = ">"
        pos += 4
    pos = data.index(b"$Nodes", pos) + len(b"$Nodes")
    end = data.index(b"\n", pos + 1)
    nn = int(data[pos:end])
    pos = end + 1
    ids, xyz = np.zeros(nn, np.int64), np.zeros((nn, 3))
    if binary:
        rec = np.dtype([("id", endian + "i4"), ("x", endian + "f8", 3)])
        a = np.frombuffer(data, rec, nn, pos)
        ids, xyz = a["id"].astype(np.int64), a["x"].copy()
        pos += nn * rec.itemsize
    else:
        for k in range(nn):
            end = data.index(b"\n", pos)
            f = data[pos:end].split()
            ids[k], xyz[k] = int(f[0]), [float(v) for v in f[1:4]]
            pos = end + 1
    index = {int(i): k for k, i in enumerate(ids)}
    pos = data.index(b"$Elements", pos) + len(b"$Elements")
    end = data.index(b"\n", pos + 1)
    ne = int(data[pos:end])
    pos = end + 1
    elems = []
    if binary:
        done = 0
        while done < ne:
            etype, nfollow, ntags = struct.unpack(endian + "3i", data[pos:pos + 12])
            pos += 12
            nnode = _GMSH[etype][1]
            for _ in range(nfollow):
                vals = struct.unpack(endian + "%di" % (1 + ntags + nnode), data[pos:pos + 4 * (1 + ntags + nnode)])
                pos += 4 * (1 + ntags + nnode)
                elems.append((etype, [index[v] for v in vals[1 + ntags:]]))
            done += nfollow
    else:
        for _ in range(ne):
            end = data.index(b"\n", pos)
            f = [int(v) for v in data[pos:end].split()]
            pos = end + 1
            etype, ntags = f[1], f[2]
            elems.append((etype, [index[v] for v in f[3 + ntags:3 + ntags + _GMSH[etype][1]]]))
    return xyz, (_dmplex_cell_order(elems) if dmplex_order and elems else elems)


# ExodusII element names -> the gmsh type codes build_mesh works with (the node orderings of these element types
# are the same in both formats)
_EXO = {"HEX": 5, "HEX8": 5, "HEXAHEDRON": 5, "WEDGE": 6, "WEDGE6": 6, "TETRA": 4, "TETRA4": 4, "TET4": 4, "PYRAMID": 7, "PYRAMID5": 7,
        "QUAD": 3, "QUAD4": 3, "SHELL": 3, "SHELL4": 3, "TRI": 2, "TRI3": 2, "TRIANGLE": 2}


def read_exodus(path):
    """ExodusII mesh -> (nodes [n,3], list of (gmsh element type, node indices 0-based)), element blocks in the order
    DMPlexCreateExodus numbers the cells (file order, wedge blocks last).  Both containers the reference's meshes come in are read without a netCDF
    library: the netCDF classic format (its unit-test meshes) with scipy's pure-Python reader, netCDF-4 (= HDF5: its
    benchmark meshes) with h5lite."""
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic == b"\x89HDF\r\n\x1a\n":
        from . import h5lite
        h = h5lite.H5File(path)
        var = lambda name: h[name] if name in h else None
        attr = lambda name, a: h.attrs(name).get(a)
        nblk = h.shape("num_el_blk")[0] if "num_el_blk" in h else 0
        nn = h.shape("num_nodes")[0]
    elif magic[:3] == b"CDF":
        from scipy.io import netcdf_file
        f = netcdf_file(path, "r", mmap=False)
        var = lambda name: np.array(f.variables[name][:]) if name in f.variables else None
        attr = lambda name, a: getattr(f.variables[name], a)
        nblk = f.dimensions.get("num_el_blk", 0)
        nn = f.dimensions["num_nodes"]
    else:
        raise ValueError("%s: neither a netCDF classic nor an HDF5-based ExodusII file" % path)
    xyz = np.zeros((nn, 3))
    c = var("coord")
    if c is not None:
        xyz[:, :c.shape[0]] = np.asarray(c, float).T
    else:
        for k, name in enumerate(("coordx", "coordy", "coordz")):
            c = var(name)
            if c is not None:
                xyz[:, k] = np.asarray(c, float)
    blocks = []
    for b in range(1, nblk + 1):
        et = attr("connect%d" % b, "elem_type")
        et = (et.decode() if isinstance(et, bytes) else str(et)).strip().upper()
        assert et in _EXO, "ExodusII element type %r is not supported" % et
        blocks.append((_EXO[et], np.asarray(var("connect%d" % b), np.int64) - 1))
    # DMPlex numbers the prism (wedge) blocks of a hybrid mesh after all the others, each group in file order: the cell
    # indices of the reference's minc_3d_refined deck (wedge block first in the file, top-layer hexahedra 0..92, wedges
    # 465..479) only fit this order
    elems = [(t, [int(i) for i in row]) for t, con in blocks for row in con]
    return xyz, _dmplex_cell_order(elems)


def read_mulgraph(path):
    """MULgraph geometry file (the fixed-format `g*.dat` files of TOUGH2 / PyTOUGH that the reference's benchmark meshes
    are generated from: VERTICES name x y; GRID columns `name centre-flag nnodes` + their vertex names; LAYERS name
    bottom centre) -> (nodes [n,3], elements) of the layered mesh: one prism (hexahedron for 4-sided, wedge for 3-sided
    columns) per column and layer, cells ordered layer by layer from the top, columns in file order -- the numbering of
    the meshes the reference ships next to these files.  Surfaces that cut columns (SURFACE section) are not read."""
    lines = open(path).read().split("\n")
    sec = {}
    cur = None
    for ln in lines[1:]:
        key = ln.strip()
        if key in ("VERTICES", "GRID", "CONNECTIONS", "LAYERS", "SURFACE", "SURFA", "WELLS"):
            cur = key
            sec[cur] = []
        elif cur is not None and ln.strip():
            sec[cur].append(ln)
    assert not sec.get("SURFACE") and not sec.get("SURFA"), "MULgraph surfaces are not supported"
    vert = {}
    for ln in sec["VERTICES"]:
        vert[ln[:3]] = (float(ln[3:13]), float(ln[13:23]))
    cols = []
    it = iter(sec["GRID"])
    for ln in it:
        nn = int(ln[4:6])
        cols.append([next(it)[:3] for _ in range(nn)])
    layers = [(float(ln[3:13]), float(ln[13:23])) for ln in sec["LAYERS"]]     # (bottom, centre); the first is the surface
    tops = [layers[0][0]] + [b for b, _ in layers[1:-1]]
    bottoms = [b for b, _ in layers[1:]]
    levels = [layers[0][0]] + bottoms
    names = sorted(vert)
    index = {(nme, k): i for i, (k, nme) in enumerate((k, nme) for k in range(len(levels)) for nme in names)}
    nodes = np.array([[vert[nme][0], vert[nme][1], levels[k]] for k in range(len(levels)) for nme in names])
    elems = []
    for k in range(len(bottoms)):                      # layer k: top = level k, bottom = level k + 1
        for col in cols:
            poly = np.array([vert[v] for v in col])
            area2 = np.sum(poly[:, 0] * np.roll(poly[:, 1], -1) - np.roll(poly[:, 0], -1) * poly[:, 1])
            ring = col if area2 > 0 else col[::-1]     # counter-clockwise seen from above
            assert len(ring) in (3, 4), "columns with %d sides are not supported" % len(ring)
            elems.append((5 if len(ring) == 4 else 6, [index[(v, k + 1)] for v in ring] + [index[(v, k)] for v in ring]))
    return nodes, elems


def read_mesh(path):
    """mesh file by extension: gmsh MSH 2.2 (.msh), ExodusII (.exo / .e / .ex2) or MULgraph geometry (.dat)"""
    low = path.lower()
    if low.endswith((".exo", ".e", ".ex2", ".exii")):
        return read_exodus(path)
    if low.endswith(".dat"):
        return read_mulgraph(path)
    return read_gmsh(path)


def _polygon_geometry(p):
    """area-weighted centroid and area of a planar polygon in 2-D (DMPlexComputeGeometryFVM for a 2-D cell)"""
    x, y = p[:, 0], p[:, 1]
    x1, y1 = np.roll(x, -1), np.roll(y, -1)
    cr = x * y1 - x1 * y
    a = 0.5 * cr.sum()
    cx, cy = ((x + x1) * cr).sum() / (6.0 * a), ((y + y1) * cr).sum() / (6.0 * a)
    return np.array([cx, cy]), abs(a)


def _face3_geometry(p):
    """centroid, area, unit normal of a (nearly) planar polygon in 3-D: fan of triangles about the vertex mean"""
    c0 = p.mean(0)
    an, cen = np.zeros(3), np.zeros(3)
    atot = 0.0
    for k in range(len(p)):
        a, b = p[k], p[(k + 1) % len(p)]
        n = 0.5 * np.cross(a - c0, b - c0)
        an += n
        ar = np.linalg.norm(n)
        cen += ar * (c0 + a + b) / 3.0
        atot += ar
    area = np.linalg.norm(an)
    return cen / atot, area, an / area


def _cell3_geometry(nodes, faces):
    """centroid and volume of a polyhedron from its faces: pyramids over the face triangles about the vertex mean"""
    pts = np.unique(np.concatenate(faces))
    c0 = nodes[pts].mean(0)
    vol, cen = 0.0, np.zeros(3)
    for f in faces:
        p = nodes[list(f)]
        fc = p.mean(0)
        for k in range(len(p)):
            a, b = p[k], p[(k + 1) % len(p)]
            v = abs(np.dot(np.cross(a - c0, b - c0), fc - c0)) / 6.0
            vol += v
            cen += v * (c0 + a + b + fc) / 4.0
    return cen / vol, vol


def _seq_mean(p):
    """mean over axis 1 of [n, k, 3] with the additions in index order (what ndarray.mean does for one small polygon)"""
    s = p[:, 0].copy()
    for k in range(1, p.shape[1]):
        s = s + p[:, k]
    return s / p.shape[1]


def _build_mesh3(nodes, cells, g, permeability_angle):
    """the 3-D part of build_mesh with numpy array operations over all cells / faces at once (the loop version takes
    20 minutes per million cells): the same formulas in the same order -- cell centroid and volume from pyramids over
    the face triangles about the vertex mean, face centroid / area / normal from the triangle fan, faces matched through
    their sorted node numbers -- and the same ordering conventions: face (c1, c2) with c1 < c2 oriented from c1, interior
    faces sorted by (c1, c2), exterior faces in the order cells and their local faces are walked.  Returns
    (cell_geom, face_cells, face_geom, exterior)."""
    nodes = np.asarray(nodes, float)
    nc = len(cells)
    types = np.array([t for t, _ in cells])
    cell_geom = np.zeros((nc, 4))
    K, F, C, L, order_key = [], [], [], [], []
    big = np.iinfo(np.int64).max
    for t in np.unique(types):
        idx = np.flatnonzero(types == t)
        conn = np.array([cells[c][1] for c in idx], np.int64)
        c0 = _seq_mean(nodes[np.sort(conn, axis=1)])               # mean over the sorted unique node numbers
        vol, cen = np.zeros(len(idx)), np.zeros((len(idx), 3))
        for lf, face in enumerate(_FACES3[int(t)]):
            fn = conn[:, list(face)]
            pts = nodes[fn]
            fcen = _seq_mean(pts)
            nfn = len(face)
            for k in range(nfn):
                a, b = pts[:, k], pts[:, (k + 1) % nfn]
                v = np.abs((np.cross(a - c0, b - c0) * (fcen - c0)).sum(1)) / 6.0
                vol = vol + v
                cen = cen + v[:, None] * (c0 + a + b + fcen) / 4.0
            key = np.full((len(idx), 4), big, np.int64)
            key[:, :nfn] = np.sort(fn, axis=1)
            ordered = np.full((len(idx), 4), -1, np.int64)
            ordered[:, :nfn] = fn
            K.append(key)
            F.append(ordered)
            C.append(idx)
            L.append(np.full(len(idx), nfn))
            order_key.append(idx * 8 + lf)                          # the order cells and local faces are walked in
        cell_geom[idx, :3] = cen / vol[:, None]
        cell_geom[idx, 3] = vol
    K, F, C, L, order_key = np.concatenate(K), np.concatenate(F), np.concatenate(C), np.concatenate(L), np.concatenate(order_key)
    walk = np.argsort(order_key, kind="stable")
    K, F, C, L = K[walk], F[walk], C[walk], L[walk]
    M = len(K)
    _, inv, counts = np.unique(K, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    assert counts.max() <= 2, "a face with more than two cells"
    first = np.full(len(counts), M, np.int64)
    np.minimum.at(first, inv, np.arange(M))
    last = np.zeros(len(counts), np.int64)
    np.maximum.at(last, inv, np.arange(M))
    faces = np.argsort(first, kind="stable")                        # faces in the order they are first met
    first, last, counts = first[faces], last[faces], counts[faces]
    c1 = C[first]
    cen, area, nrm = np.zeros((len(first), 3)), np.zeros(len(first)), np.zeros((len(first), 3))
    for nfn in (3, 4):
        sel = np.flatnonzero(L[first] == nfn)
        if len(sel) == 0:
            continue
        pts = nodes[F[first[sel], :nfn]]
        p0 = _seq_mean(pts)
        an, cc, atot = np.zeros((len(sel), 3)), np.zeros((len(sel), 3)), np.zeros(len(sel))
        for k in range(nfn):
            a, b = pts[:, k], pts[:, (k + 1) % nfn]
            n = 0.5 * np.cross(a - p0, b - p0)
            an = an + n
            ar = np.sqrt((n * n).sum(1))
            cc = cc + ar[:, None] * (p0 + a + b) / 3.0
            atot = atot + ar
        ar = np.sqrt((an * an).sum(1))
        cen[sel], area[sel], nrm[sel] = cc / atot[:, None], ar, an / ar[:, None]
    flip = ((cen - cell_geom[c1, :3]) * nrm).sum(1) < 0.0
    nrm[flip] = -nrm[flip]                                          # outward from the first owner
    inner = np.flatnonzero(counts == 2)
    outer = np.flatnonzero(counts == 1)
    a1, a2 = c1[inner], C[last[inner]]
    assert (a1 < a2).all()
    srt = np.lexsort((a2, a1))
    inner, a1, a2 = inner[srt], a1[srt], a2[srt]
    fcn, fnr, far = cen[inner], nrm[inner], area[inner]
    # face%calculate_distances (src/face.F90:230-250)
    d1 = ((fcn - cell_geom[a1, :3]) * fnr).sum(1)
    d2 = ((cell_geom[a2, :3] - fcn) * fnr).sum(1)
    d12 = ((cell_geom[a2, :3] - cell_geom[a1, :3]) * fnr).sum(1)
    corr = d12 / (d1 + d2)
    fg = np.zeros((len(inner), 12))
    fg[:, 0], fg[:, 1], fg[:, 2], fg[:, 3] = far, d1 * corr, d2 * corr, d12
    fg[:, 4:7], fg[:, 7], fg[:, 8:11] = fnr, fnr @ g, fcn
    rot = fnr.copy()
    if permeability_angle != 0.0:
        c, s_ = np.cos(permeability_angle), np.sin(permeability_angle)
        rot[:, :2] = fnr[:, :2] @ np.array([[c, s_], [-s_, c]]).T
    fg[:, 11] = np.argmax(np.abs(rot), axis=1) + 1
    fc = np.stack([a1, a2], 1).astype(np.int32)
    dist = ((cen[outer] - cell_geom[c1[outer], :3]) * nrm[outer]).sum(1)
    exterior = [(int(c1[k]), cen[k].copy(), float(area[k]), nrm[k].copy(), float(d)) for k, d in zip(outer, dist)]
    return cell_geom, fc, fg, exterior


def build_mesh(nodes, elems, thickness=1.0, radial=False, gravity=None, permeability_angle=0.0, vectorized=True):
    """Mesh (waiwera_b200.mesh.Mesh) of the top-dimensional elements.  Returns (mesh, exterior) where exterior is a
    list of (cell, face centroid, area, outward unit normal, distance) of the boundary faces, for boundary ghosts.
    3-D meshes go through the array version _build_mesh3 (vectorized=False: the cell-by-cell loop it was checked against)."""
    dim = max(_GMSH[t][0] for t, _ in elems)
    cells = [(t, n) for t, n in elems if _GMSH[t][0] == dim]
    nc = len(cells)
    g = np.zeros(3)
    if gravity is None:
        if dim == 3:
            g[2] = -9.8                                    # default_gravity_3D; 2-D default: none
    elif np.ndim(gravity) == 0:
        g[dim - 1] = -float(gravity)
    else:
        g[:len(gravity)] = gravity
    if dim == 3 and vectorized:
        cell_geom, fc, fg, exterior = _build_mesh3(nodes, cells, g, permeability_angle)
        m = wmesh.Mesh(ncell=nc, ninterior=nc, nowned=nc, face_cells=np.ascontiguousarray(fc), face_geom=np.ascontiguousarray(fg),
                       cell_geom=np.ascontiguousarray(cell_geom), rock=wmesh.default_rock(nc, None, heterogeneous=False),
                       dims=(nc, 1, 1), natural=np.arange(nc, dtype=np.int64), ncell_global=nc)
        m.gravity, m.dim, m.permeability_angle = g, dim, permeability_angle
        return m, exterior
    cell_geom = np.zeros((nc, 4))
    facemap = {}                                           # sorted node tuple -> [(cell, ordered nodes)]
    for c, (t, n) in enumerate(cells):
        if dim == 2:
            cen, area = _polygon_geometry(nodes[n, :2])
            cell_geom[c, :2] = cen
            # modify_cell_geometry (src/mesh.F90:369-395): thickness, or Pappus for a solid of revolution
            cell_geom[c, 3] = area * (2.0 * np.pi * cen[0] if radial else thickness)
            local = [(n[k], n[(k + 1) % len(n)]) for k in range(len(n))]
        else:
            local = [tuple(n[i] for i in f) for f in _FACES3[t]]
            cell_geom[c, :3], cell_geom[c, 3] = _cell3_geometry(nodes, local)
        for f in local:
            facemap.setdefault(tuple(sorted(f)), []).append((c, f))
    interior, exterior = [], []
    for key, owners in facemap.items():
        c1, f = owners[0]
        if dim == 2:
            a, b = nodes[f[0], :2], nodes[f[1], :2]
            cen = np.zeros(3)
            cen[:2] = 0.5 * (a + b)
            length = np.linalg.norm(b - a)
            nrm = np.zeros(3)
            nrm[:2] = np.array([b[1] - a[1], -(b[0] - a[0])]) / length
            # modify_face_geometry (src/mesh.F90:399-432)
            area = length * (2.0 * np.pi * cen[0] if radial else thickness)
        else:
            cen, area, nrm = _face3_geometry(nodes[list(f)])
        if np.dot(cen - cell_geom[c1, :3], nrm) < 0.0:
            nrm = -nrm                                     # outward from the first owner
        if len(owners) == 2:
            c2 = owners[1][0]
            if c2 < c1:
                c1, c2, nrm = c2, c1, -nrm
            interior.append((c1, c2, cen, area, nrm))
        else:
            exterior.append((c1, cen, area, nrm, float(np.dot(cen - cell_geom[c1, :3], nrm))))
    interior.sort(key=lambda r: (r[0], r[1]))
    nf = len(interior)
    fc, fg = np.zeros((nf, 2), np.int32), np.zeros((nf, 12))
    for k, (c1, c2, cen, area, nrm) in enumerate(interior):
        fc[k] = (c1, c2)
        # face%calculate_distances (src/face.F90:230-250)
        d1 = np.dot(cen - cell_geom[c1, :3], nrm)
        d2 = np.dot(cell_geom[c2, :3] - cen, nrm)
        d12 = np.dot(cell_geom[c2, :3] - cell_geom[c1, :3], nrm)
        corr = d12 / (d1 + d2)
        fg[k, 0], fg[k, 1], fg[k, 2], fg[k, 3] = area, d1 * corr, d2 * corr, d12
        fg[k, 4:7], fg[k, 7], fg[k, 8:11] = nrm, np.dot(g, nrm), cen
        fg[k, 11] = permeability_direction(nrm, dim, permeability_angle)
    m = wmesh.Mesh(ncell=nc, ninterior=nc, nowned=nc, face_cells=np.ascontiguousarray(fc), face_geom=np.ascontiguousarray(fg),
                   cell_geom=np.ascontiguousarray(cell_geom), rock=wmesh.default_rock(nc, None, heterogeneous=False),
                   dims=(nc, 1, 1), natural=np.arange(nc, dtype=np.int64), ncell_global=nc)
    m.gravity = g
    m.dim = dim
    m.permeability_angle = permeability_angle
    return m, exterior


def permeability_direction(normal, dim, angle=0.0):
    """face%calculate_permeability_direction (src/face.F90:210-226): axis of the (rotated) permeability tensor
    closest to the face normal, 1-based"""
    n = np.array(normal[:3], float)
    if angle != 0.0:
        c, s = np.cos(angle), np.sin(angle)
        n[:2] = np.array([[c, s], [-s, c]]) @ n[:2]
    return int(np.argmax(np.abs(n))) + 1


def add_boundary_faces(m, exterior, specs):
    """Dirichlet ghost cells for the "boundaries" of the input: every spec {"faces": {"cells": [...], "normal":
    [...]}} (or a list of those) selects, per listed cell, the exterior face whose outward normal is closest to the
    given one (dm_cell_normal_face, src/mesh.F90:1772-1800).  Returns (mesh, boundary index of every ghost cell)."""
    by_cell = {}
    for e in exterior:
        by_cell.setdefault(e[0], []).append(e)
    cells, owner, rows = [], [], []
    for ib, spec in enumerate(specs):
        faces = spec.get("faces", {})
        for fs in (faces if isinstance(faces, list) else [faces]):
            nrm = np.zeros(3)
            given = fs.get("normal", [0.0, 0.0, 1.0])
            nrm[:len(given)] = given
            for c in fs.get("cells", []):
                cand = by_cell.get(c, [])
                assert cand, "cell %d has no exterior face" % c
                e = max(cand, key=lambda r: float(np.dot(r[3], nrm)))
                cells.append(c)
                owner.append(ib)
                rows.append(e)
    n0 = m.ncell
    nb = len(cells)
    fg = np.zeros((nb, 12))
    gg = np.zeros((nb, 4))
    for k, (c, cen, area, nrm, dist) in enumerate(rows):
        # mesh_boundary_face_geometry (src/mesh.F90:583-664): distances (d1, 0), ghost volume 0 at the face centroid
        fg[k, 0], fg[k, 1], fg[k, 2], fg[k, 3] = area, dist, 0.0, dist
        fg[k, 4:7], fg[k, 7], fg[k, 8:11] = nrm, np.dot(m.gravity, nrm), cen
        fg[k, 11] = permeability_direction(nrm, m.dim, m.permeability_angle)
        gg[k, :3] = cen
    ghosts = n0 + np.arange(nb)
    out = wmesh.Mesh(ncell=n0 + nb, ninterior=m.ninterior, nowned=m.nowned,
                     face_cells=np.ascontiguousarray(np.vstack([m.face_cells, np.stack([cells, ghosts], 1)]).astype(np.int32)) if nb else m.face_cells,
                     face_geom=np.ascontiguousarray(np.vstack([m.face_geom, fg])),
                     cell_geom=np.ascontiguousarray(np.vstack([m.cell_geom, gg])),
                     rock=np.ascontiguousarray(np.vstack([m.rock, m.rock[cells]])) if nb else m.rock,
                     dims=m.dims, natural=m.natural, ncell_global=m.ncell_global,
                     boundary={"ghost_cells": ghosts.astype(np.int32), "interior_cells": np.array(cells, np.int32)} if nb else {})
    out.gravity, out.dim, out.permeability_angle = m.gravity, m.dim, m.permeability_angle
    out.minc_levels, out.minc_base = m.minc_levels, m.minc_base
    if hasattr(m, "minc_zone"):
        out.minc_zone, out.minc_cells, out.minc_parent, out.minc_level = m.minc_zone, m.minc_cells, m.minc_parent, m.minc_level
    return out, np.array(owner, np.int32)


def _zone_type(spec):
    """get_zone_type (src/zone.F90:91-163): an array is a cell list; an object says its "type" (array / box / combine) or
    shows it by its keys: "cells"; "x" "y" "z" "r"; "+" "-" "*" """
    if isinstance(spec, list):
        return "array"
    if isinstance(spec, dict):
        t = str(spec.get("type", "")).lower()
        if t in ("array", "box", "combine"):
            return t
        if "cells" in spec:
            return "array"
        if any(k in spec for k in ("x", "y", "z", "r")):
            return "box"
        if any(k in spec for k in ("+", "-", "*")):
            return "combine"
    return None


def _zone_cells(zone, m, zones=None, _seen=()):
    """interior cells of a zone (src/zone.F90): cell list; box in the cell-centroid coordinates, limits inclusive, a
    missing axis unbounded ("r" is the first coordinate of a radial mesh); combination: the union of the "+" zones (all
    cells if there are none), intersected with every "*" zone, without the "-" zones.  `zone`: a name in `zones` or a
    zone value."""
    zones = zones or {}
    if isinstance(zone, str):
        assert zone in zones, "unknown zone %r" % zone
        assert zone not in _seen, "zone %r depends on itself" % zone
        return _zone_cells(zones[zone], m, zones, _seen + (zone,))
    n = m.ninterior
    kind = _zone_type(zone)
    if kind == "array":
        cells = zone if isinstance(zone, list) else zone.get("cells", [])
        return np.array(sorted(set(int(c) for c in cells)), np.int64)
    if kind == "box":
        sel = np.ones(n, bool)
        for ax, k in (("x", 0), ("r", 0), ("y", 1), ("z", 2)):
            if ax in zone:
                lo, hi = zone[ax]
                sel &= (m.cell_geom[:n, k] >= lo) & (m.cell_geom[:n, k] <= hi)
        return np.nonzero(sel)[0]
    if kind == "combine":
        def names(key):
            v = zone.get(key)
            return [] if v is None else ([v] if isinstance(v, str) else list(v))
        sel = np.zeros(n, bool)
        if names("+"):
            for z in names("+"):
                sel[_zone_cells(z, m, zones, _seen)] = True
        else:
            sel[:] = True
        for z in names("*"):
            keep = np.zeros(n, bool)
            keep[_zone_cells(z, m, zones, _seen)] = True
            sel &= keep
        for z in names("-"):
            sel[_zone_cells(z, m, zones, _seen)] = False
        return np.nonzero(sel)[0]
    raise ValueError("unrecognised zone %r" % (zone,))


def _rock_type_cells(rt, m, zones):
    idx = list(rt.get("cells", []))
    zs = rt.get("zones", [])
    for z in ([zs] if isinstance(zs, str) else zs):
        idx += _zone_cells(z, m, zones).tolist()
    return np.array(sorted(set(idx)), np.int64)


def _rank(v):
    return 0 if not isinstance(v, (list, tuple)) else 1 + max([_rank(x) for x in v], default=0)


def rock_controls(spec, m, zones=None):
    """Rock properties given as tables in time (setup_permeability_rock_control / setup_porosity_rock_control,
    src/rock_setup.F90:383-463): a rock type whose "permeability" is an array of rank 2 (rows [t, k] or
    [t, k1, k2(, k3)]) or whose "porosity" is an array (rows [t, phi]), with the "interpolation" of the type (linear or
    step).  Returns a list of (column of the rock record, number of columns, cells, table, interpolation) in the order
    of the reference's control list (per type: permeability, then porosity); rock_at applies them."""
    out = []
    for rt in (spec or {}).get("types", []):
        idx = _rock_type_cells(rt, m, zones)
        if len(idx) == 0:
            continue
        interp = rt.get("interpolation", "linear")
        assert interp in ("linear", "step"), "%s interpolation of a rock property table is not built" % interp
        if _rank(rt.get("permeability")) == 2:
            tab = np.array(rt["permeability"], float)
            assert 2 <= tab.shape[1] <= 4, "rock permeability table: rows are [time, k] or [time, k1, k2(, k3)]"
            out.append((0, tab.shape[1] - 1, idx, tab, interp))
        if _rank(rt.get("porosity")) >= 1:
            tab = np.array(rt["porosity"], float).reshape(-1, 2)
            out.append((5, 1, idx, tab, interp))
    return out


def apply_rock_controls(rock, controls, t):
    """the update of every rock control at time t on the 8-double records `rock` (in place): table%interpolate(t),
    a one-column permeability table sets all three directions, a wider one the leading directions
    (permeability_table_rock_control_update / porosity_table_rock_control_update, src/rock_control.F90:49-116)"""
    for col, ncol, idx, tab, interp in controls:
        v = [_table_value(tab[:, [0, 1 + j]], interp, t) for j in range(ncol)]
        if col == 0 and ncol == 1:
            rock[idx, 0:3] = v[0]
        else:
            rock[idx, col:col + ncol] = v
    return rock


def rock_at(p, t):
    """rock records of the interior cells at time t, for wb_set_rock before a time step is tried
    (flow_simulation_update_rock_properties called from pre_try_timestep with the time the step ends at,
    src/flow_simulation.F90:2040-2089, src/timestepper.F90:2333); None if the input has no rock controls"""
    if not getattr(p, "rock_controls", None):
        return None
    n = p.mesh.ninterior
    return apply_rock_controls(np.array(p.mesh.rock[:n], float), p.rock_controls, t)


def rock_records(spec, m, zones=None, time=0.0):
    """8-double rock records from the "rock" value of the input (src/rock_setup.F90:236-465; defaults src/rock.F90:69-76);
    properties given as tables in time (rock_controls) take their value at `time`, the start of the run
    (src/flow_simulation.F90:971)"""
    n = m.ninterior
    rock = np.zeros((n, 8))
    rock[:, 0:3], rock[:, 3:5], rock[:, 5], rock[:, 6], rock[:, 7] = 1e-13, 2.5, 0.1, 2200.0, 1000.0
    for rt in (spec or {}).get("types", []):
        idx = _rock_type_cells(rt, m, zones)
        if len(idx) == 0:
            continue
        if rt.get("permeability") is not None and _rank(rt["permeability"]) < 2:
            k = np.atleast_1d(np.array(rt["permeability"], float))
            rock[idx, 0:len(k)] = k
            if len(k) == 1:
                rock[idx, 0:3] = k[0]
        for key, col in (("wet_conductivity", 3), ("dry_conductivity", 4), ("porosity", 5), ("density", 6), ("specific_heat", 7)):
            if rt.get(key) is not None and _rank(rt[key]) == 0:
                rock[idx, col] = rt[key]
        if rt.get("dry_conductivity") is None and rt.get("wet_conductivity") is not None:
            rock[idx, 4] = rt["wet_conductivity"]          # dry defaults to wet (src/rock_setup.F90)
    return apply_rock_controls(rock, rock_controls(spec, m, zones), time)


def apply_minc(m, minc, rock_spec, zones=None):
    """"mesh.minc" of the input (src/minc.F90:73-190 geometry and zones, src/mesh.F90:3186-3375 rock properties) on a
    mesh without boundary ghosts -> the MINC mesh of mesh.add_minc_zones.  "minc" is one entry or a list of entries with
    their own geometries; the cells of an entry are the union of the "zones" and of the cells of the rock "types" of
    every item of its "rock".  There the fracture cells take the properties the "fracture" rock type gives (the others
    stay), the matrix cells those of the "matrix" rock type, porosity by default the one that keeps the void fraction of
    the original cell."""
    types = {rt.get("name"): rt for rt in (rock_spec or {}).get("types", [])}

    def cells_of_type(name):
        assert name in types, "unrecognised rock type %r" % name
        rt = types[name]
        idx = list(rt.get("cells", []))
        zs = rt.get("zones", [])
        for z in ([zs] if isinstance(zs, str) else zs):
            idx += _zone_cells(z, m, zones).tolist()
        return idx

    def properties(entry, which):
        """the 8 rock properties of the named rock type, -1 where it does not give one"""
        out = np.full(8, -1.0)
        if which not in entry:
            return out
        assert "type" in entry[which], "mesh.minc.rock: %s.type not found" % which
        name = entry[which]["type"]
        assert name in types, "unrecognised rock type %r" % name
        rt = types[name]
        if rt.get("permeability") is not None:
            k = np.atleast_1d(np.asarray(rt["permeability"], float))
            if len(k) == 1:
                out[0:3] = k[0]
            else:
                out[0:len(k)] = k
        for key, col in (("wet_conductivity", 3), ("dry_conductivity", 4), ("porosity", 5), ("density", 6), ("specific_heat", 7)):
            if rt.get(key) is not None:
                out[col] = rt[key]
        return out

    n = m.ninterior
    orig = m.rock[:n].copy()
    specs = []
    for spec in ([minc] if isinstance(minc, dict) else list(minc)):
        geom = spec.get("geometry", {})
        fr, mx = geom.get("fracture", {}), geom.get("matrix", {})
        mvol = mx.get("volume")
        if "volume" in fr:
            fvol = float(fr["volume"])
            mvol = [1.0 - fvol] if mvol is None else list(np.atleast_1d(mvol).astype(float))
        else:
            mvol = [0.9] if mvol is None else list(np.atleast_1d(mvol).astype(float))
            fvol = 1.0 - sum(mvol)
        volumes = np.array([fvol] + mvol)
        volumes = volumes / volumes.sum()
        planes = int(fr.get("planes", 1))
        sp = np.atleast_1d(np.asarray(fr.get("spacing", 50.0), float))
        spacing = np.full(planes, sp[0])
        spacing[:min(len(sp), planes)] = sp[:planes]
        entry_of = np.full(n, -1)
        rocks = spec.get("rock", [])
        rocks = [rocks] if isinstance(rocks, dict) else list(rocks)
        for k, entry in enumerate(rocks):
            zs = entry.get("zones", [])
            for z in ([zs] if isinstance(zs, str) else zs):
                entry_of[_zone_cells(z, m, zones)] = k
            ts = entry.get("types", [])
            for t in ([ts] if isinstance(ts, str) else ts):
                entry_of[np.array(cells_of_type(t), np.int64)] = k
        zone = np.nonzero(entry_of >= 0)[0]
        matrix = orig[zone].copy()
        for k, entry in enumerate(rocks):
            rows = np.nonzero(entry_of[zone] == k)[0]
            sel = zone[rows]
            if len(sel) == 0:
                continue
            fp, mp = properties(entry, "fracture"), properties(entry, "matrix")
            fpor = np.where(fp[5] < 0, orig[sel, 5], fp[5])
            mpor = (orig[sel, 5] - fpor * volumes[0]) / (1.0 - volumes[0]) if mp[5] < 0 else np.full(len(sel), mp[5])
            matrix[rows] = np.where(mp > 0, mp, orig[sel])
            matrix[rows, 5] = mpor
            m.rock[sel] = np.where(fp > 0, fp, orig[sel])
            m.rock[sel, 5] = fpor
        specs.append(dict(cells=zone, volumes=volumes, spacing=spacing, matrix_rock=matrix,
                          fracture_connection_distance=float(fr.get("connection", 0.0))))
    out = wmesh.add_minc_zones(m, specs)
    out.gravity, out.dim, out.permeability_angle = m.gravity, m.dim, m.permeability_angle
    out.minc_cells = n
    out.minc_zone = np.unique(np.concatenate([z["cells"] for z in specs])) if specs else np.zeros(0, np.int64)
    return out


_EOS = {"we": ("EOS_WE", 2), "w": ("EOS_W", 1), "wce": ("EOS_WCE", 3), "wae": ("EOS_WAE", 3)}


def make_params(mod, doc, gravity):
    """wb_params from the eos / thermodynamics / rock curves.  mod: the module whose make_params / make_relperm /
    make_cappress build the struct (waiwera_b200.flow; the parity tests pass their checker's module, same layout)"""
    eos = doc.get("eos", "we")
    name = eos if isinstance(eos, str) else eos.get("name", "we")
    assert name in _EOS, "eos %r is not built" % name
    thermo = doc.get("thermodynamics", "iapws")
    thermo = thermo if isinstance(thermo, str) else thermo.get("name", "iapws")
    rock = doc.get("rock") or {}
    rp = rock.get("relative_permeability") or {"type": "linear"}
    kw = {k: v for k, v in rp.items() if k != "type"}
    if rp.get("type", "linear") == "linear":
        kw = {"liquid": tuple(rp.get("liquid", (0.0, 1.0))), "vapour": tuple(rp.get("vapour", (0.0, 1.0)))}
    relperm = mod.make_relperm(rp.get("type", "linear").lower().replace(" ", "_"), **kw)
    cp = rock.get("capillary_pressure") or {"type": "zero"}
    ckw = {k: v for k, v in cp.items() if k != "type"}
    if "saturation_limits" in ckw:
        ckw["saturation_limits"] = tuple(ckw["saturation_limits"])
    cappress = mod.make_cappress(cp.get("type", "zero").lower().replace(" ", "_"), **ckw)
    kwargs = dict(eos=getattr(mod, _EOS[name][0]), thermo=mod.THERMO_IFC67 if thermo.lower() == "ifc67" else mod.THERMO_IAPWS,
                  relperm=relperm, cappress=cappress, gravity=tuple(gravity))
    if name == "w" and not isinstance(eos, str) and "temperature" in eos:
        kwargs["eos_w_temperature"] = eos["temperature"]
    prm = mod.make_params(**kwargs)
    # eos.primary.scale (src/eos_we.F90:75-109, src/eos_wge.F90:96-110): pressure, temperature, partial_pressure
    # (a number, or "pressure" for the adaptive Pg / P scaling, the default)
    scale = ({} if isinstance(eos, str) else eos.get("primary", {}).get("scale", {})) or {}
    if "pressure" in scale:
        prm.pressure_scale = float(scale["pressure"])
    if "temperature" in scale:
        prm.temperature_scale = float(scale["temperature"])
    pp = scale.get("partial_pressure", scale.get("air_partial_pressure", scale.get("CO2_partial_pressure")))
    if pp is not None:
        prm.partial_pressure_scale = 0.0 if isinstance(pp, str) else float(pp)
    return prm, _EOS[name][1]


def _component(v, np_):
    """component number of a name or number (get_component, src/source_setup.F90:2087-2121): mass components by the
    EOS's names, "energy" = the last primary variable"""
    names = {"water": 1, "energy": np_, "co2": 2, "air": 2, "ncg": 2, "gas": 2}
    return names[v.lower()] if isinstance(v, str) else int(v)


def _tracer_rates(v, tracers):
    """tracer injection rates of a source: one number for all tracers, a list, or {tracer name: rate} (the others 0)"""
    nt = max(len(tracers), 1)
    if isinstance(v, list) and v and isinstance(v[0], list):
        return [0.0] * nt                                  # a table in time: tracer_rates_at
    if isinstance(v, dict):
        names = [t.get("name") for t in tracers]
        return [float(v.get(nm, 0.0)) for nm in names] + [0.0] * (nt - len(names))
    return np.broadcast_to(np.atleast_1d(np.asarray(v, float)), (nt,)).tolist()


def expand_sources(doc, m, np_, zones=None):
    """The source list in natural source order (setup_source_network, src/source_setup.F90:240-340, 680-760): a
    specification with "cell", with "cells" (a number or a list) or with "zones" gives one source per cell (zone cells in
    natural order), all with the same parameters; one without any of them a source without a cell, which does nothing.
    Returns (the specifications per source with "cell" set, sources without a cell left out; the table of all sources:
    cell (-1: none), rate, injection enthalpy, injection / production component as the reference stores them)."""
    eos = doc.get("eos", "we")
    eos = eos if isinstance(eos, str) else eos.get("name", "we")
    tr = doc.get("tracer")
    tracers = [] if tr is None else ([tr] if isinstance(tr, dict) else list(tr))
    out, table = [], []
    for s in doc.get("source") or []:
        if "cell" in s and s["cell"] is not None:
            cells = [int(s["cell"])]
        elif "cells" in s:
            cells = [int(c) for c in np.atleast_1d(s["cells"])]
        elif "zones" in s:
            zs = s["zones"]
            sel = set()
            for z in ([zs] if isinstance(zs, str) else zs):
                sel |= set(_zone_cells(z, m, zones).tolist())
            cells = sorted(sel)
        else:
            cells = [-1]
        inj = _component(s["component"], np_) if "component" in s else 0
        if "production_component" in s:
            prod = _component(s["production_component"], np_)
        else:
            prod = inj if (inj == np_ and np_ > 1 and eos != "w") else 0
        rate = s.get("rate", 0.0)
        h = 0.0 if (inj == np_ and np_ > 1 and eos != "w") else (s["enthalpy"] if isinstance(s.get("enthalpy"), (int, float)) else 83.9e3)
        for c in cells:
            table.append(dict(cell=c, rate=rate if isinstance(rate, (int, float)) else None, enthalpy=float(h),
                              injection_component=inj, production_component=prod, tracer=_tracer_rates(s.get("tracer", 0.0), tracers),
                              name=s.get("name")))
            if c >= 0:
                one = {k: v for k, v in s.items() if k not in ("cells", "zones")}
                one["cell"] = c
                out.append(one)
    return out, table


def load_json(path):
    """a JSON input as Waiwera's parser (fson) accepts it: standard JSON, and also numbers written "20." or ".5" and a
    comma before a closing bracket -- decks written for the reference contain these (its own test inputs do)"""
    import re
    text = open(path).read()
    try:
        return json.loads(text)
    except json.JSONDecodeError:
        pass
    parts = re.split(r'("(?:\\.|[^"\\])*")', text)
    for i in range(0, len(parts), 2):                         # the segments outside string literals
        seg = parts[i]
        seg = re.sub(r"(?<![\w.])(\d+)\.(?![\d])", r"\1.0", seg)
        seg = re.sub(r"(?<![\w.])\.(\d)", r"0.\1", seg)
        seg = re.sub(r",(\s*[}\]])", r"\1", seg)
        parts[i] = seg
    return json.loads("".join(parts))


class Problem:
    """what load() returns: mesh, initial state, boundary values, sources, tracers, time stepping"""


def load(path, mod=None, mesh_path=None):
    """Reads <path> (Waiwera JSON input) and the gmsh mesh it names.  mod: waiwera_b200.flow (see make_params),
    None: no parameter struct."""
    doc = load_json(path)
    if doc.get("network"):
        raise NotImplementedError("%s: source networks (\"network\": groups, reinjectors) are not built" % path)
    eos_name = doc.get("eos", "we")
    eos_name = eos_name if isinstance(eos_name, str) else eos_name.get("name", "we")
    if eos_name not in _EOS:
        raise NotImplementedError("%s: eos %r is not built (w, we, wce, wae are)" % (path, eos_name))
    mspec = doc["mesh"] if isinstance(doc["mesh"], dict) else {"filename": doc["mesh"]}
    mfile = mesh_path or os.path.join(os.path.dirname(path), mspec["filename"])
    nodes, elems = read_mesh(mfile)
    m, exterior = build_mesh(nodes, elems, thickness=mspec.get("thickness", 1.0), radial=bool(mspec.get("radial", False)),
                             gravity=doc.get("gravity"), permeability_angle=np.deg2rad(mspec.get("permeability_angle", 0.0)))
    # "mesh.faces": permeability directions set by hand for the face between two cells (src/mesh.F90:1268-1350, 1355-1410)
    for fs in mspec.get("faces") or []:
        cells = sorted(int(c) for c in fs.get("cells", []))
        if len(cells) == 2:
            hit = np.nonzero((np.sort(m.face_cells.reshape(-1, 2), 1) == cells).all(1))[0]
            if len(hit) == 1:
                m.face_geom[hit[0], 11] = float(fs.get("permeability_direction", 1))
    start_time = float((doc.get("time") or {}).get("start", 0.0) or 0.0)
    rock = rock_records(doc.get("rock"), m, mspec.get("zones"), time=start_time)
    rctl = rock_controls(doc.get("rock"), m, mspec.get("zones"))
    m.rock[:] = rock
    assert not (rctl and mspec.get("minc")), "rock property tables on MINC meshes are not built"
    if mspec.get("minc"):
        m = apply_minc(m, mspec["minc"], doc.get("rock"), mspec.get("zones"))
    bspecs = doc.get("boundaries") or []
    m, bowner = add_boundary_faces(m, exterior, bspecs)
    p = Problem()
    p.doc, p.mesh = doc, m
    p.rock_controls = rctl                                  # cells are interior cells: boundary ghosts come after them
    eos_name = doc.get("eos", "we")
    p.eos = eos_name if isinstance(eos_name, str) else eos_name.get("name", "we")
    if mod is not None:
        p.params, p.np = make_params(mod, doc, m.gravity)
    else:
        eos = doc.get("eos", "we")
        p.params, p.np = None, _EOS[eos if isinstance(eos, str) else eos.get("name", "we")][1]
    n = m.ninterior
    init = doc.get("initial") or {}
    if "primary" in init:
        prim = np.array(init["primary"], float)
        reg = np.array(init.get("region", 1))
        n0 = getattr(m, "minc_cells", n)
        if n0 < n and not init.get("minc", False) and prim.ndim == 2:
            # values for the original cells only: a matrix cell starts from its fracture cell (src/initial.F90:976-1060)
            parents = m.minc_parent
            prim = prim.reshape(n0, -1)[parents]
            reg = reg if reg.ndim == 0 else reg[parents]
        p.primary = np.tile(prim, (n, 1)) if prim.ndim == 1 else prim.reshape(n, -1)
        p.region = (np.full(n, int(reg)) if reg.ndim == 0 else reg).astype(np.int32)
        # scaled with the SAME scales the parameters carry (eos.primary.scale of the input, else the defaults)
        sc = dict(pressure_scale=1e6, temperature_scale=1e2, partial_pressure_scale=0.0)
        if p.params is not None:
            sc = dict(pressure_scale=p.params.pressure_scale if p.params.pressure_scale > 0 else 1e6,
                      temperature_scale=p.params.temperature_scale if p.params.temperature_scale > 0 else 1e2,
                      partial_pressure_scale=max(p.params.partial_pressure_scale, 0.0))
        p.primary_scales = sc
        p.y = np.ascontiguousarray(wmesh.scale_primaries(p.primary, p.region, **sc)).reshape(-1)
    elif "filename" in init and os.path.exists(os.path.join(os.path.dirname(path), init["filename"])):
        # restart from a Waiwera HDF5 output file (src/initial.F90:421-507): the fields of the time index "index"
        # (default -1: the last one, src/initial.F90:776; negative: from the end), read without an HDF5 library
        # (waiwera_b200/h5lite.py)
        from . import output
        eos_ = doc.get("eos", "we")
        eos_ = eos_ if isinstance(eos_, str) else eos_.get("name", "we")
        p.primary, p.region, p.restart_time = output.read_restart(os.path.join(os.path.dirname(path), init["filename"]), eos_,
                                                                  int(init.get("index", -1)))
        n0 = getattr(m, "minc_cells", n)
        if n0 < n and len(p.region) == n0 and not init.get("minc", False):
            # a file of the original cells for a MINC mesh: a matrix cell starts from its fracture cell
            p.primary, p.region = p.primary[m.minc_parent], p.region[m.minc_parent]
        assert len(p.region) == n, "restart file holds %d cells, the mesh %d" % (len(p.region), n)
        sc = dict(pressure_scale=1e6, temperature_scale=1e2, partial_pressure_scale=0.0)
        if p.params is not None:
            sc = dict(pressure_scale=p.params.pressure_scale if p.params.pressure_scale > 0 else 1e6,
                      temperature_scale=p.params.temperature_scale if p.params.temperature_scale > 0 else 1e2,
                      partial_pressure_scale=max(p.params.partial_pressure_scale, 0.0))
        p.primary_scales = sc
        p.y = np.ascontiguousarray(wmesh.scale_primaries(p.primary, p.region, **sc)).reshape(-1)
    elif "filename" in init:
        p.primary = p.region = p.y = None                  # restart file not at hand: pass the arrays
    else:
        # no initial conditions: the default primaries of the EOS in region 1 everywhere (src/initial.F90:941-951)
        p.primary = np.tile(np.array([1.0e5, 20.0, 0.0][:p.np]), (n, 1))
        p.region = np.ones(n, np.int32)
        sc = dict(pressure_scale=1e6, temperature_scale=1e2, partial_pressure_scale=0.0)
        if p.params is not None:
            sc = dict(pressure_scale=p.params.pressure_scale if p.params.pressure_scale > 0 else 1e6,
                      temperature_scale=p.params.temperature_scale if p.params.temperature_scale > 0 else 1e2,
                      partial_pressure_scale=max(p.params.partial_pressure_scale, 0.0))
        p.primary_scales = sc
        p.y = np.ascontiguousarray(wmesh.scale_primaries(p.primary, p.region, **sc)).reshape(-1)
    tr = doc.get("tracer")
    p.tracers = [] if tr is None else ([tr] if isinstance(tr, dict) else list(tr))
    nt = len(p.tracers)
    # eos%default_primary / default_region (src/eos_w.F90:80, eos_we.F90:90, eos_wge.F90:80) where the input gives none
    default_primary = [1.0e5, 20.0, 0.0][:p.np]
    p.boundary_primary = np.array([bspecs[i].get("primary", default_primary) for i in bowner], float).reshape(len(bowner), p.np)
    p.boundary_region = np.array([bspecs[i].get("region", 1) for i in bowner], np.int32)
    p.boundary_tracer = np.array([np.broadcast_to(np.atleast_1d(bspecs[i].get("tracer", 0.0)), (max(nt, 1),)) for i in bowner],
                                 float).reshape(len(bowner), max(nt, 1))
    p.initial_tracer = np.broadcast_to(np.atleast_1d(init.get("tracer", 0.0)), (max(nt, 1),)).astype(float)
    src, p.source_specs = expand_sources(doc, m, p.np, mspec.get("zones"))
    for s in src:
        unsupported = set(s) - {"cell", "rate", "component", "production_component", "enthalpy", "name", "tracer",
                                "interpolation", "averaging", "deliverability", "direction", "limiter", "separator",
                                "recharge", "injectivity", "factor"}
        assert not unsupported, "source controls are not built: %s" % sorted(unsupported)
        dl = s.get("deliverability")
        assert dl is None or "threshold" not in dl, "deliverability thresholds are not built"

    def as_table(v, s, sub=None):
        """a control parameter given as a table in time: [[t, v], ...] or {"time": [[t, v], ...]} (a sub-object may carry
        its own "interpolation" / "averaging", else the source's: src/source_setup.F90); None for a plain number"""
        own = sub if isinstance(sub, dict) else {}
        if isinstance(v, dict) and "time" in v:
            own = dict(own, **{k: v[k] for k in ("interpolation", "averaging") if k in v})
            v = v["time"]
        if isinstance(v, list) and v and isinstance(v[0], list):
            return (np.array(v, float), own.get("interpolation", s.get("interpolation", "linear")),
                    own.get("averaging", s.get("averaging", "integrate")))
        return None
    # a rank-2 "rate" is a table source control (src/source_control.F90: table of (time, rate)); kept as a table
    # that rates_at() evaluates over each time step, "step" or "linear" interpolation, "endpoint" averaging
    p.source_tables = {}
    kept = []
    for s in src:
        if isinstance(s.get("rate"), list):
            p.source_tables[len(kept)] = (np.array(s["rate"], float), s.get("interpolation", "linear"),
                                          s.get("averaging", "integrate"))
            s = dict(s, rate=0.0)
            kept.append(s)
        elif s.get("rate", 0.0) != 0.0 or "deliverability" in s or "recharge" in s or "injectivity" in s:
            kept.append(s)
    src = kept

    def separator_pressures(s):
        """get_separator_pressure (src/source_setup.F90:2255-2330): "separator": true | {"pressure": p | [p1, p2]}, or the
        "separator_pressure" of a water / steam limiter in the single-type syntax; [] = no separator"""
        sep = s.get("separator")
        if sep is not None:
            if sep is True or (isinstance(sep, dict) and "pressure" not in sep):
                return [0.55e6]                      # default_separator_pressure, src/separator.F90:34
            if sep is False:
                return []
            return [float(v) for v in np.atleast_1d(sep["pressure"])]
        lm = s.get("limiter")
        if lm is not None and str(lm.get("type", "total")).lower() != "total" and "separator_pressure" in lm:
            return [float(v) for v in np.atleast_1d(lm["separator_pressure"])]
        return []

    def limits(s):
        """add_limiter (src/source_setup.F90:3117-3276): {"type": t, "limit": x} (one flow type, default total, default
        limit 1) or {"total": x, "water": y, "steam": z}"""
        lm = s.get("limiter")
        out = {"total": 0.0, "water": 0.0, "steam": 0.0}
        if lm is None:
            return out
        if "limit" in lm or "type" in lm:
            out[str(lm.get("type", "total")).lower()] = float(lm.get("limit", 1.0))
        else:
            for k in out:
                if k in lm:
                    out[k] = float(lm[k])
        return out
    # source controls (see wb_set_source_controls): deliverability (productivity None: to be calculated from the
    # initial rate, src/source_control.F90:407-468), direction, total-flow limiter
    # parameters given as tables in time are kept in p.source_control_tables[(source, key)] and evaluated over each
    # time step by controls_at(); key: "productivity", "reference_pressure", "limit", "limit_water", "limit_steam",
    # "factor" (rate_factor_source_control: the rate after the other controls times the factor)
    p.source_controls = []
    p.source_control_tables = {}
    # reference pressures of sources on deliverability tabulated against the flowing enthalpy or the pressure
    # (wb_set_source_pressure_table; src/source_setup.F90:2704-2717): source, coordinate (0 enthalpy, 1 pressure), points,
    # step interpolation flag
    p.source_pressure_tables = []
    for k, s in enumerate(src):
        if "factor" in s:
            tab = as_table(s["factor"], s, s["factor"])
            if tab is not None:
                p.source_control_tables[(k, "factor")] = tab
            else:
                p.source_control_tables[(k, "factor")] = (np.array([[0.0, float(s["factor"])]]), "step", "integrate")
        if "deliverability" in s or "limiter" in s or "direction" in s or "recharge" in s or "injectivity" in s:
            dl = dict(s.get("deliverability") or {})
            pr = dl.get("pressure")
            if isinstance(pr, dict) and "time" not in pr and ("enthalpy" in pr or "pressure" in pr):
                coord = 0 if "enthalpy" in pr else 1
                pts = np.array(pr["enthalpy" if coord == 0 else "pressure"], float).reshape(-1, 2)
                interp = str(pr.get("interpolation", s.get("interpolation", "linear"))).lower()
                assert interp in ("linear", "step"), "%s interpolation of a reference pressure table is not built" % interp
                assert len(pts) <= 8, "reference pressure tables of more than 8 points are not built"
                p.source_pressure_tables.append(dict(source=k, coordinate=coord, table=pts.tolist(), step=int(interp == "step")))
                dl["pressure"] = float(pts[0, 1])
            for key, name in (("productivity", "productivity"), ("pressure", "reference_pressure")):
                tab = as_table(dl.get(key), s, dl)
                if tab is not None:
                    p.source_control_tables[(k, name)] = tab
                    dl[key] = float(tab[0][0, 1])
            lm = s.get("limiter") or {}
            for key, name in (("limit", "limit"), ("total", "limit"), ("water", "limit_water"), ("steam", "limit_steam")):
                tab = as_table(lm.get(key), s, lm)
                if tab is not None:
                    if key == "limit":
                        name = {"total": "limit", "water": "limit_water", "steam": "limit_steam"}[str(lm.get("type", "total")).lower()]
                    p.source_control_tables[(k, name)] = tab
                    lm[key] = float(tab[0][0, 1])
            p.source_controls.append(dict(
                source=k, deliverability="deliverability" in s, productivity=dl.get("productivity"),
                reference_pressure=dl.get("pressure", 1.0e5),
                direction={"both": 0, "production": 1, "out": 1, "injection": 2, "in": 2}[str(s.get("direction", "both")).lower()],
                limit=limits(s)["total"]))
        if "deliverability" in s and "rate" not in s:
            s["rate"] = -1.0          # placeholder: producing, the control sets the rate
    # recharge / injectivity controls (see wb_set_source_recharge; src/source_setup.F90:2925-3092): coefficient (default
    # 1e-2, src/source_control.F90:37), reference pressure (a number, or None = "initial": the pressure of the cell
    # at the start of the run)
    p.source_recharge = []
    for k, s in enumerate(src):
        rc = s.get("recharge", s.get("injectivity"))
        if rc is not None:
            assert np.ndim(rc.get("coefficient", 0.0)) == 0 and not isinstance(rc.get("pressure"), (list, dict)), \
                "table-valued recharge parameters are not built"
            pr = rc.get("pressure", "initial")
            p.source_recharge.append(dict(source=k, coefficient=float(rc.get("coefficient", 1.0e-2)),
                                          reference_pressure=None if isinstance(pr, str) else float(pr)))
            s.setdefault("rate", 0.0)
    # separators and limits on the separated water / steam flows (see wb_set_source_separators)
    p.source_separators = []
    for k, s in enumerate(src):
        pr, lm = separator_pressures(s), limits(s)
        if pr:
            # without a separator the separated flows are zero and a water / steam limit never acts
            # (source_network_node.F90:116-156, separator pressures: source_setup.F90:2255-2328)
            assert len(pr) <= 2, "separators with more than two stages are not built"
            p.source_separators.append(dict(source=k, pressure=pr, limit_water=lm["water"], limit_steam=lm["steam"]))
    # tracer injection rates: numbers or (time, rate) tables per source
    p.source_tracer_tables = {}
    for k, s in enumerate(src):
        tr_ = s.get("tracer")
        if isinstance(tr_, list) and tr_ and isinstance(tr_[0], list):
            p.source_tracer_tables[k] = (np.array(tr_, float), s.get("interpolation", "linear"), s.get("averaging", "integrate"))
            s["tracer"] = 0.0

    p.source_cells = np.array([s["cell"] for s in src], np.int32)
    p.source_rates = np.array([s["rate"] for s in src], float)
    # get_components (src/source_setup.F90:2052-2083; doc/user/setup_sources.rst): injection uses "component"
    # (default water = 1); production uses "production_component", which defaults to energy if "component" is
    # energy and to 0 (all mass components) otherwise
    def comp(v):
        return _component(v, p.np)

    def components(s):
        inj = comp(s.get("component", 1))
        if "production_component" in s:
            prod = comp(s["production_component"])
        else:
            prod = inj if (inj == p.np and p.np > 1 and name_of_eos != "w") else 0
        return inj, prod
    eos_doc = doc.get("eos", "we")
    name_of_eos = eos_doc if isinstance(eos_doc, str) else eos_doc.get("name", "we")
    both = [components(s) for s in src]
    # the reference picks the injection or the production component from the sign of the CURRENT rate at every update
    # (src/source.F90:372-380, 469-476): both travel to the engine (wb_set_source_components); source_components is the
    # choice for the initial rate, for callers that only have fixed-sign sources
    p.source_injection_components = np.array([b[0] for b in both], np.int32)
    p.source_production_components = np.array([b[1] for b in both], np.int32)
    p.source_components = np.array([b[0] if s["rate"] >= 0 else b[1] for s, b in zip(src, both)], np.int32)
    p.source_enthalpies = np.array([s.get("enthalpy", 83.9e3) for s in src], float)
    p.source_tracer = np.array([_tracer_rates(s.get("tracer", 0.0), p.tracers) for s in src], float).reshape(len(src), max(nt, 1))
    p.time = doc.get("time", {})
    return p


def _table_value(tab, interp, t):
    """table%interpolate: linear, or step (right-continuous at the data points); constant beyond both ends"""
    if interp == "step":
        return tab[max(np.searchsorted(tab[:, 0], t, side="right") - 1, 0), 1]
    return np.interp(t, tab[:, 0], tab[:, 1])


def _table_average(tab, interp, t0, t1, averaging="integrate"):
    """table%average over [t0, t1] (src/interpolation.F90:565-680): "integrate" (the default of the JSON input,
    default_averaging_str) integrates the interpolant exactly over the interval, "endpoint" is the mean of the
    values at both ends"""
    if averaging == "endpoint":
        return 0.5 * (_table_value(tab, interp, t0) + _table_value(tab, interp, t1))
    if t1 - t0 < 1.e-15:
        return _table_value(tab, interp, t0)
    xs = np.concatenate([[t0], tab[(tab[:, 0] > t0) & (tab[:, 0] < t1), 0], [t1]])
    total = 0.0
    for a, b in zip(xs[:-1], xs[1:]):
        if interp == "step":
            total += _table_value(tab, interp, a) * (b - a)
        else:
            total += 0.5 * (_table_value(tab, interp, a) + _table_value(tab, interp, b)) * (b - a)
    return total / (t1 - t0)


def tracer_rates_at(p, t0, t1):
    """tracer injection rates [nsources, ntracers] for the time step [t0, t1] (a table applies to every tracer)"""
    r = p.source_tracer.copy()
    for k, (tab, interp, averaging) in p.source_tracer_tables.items():
        r[k, :] = _table_average(tab, interp, t0, t1, averaging)
    return r


def components_at(p, rates):
    """component of every source for rates of these signs (source%update_flow, src/source.F90:372-380, 469-476)"""
    r = np.asarray(rates, float)
    return np.where(r > 0, p.source_injection_components, p.source_production_components).astype(np.int32)


def rates_at(p, t0, t1):
    """source rates for the time step [t0, t1]: fixed rates, and table sources averaged over the step; a "factor"
    control of a fixed-rate source scales its rate (sources on deliverability: see controls_at)"""
    r = p.source_rates.copy()
    for k, (tab, interp, averaging) in p.source_tables.items():
        r[k] = _table_average(tab, interp, t0, t1, averaging)
    on_deliv = {c["source"] for c in getattr(p, "source_controls", []) if c["deliverability"]}
    for (k, key), (tab, interp, averaging) in getattr(p, "source_control_tables", {}).items():
        if key == "factor" and k not in on_deliv:
            r[k] = r[k] * _table_average(tab, interp, t0, t1, averaging)
    return r


def controls_at(p, t0, t1):
    """the source controls and separator limits for the time step [t0, t1]: copies of p.source_controls /
    p.source_separators with every table-valued parameter replaced by its average over the step (table_object_control
    update, src/control.F90 with interpolation_table%average), and the "factor" of a source on deliverability folded
    into its productivity index (the deliverability rate is linear in it).  Pass them to wb_set_source_controls /
    wb_set_source_separators before the step."""
    tabs = getattr(p, "source_control_tables", {})
    avg = lambda k, key: _table_average(*tabs[(k, key)][:2], t0, t1, tabs[(k, key)][2])
    ctrl = [dict(c) for c in p.source_controls]
    for c in ctrl:
        k = c["source"]
        for key in ("productivity", "reference_pressure", "limit"):
            if (k, key) in tabs:
                c[key] = avg(k, key)
        if (k, "factor") in tabs and c["deliverability"] and c["productivity"] is not None:
            c["productivity"] = c["productivity"] * avg(k, "factor")
    seps = [dict(q) for q in getattr(p, "source_separators", [])]
    for q in seps:
        for key in ("limit_water", "limit_steam"):
            if (q["source"], key) in tabs:
                q[key] = avg(q["source"], key)
    return ctrl, seps
