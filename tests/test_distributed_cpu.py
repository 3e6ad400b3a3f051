"""World-size-2 checks of the N>1 host logic on CPU (gloo, 127.0.0.1): the box partition with overlap-1 ghost
cells and the halo plan that `wb_set_halo` consumes.  Each rank builds its local mesh, exchanges ghost entries
exactly as the SpMV halo does (send owned cells listed in send_idx, receive into the ghost cells of recv_idx),
and the distributed block SpMV / dot product reproduce the serial ones.  (reference:
DMPlexDistribute overlap 1 src/mesh.F90:143-171, MatMult_MPIBAIJ ghost scatter.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from waiwera_b200 import mesh as wmesh

DIMS = (6, 5, 8)
BS = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def block_value(i, j):
    """deterministic bs x bs block of the global matrix entry (i, j), natural numbering"""
    rng = np.random.default_rng([int(i), int(j), 77])
    b = rng.uniform(-1, 1, (BS, BS))
    if i == j:
        b += 8.0 * np.eye(BS)
    return b


def pattern(m):
    """FV adjacency rows over owned cells: self + face neighbours with dofs (src/dm_utils.F90:1041-1051),
    local column numbering, sorted"""
    rows = [[i] for i in range(m.nowned)]
    for c1, c2 in m.face_cells:
        if c1 < m.ninterior and c2 < m.ninterior:
            if c1 < m.nowned:
                rows[c1].append(int(c2))
            if c2 < m.nowned:
                rows[c2].append(int(c1))
    return [sorted(r) for r in rows]


def local_spmv(m, xloc):
    y = np.zeros((m.nowned, BS))
    for i, cols in enumerate(pattern(m)):
        for j in cols:
            y[i] += block_value(m.natural[i], m.natural[j]) @ xloc[j]
    return y


def halo_exchange(m, xloc):
    """ghost entries of xloc <- owners, in the message order of the halo plan"""
    reqs, bufs = [], []
    for n, r in enumerate(m.neigh_rank):
        s = torch.from_numpy(np.ascontiguousarray(xloc[m.send_idx[m.send_ptr[n]:m.send_ptr[n + 1]]]))
        rb = torch.zeros((m.recv_ptr[n + 1] - m.recv_ptr[n], BS), dtype=torch.float64)
        reqs.append(dist.isend(s, int(r)))
        reqs.append(dist.irecv(rb, int(r)))
        bufs.append((n, rb))
    for q in reqs:
        q.wait()
    for n, rb in bufs:
        xloc[m.recv_idx[m.recv_ptr[n]:m.recv_ptr[n + 1]]] = rb.numpy()


def worker(rank, world, port, parts, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gm = wmesh.structured(*DIMS, top_boundary=True)
        owner = wmesh.box_owner(gm, parts)
        m = wmesh.partition(gm, owner, rank, world)
        gx = np.random.default_rng(5).uniform(-1, 1, (gm.ninterior, BS))
        # ---- structure of the local mesh
        assert m.nowned == int((owner == rank).sum())
        assert np.array_equal(m.natural[:m.nowned], np.flatnonzero(owner == rank))
        assert (owner[m.natural[m.nowned:m.ninterior]] != rank).all()
        # ghost cells of the receive list are numbered contiguously after the owned cells (no unpack needed)
        assert np.array_equal(m.recv_idx, m.nowned + np.arange(m.ninterior - m.nowned))
        assert (m.send_idx < m.nowned).all()
        # local geometry = global geometry under the natural map; local faces keep the global face order
        assert np.array_equal(m.cell_geom[:m.ninterior], gm.cell_geom[m.natural])
        assert np.array_equal(m.rock[:m.ninterior], gm.rock[m.natural])
        # ---- halo + SpMV
        xloc = np.zeros((m.ninterior, BS))
        xloc[:m.nowned] = gx[m.natural[:m.nowned]]
        halo_exchange(m, xloc)
        assert np.array_equal(xloc, gx[m.natural]), "ghost entries differ from their owners' values"
        y = local_spmv(m, xloc)
        # ---- global dot product (VecMDot: local partial + allreduce)
        d = torch.tensor([float((xloc[:m.nowned] * y).sum())], dtype=torch.float64)
        dist.all_reduce(d)
        ys = [None] * world
        dist.all_gather_object(ys, (m.natural[:m.nowned], y))
        if rank == 0:
            gy = np.zeros((gm.ninterior, BS))
            for nat, yy in ys:
                gy[nat] = yy
            ref = local_spmv(gm, gx)
            assert np.abs(gy - ref).max() <= 1e-13 * np.abs(ref).max()
            assert abs(d.item() - (gx * ref).sum()) <= 1e-12 * abs((gx * ref).sum())
            # the ranks' first_cell offsets tile the global (rank-contiguous) numbering
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,parts", [(2, (1, 1, 2)), (2, (2, 1, 1)), (4, (1, 2, 2))])
def test_partition_halo_spmv_gloo(world, parts):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, parts, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_partition_offsets_and_faces():
    """first_cell offsets tile [0, N); every global face is kept by the ranks owning its support cells, and the
    union of the ranks' owned cells is a partition"""
    gm = wmesh.structured(*DIMS, top_boundary=True)
    for world in (2, 4, 8):
        owner = wmesh.box_owner(gm, wmesh.default_parts(world))
        ms = [wmesh.partition(gm, owner, r, world) for r in range(world)]
        assert sum(m.nowned for m in ms) == gm.ninterior
        assert [m.first_cell for m in ms] == list(np.cumsum([0] + [m.nowned for m in ms[:-1]]))
        seen = np.zeros(gm.ninterior, int)
        for m in ms:
            seen[m.natural[:m.nowned]] += 1
            # every neighbour pair agrees on message sizes
            for n, r in enumerate(m.neigh_rank):
                o = ms[r]
                k = list(o.neigh_rank).index(m.rank)
                assert m.send_ptr[n + 1] - m.send_ptr[n] == o.recv_ptr[k + 1] - o.recv_ptr[k]
                # and on the cells: what I send is what the neighbour's ghost list expects, in order
                assert np.array_equal(m.natural[m.send_idx[m.send_ptr[n]:m.send_ptr[n + 1]]],
                                      o.natural[o.recv_idx[o.recv_ptr[k]:o.recv_ptr[k + 1]]])
        assert (seen == 1).all()


@pytest.mark.parametrize("deck,world", [("minc_3d_refined", 2), ("minc_3d_refined", 3), ("minc_3d_refined", 4), ("problem5b", 2),
                                        ("problem6", 8)])
def test_unstructured_meshes_partition_like_the_boxes(deck, world):
    """an ingested deck (3-D hybrid mesh with a MINC zone and boundary faces; a 2-D areal mesh; the 125-cell 3-D field
    on 8 ranks) split by mesh.coordinate_owner: balanced, matrix cells with their fracture cell, halo plans of
    neighbouring ranks agree, and the function evaluation of every rank on its local mesh -- ghost values filled as the
    halo exchange would -- gives the rows of the serial one (oracle, bit for bit apart from the order of the inflow sum)"""
    from oracle import wo
    from waiwera_b200 import ingest
    here = os.path.dirname(os.path.abspath(__file__))
    p = ingest.load(os.path.join(here, "golden", "inputs", deck + ".input.json"), mod=wo)
    gm = p.mesh
    n = gm.ninterior
    owner = wmesh.coordinate_owner(gm, world)
    counts = np.bincount(owner, minlength=world)
    assert counts.min() > 0 and counts.max() <= 1.35 * n / world + 3
    if hasattr(gm, "minc_parent"):
        assert np.array_equal(owner, owner[gm.minc_parent])
    ms = [wmesh.partition(gm, owner, r, world) for r in range(world)]
    assert sum(m.nowned for m in ms) == n
    for m in ms:
        for k, r in enumerate(m.neigh_rank):
            o = ms[r]
            j = list(o.neigh_rank).index(m.rank)
            assert np.array_equal(m.natural[m.send_idx[m.send_ptr[k]:m.send_ptr[k + 1]]],
                                  o.natural[o.recv_idx[o.recv_ptr[j]:o.recv_ptr[j + 1]]])
    npv = p.np

    def evaluate(m, y, region, boundary_rows, ghost_fluid=None):
        f = wo.Flow(p.params, m.ncell, m.ninterior, m.nowned, m.face_cells.reshape(-1), m.face_geom.reshape(-1),
                    m.cell_geom.reshape(-1), m.rock.reshape(-1))
        for gl, il, row in boundary_rows:
            assert f.set_boundary(int(gl), int(il), p.boundary_primary[row], int(p.boundary_region[row])) == 0
        assert f.fluid_init(y, region) == 0
        if ghost_fluid is not None:          # the fluid vector's global-to-local scatter (the halo exchange)
            f.fluid()[m.nowned:m.ninterior] = ghost_fluid
        err, L0 = f.lhs(y)
        assert err == 0
        err, lhs, rhs, r = f.residual(y, L0, 1.0e5)
        assert err == 0
        return lhs, rhs, f.fluid()[:m.ninterior].copy()

    rows = [(gm.boundary["ghost_cells"][k], gm.boundary["interior_cells"][k], k) for k in range(len(p.boundary_region))]
    lhs_g, rhs_g, fluid_g = evaluate(gm, p.y, p.region, rows)
    yg = p.y.reshape(n, npv)
    row_of_ghost = {int(g): k for k, g in enumerate(gm.boundary.get("ghost_cells", []))}
    for m in ms:
        y = np.ascontiguousarray(yg[m.natural]).reshape(-1)          # owned + partition ghosts, as after the halo exchange
        region = np.ascontiguousarray(p.region[m.natural])
        brows = []
        if m.boundary:
            brows = [(gl, il, row_of_ghost[int(gg)]) for gl, il, gg in
                     zip(m.boundary["ghost_cells"], m.boundary["interior_cells"], m.boundary["global_ghost"])]
        lhs, rhs, _ = evaluate(m, y, region, brows, ghost_fluid=fluid_g[m.natural[m.nowned:m.ninterior]])
        own = m.natural[:m.nowned]
        assert np.array_equal(lhs.reshape(-1, npv)[:m.nowned], lhs_g.reshape(-1, npv)[own])
        a, b = rhs.reshape(-1, npv)[:m.nowned], rhs_g.reshape(-1, npv)[own]
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
