"""Host-side checks of the bench problems (bench.Problem): the weak-scaling layout of config 5 (MINC, one box of
fracture cells + their matrix cells per GPU, boxes side by side so that every N has the same depth), its partition and
halo plan, and the sub-domain (cube) assignment the preconditioner uses.  CPU only: nothing here touches the GPU library
or the oracle."""
import argparse

import numpy as np
import pytest

import bench
from waiwera_b200 import mesh as wmesh

PER = (6, 5, 4)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_config5_weak_layout_and_partition(world):
    p = bench.Problem(5, world, PER)
    nfrac = PER[0] * PER[1] * PER[2] * world
    assert p.scaling == "weak" and p.minc
    assert p.dims[2] == PER[2], "boxes are placed side by side: the depth (the hydrostatic column) must not change with N"
    assert p.dims[0] * p.dims[1] * p.dims[2] == nfrac
    assert p.mesh.ninterior == 2 * nfrac                     # one matrix level: as many matrix cells as fracture cells
    assert len(p.y) == 3 * p.mesh.ninterior and len(p.region) == p.mesh.ninterior
    name = p.name(argparse.Namespace(ksp="gmres", restart=30, pc="ilu0"))
    assert "config 5" in name and "MINC" in name and "%d cells" % p.mesh.ninterior in name
    if world == 1:
        assert p.owner is None
        return
    owner = p.owner
    # equal work per GPU, matrix cells with their fracture cell
    assert np.array_equal(np.bincount(owner, minlength=world), np.full(world, 2 * nfrac // world))
    assert np.array_equal(owner[:nfrac], owner[nfrac:2 * nfrac])
    ms = [wmesh.partition(p.mesh, owner, r, world) for r in range(world)]
    seen = np.zeros(p.mesh.ninterior, int)
    for m in ms:
        seen[m.natural[:m.nowned]] += 1
        for n, r in enumerate(m.neigh_rank):
            o = ms[r]
            k = list(o.neigh_rank).index(m.rank)
            assert np.array_equal(m.natural[m.send_idx[m.send_ptr[n]:m.send_ptr[n + 1]]],
                                  o.natural[o.recv_idx[o.recv_ptr[k]:o.recv_ptr[k + 1]]])
        # sub-domains: a matrix cell is in the sub-domain of its fracture cell; ids are dense
        blk = p.blocks(m, 2)
        assert blk.min() == 0 and len(np.unique(blk)) == blk.max() + 1
        nat = m.natural[:m.nowned]
        frac_of = np.where(nat >= nfrac, nat - nfrac, nat)
        by_frac = {}
        for b, fcell in zip(blk, frac_of):
            assert by_frac.setdefault(int(fcell), int(b)) == int(b)
    assert (seen == 1).all()


@pytest.mark.parametrize("cfg,dims", [(2, (6, 6, 6)), (4, (6, 6, 10))])
def test_strong_scaling_problems_split_evenly(cfg, dims):
    for world in (2, 4, 8):
        p = bench.Problem(cfg, world, dims)
        assert p.scaling == "strong" and p.mesh.ninterior == dims[0] * dims[1] * dims[2]
        counts = np.bincount(p.owner, minlength=world)
        assert counts.sum() == p.mesh.ninterior and counts.max() - counts.min() <= counts.max() // 2


def test_solver_bytes_per_iteration():
    """the algorithmic byte count behind bench.py's `roofline_solver`: config 2's pattern at a small size against a
    count by hand, and the full-size formula against the figure DESIGN.md quotes (1.09 GB per GMRES(30) iteration)"""
    nx = 4
    m = wmesh.structured(nx, nx, nx)
    nb = m.ninterior
    rows = [[i] for i in range(nb)]
    for c1, c2 in m.face_cells:
        if c1 < nb and c2 < nb:
            rows[c1].append(int(c2))
            rows[c2].append(int(c1))
    rowptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    colidx = np.concatenate([sorted(r) for r in rows])
    nnzb = len(colidx)
    assert nnzb == nb + 2 * 3 * nx * nx * (nx - 1)
    bor = wmesh.cube_blocks(m, 2)
    # couplings cut by 2^3 cubes: the faces between cubes, one plane per direction in a 4^3 mesh
    nnzf = nnzb - 2 * 3 * nx * nx
    got = bench.solver_bytes_per_iteration(nb, 2, rowptr, colidx, bor, 30, "gmres")
    assert got == (nnzb + nnzf) * 36 + 37 * nb * 16
    assert bench.solver_bytes_per_iteration(nb, 2, rowptr, colidx, None, 30, "gmres") == 2 * nnzb * 36 + 37 * nb * 16
    assert bench.solver_bytes_per_iteration(nb, 3, rowptr, colidx, bor, 30, "bcgs") == 2 * (nnzb + nnzf) * 76 + 22 * nb * 24
    # full size: 100^3 cells, 10^3 cubes
    N = 100
    nnz = N ** 3 + 6 * N * N * (N - 1)
    nnf = nnz - 2 * 3 * N * N * 9
    assert abs(((nnz + nnf) * 36 + 37 * N ** 3 * 16) / 1e9 - 1.072) < 0.001   # 1.07-1.09 GB, ncu: 1.11 GB
