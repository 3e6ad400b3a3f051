"""The linear-algebra half of the oracle (BAIJ SpMV, block ILU(0), GMRES(30), BiCGStab) is PETSc code that is not in
the reference tree, so no reference vector pins it ("parity unpinned").  These CPU tests pin it
against INDEPENDENT implementations instead: scipy's BSR product and Krylov solvers, and a dense textbook block
ILU(0) written here from the definition (L U = A on the sparsity pattern)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from waiwera_b200 import mesh as wmesh

SEED = 20240917


def random_system(wo, dims, bs, seed, diag_boost=4.0):
    m = wmesh.structured(*dims)
    wm = wo.Mesh()
    keep = (np.ascontiguousarray(m.face_cells.reshape(-1), np.int32), m.face_geom.reshape(-1).copy(),
            m.cell_geom.reshape(-1).copy(), m.rock.reshape(-1).copy())
    wm.ncell, wm.ninterior, wm.nowned, wm.nface = m.ncell, m.ninterior, m.nowned, m.nface
    wm.face_cells, wm.face_geom, wm.cell_geom, wm.rock = wo.ip(keep[0]), wo.dp(keep[1]), wo.dp(keep[2]), wo.dp(keep[3])
    A = wo.lib().wo_bsr_from_mesh(C.byref(wm), bs)
    rowptr, colidx, val = wo.bsr_arrays(A)
    rng = np.random.default_rng(seed)
    val[:] = rng.uniform(-1, 1, val.shape)
    rows = np.repeat(np.arange(A.contents.nb), np.diff(rowptr))
    diag = np.flatnonzero(colidx == rows)
    val[diag] += diag_boost * np.eye(bs).reshape(-1)
    # scipy BSR wants row-major blocks; BAIJ blocks are column-major
    blocks = val.reshape(-1, bs, bs).transpose(0, 2, 1).copy()
    S = sp.bsr_matrix((blocks, colidx.copy(), rowptr.copy()), shape=(A.contents.nb * bs,) * 2)
    return A, S, keep


def dense_block_ilu0(S, bs, block_of_row=None):
    """textbook IKJ ILU(0) on the block pattern (dense arithmetic on a small matrix); couplings between different
    block-Jacobi sub-domains are dropped first"""
    Ad = S.toarray()
    nb = Ad.shape[0] // bs
    pat = np.zeros((nb, nb), bool)
    B = S.tobsr(blocksize=(bs, bs))
    for i in range(nb):
        for k in range(B.indptr[i], B.indptr[i + 1]):
            j = B.indices[k]
            if block_of_row is None or block_of_row[i] == block_of_row[j]:
                pat[i, j] = True
    blk = lambda M, i, j: M[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
    LU = np.zeros_like(Ad)
    for i in range(nb):
        for j in range(nb):
            if pat[i, j]:
                blk(LU, i, j)[:] = blk(Ad, i, j)
    for i in range(nb):
        for k in range(i):
            if not pat[i, k]:
                continue
            blk(LU, i, k)[:] = blk(LU, i, k) @ np.linalg.inv(blk(LU, k, k))
            for j in range(k + 1, nb):
                if pat[i, j] and pat[k, j]:
                    blk(LU, i, j)[:] -= blk(LU, i, k) @ blk(LU, k, j)
    L = np.tril(LU, -1)
    U = np.triu(LU)
    # block lower / upper split: the diagonal blocks belong to U entirely
    for i in range(nb):
        d = blk(LU, i, i).copy()
        blk(L, i, i)[:] = np.eye(bs)
        blk(U, i, i)[:] = d
    return L, U, pat


@pytest.mark.parametrize("bs", [1, 2, 3])
def test_spmv_matches_scipy(wo, bs):
    A, S, keep = random_system(wo, (6, 5, 4), bs, SEED + bs)
    x = np.random.default_rng(1).uniform(-1, 1, S.shape[0])
    y = np.zeros_like(x)
    wo.lib().wo_bsr_spmv(A, wo.dp(x), wo.dp(y))
    ref = S @ x
    assert np.abs(y - ref).max() <= 1e-14 * np.abs(ref).max()
    wo.lib().wo_bsr_destroy(A)


@pytest.mark.parametrize("bs,nblocks", [(2, 1), (3, 1), (2, 4)])
def test_ilu0_matches_textbook_definition(wo, bs, nblocks):
    """M^-1 r of the oracle == U^-1 L^-1 r of a dense ILU(0) built from the definition, and (L U) == A on the pattern"""
    A, S, keep = random_system(wo, (4, 3, 3), bs, SEED + 10 * bs + nblocks)
    nb = A.contents.nb
    bor = None if nblocks == 1 else ((np.arange(nb, dtype=np.int64) * nblocks) // nb).astype(np.int32)
    L, U, pat = dense_block_ilu0(S, bs, bor)
    Ad = S.toarray()
    LU = L @ U
    for i in range(nb):
        for j in range(nb):
            if pat[i, j]:
                assert np.allclose(LU[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs], Ad[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs],
                                   rtol=1e-11, atol=1e-12)
    pc = wo.lib().wo_pc_create(A, wo.PC_BJACOBI_ILU0, wo.ip(bor))
    r = np.random.default_rng(2).uniform(-1, 1, nb * bs)
    z = np.zeros_like(r)
    wo.lib().wo_pc_apply(pc, wo.dp(r), wo.dp(z))
    ref = np.linalg.solve(U, np.linalg.solve(L, r))
    assert np.abs(z - ref).max() <= 1e-11 * np.abs(ref).max()
    wo.lib().wo_pc_destroy(pc)
    wo.lib().wo_bsr_destroy(A)


@pytest.mark.parametrize("ksp", ["gmres", "bcgs"])
def test_krylov_matches_scipy(wo, ksp):
    """same preconditioned system solved by scipy (restarted GMRES(30) / BiCGStab): both reach the tolerance, the
    solutions agree to it, and the iteration counts are of the same size"""
    bs = 2
    A, S, keep = random_system(wo, (8, 7, 6), bs, SEED + 77, diag_boost=3.0)
    nb = A.contents.nb
    n = nb * bs
    pc = wo.lib().wo_pc_create(A, wo.PC_BJACOBI_ILU0, None)
    b = np.random.default_rng(3).uniform(-1, 1, n)

    def apply_pc(r):
        z = np.zeros(n)
        wo.lib().wo_pc_apply(pc, wo.dp(np.ascontiguousarray(r, np.float64)), wo.dp(z))
        return z

    M = spla.LinearOperator((n, n), matvec=apply_pc)
    o = wo.KspOpts()
    o.type = wo.KSP_GMRES if ksp == "gmres" else wo.KSP_BCGS
    o.restart, o.maxit, o.rtol, o.atol, o.dtol = 30, 10000, 1e-10, 1e-50, 1e5
    x = np.zeros(n)
    its, rn = C.c_int(), C.c_double()
    reason = wo.lib().wo_ksp_solve(A, pc, C.byref(o), wo.dp(b), wo.dp(x), C.byref(its), C.byref(rn))
    assert reason > 0
    count = [0]

    def cb(_):
        count[0] += 1
    if ksp == "gmres":
        xs, info = spla.gmres(S, b, M=M, restart=30, rtol=1e-12, atol=0.0, maxiter=500, callback=cb, callback_type="pr_norm")
    else:
        xs, info = spla.bicgstab(S, b, M=M, rtol=1e-12, atol=0.0, maxiter=2000, callback=cb)
    assert info == 0
    assert np.linalg.norm(S @ x - b) <= 1e-7 * np.linalg.norm(b)
    assert np.abs(x - xs).max() <= 1e-7 * np.abs(xs).max()
    assert 0.3 * count[0] <= its.value <= 3 * max(count[0], 1), (its.value, count[0])
    wo.lib().wo_pc_destroy(pc)
    wo.lib().wo_bsr_destroy(A)


@pytest.mark.parametrize("bs,box", [(2, (2, 3, 3)), (1, (4, 3, 1)), (3, (2, 2, 2))])
def test_asm_overlap1_matches_definition(wo, bs, box):
    """restricted additive Schwarz with overlap 1 from its definition: for every sub-domain, the dense ILU(0) of the
    matrix restricted to the sub-domain plus one layer of matrix neighbours (rows in ascending order), applied to the
    restricted residual; only the owned rows are kept"""
    dims = (4, 3, 3)
    A, S, keep = random_system(wo, dims, bs, SEED + 40 + bs)
    nb = A.contents.nb
    idx = np.arange(nb)
    i, j, k = idx % dims[0], (idx // dims[0]) % dims[1], idx // (dims[0] * dims[1])
    key = (i // box[0]) + 10 * (j // box[1]) + 100 * (k // box[2])
    bor = np.unique(key, return_inverse=True)[1].astype(np.int32)
    B = S.tobsr(blocksize=(bs, bs))
    r = np.random.default_rng(5).uniform(-1, 1, nb * bs)
    ref = np.zeros_like(r)
    for b in range(bor.max() + 1):
        own = np.flatnonzero(bor == b)
        ext = sorted(set(own) | {int(B.indices[q]) for row in own for q in range(B.indptr[row], B.indptr[row + 1])})
        assert len(ext) > len(own)
        dof = np.concatenate([np.arange(e * bs, (e + 1) * bs) for e in ext])
        Sb = sp.csr_matrix(S.toarray()[np.ix_(dof, dof)])
        L, U, _ = dense_block_ilu0(Sb, bs)
        zb = np.linalg.solve(U, np.linalg.solve(L, r[dof]))
        for pos, e in enumerate(ext):
            if bor[e] == b:
                ref[e * bs:(e + 1) * bs] = zb[pos * bs:(pos + 1) * bs]
    pc = wo.lib().wo_pc_create(A, wo.PC_ASM_ILU0, wo.ip(bor))
    assert pc
    z = np.zeros_like(r)
    wo.lib().wo_pc_apply(pc, wo.dp(r), wo.dp(z))
    assert np.abs(z - ref).max() <= 1e-11 * np.abs(ref).max()
    # one sub-domain: the extension adds nothing, ASM == global ILU(0)
    pc1 = wo.lib().wo_pc_create(A, wo.PC_ASM_ILU0, None)
    pc2 = wo.lib().wo_pc_create(A, wo.PC_BJACOBI_ILU0, None)
    z1, z2 = np.zeros_like(r), np.zeros_like(r)
    wo.lib().wo_pc_apply(pc1, wo.dp(r), wo.dp(z1))
    wo.lib().wo_pc_apply(pc2, wo.dp(r), wo.dp(z2))
    assert np.array_equal(z1, z2)
    for p in (pc, pc1, pc2):
        wo.lib().wo_pc_destroy(p)
    wo.lib().wo_bsr_destroy(A)
