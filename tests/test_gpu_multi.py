"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the partitioned path through the C ABI with the
NCCL halo -- residual, SpMV, Krylov solve and a Newton solve on 2 ranks -- against the single-GPU path on the same
problem.  One process per GPU, rendezvous on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DIMS = (12, 10, 16)
DT = 1.0e6
DT_KSP = 1.0e3


def relerr_(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def problem(kind="we"):
    """we: structured eos_we; wce: eos_wce (3 primaries, halo width 4); minc: eos_we on a MINC mesh with two matrix
    levels (irregular rows, matrix cells owned by their fracture cell's rank)"""
    from waiwera_b200 import mesh as wmesh
    gm = wmesh.structured(*DIMS, dx=10.0, seed=wmesh.SEED)
    if kind == "wce":
        primary, region = wmesh.wce_state(gm, seed=wmesh.SEED)
    else:
        primary, region = wmesh.hydrostatic_state(gm, seed=wmesh.SEED)
    if kind == "minc":
        n = gm.ninterior
        gm = wmesh.add_minc(gm, volumes=(0.1, 0.3, 0.6), spacing=(50., 50., 50.), matrix_permeability_factor=0.01)
        rng = np.random.default_rng(11)
        prims = [primary]
        for _ in range(2):
            pm = primary.copy()
            pm[:, 0] *= 1.0 + 1e-3 * rng.uniform(-1, 1, n)
            prims.append(pm)
        primary, region = np.concatenate(prims), np.concatenate([region] * 3)
    y = np.ascontiguousarray(wmesh.scale_primaries(primary, region)).reshape(-1)
    return gm, y, region


def params_for(kind):
    from waiwera_b200 import flow
    return flow.make_params(eos=flow.EOS_WCE if kind == "wce" else flow.EOS_WE)


def owner_for(kind, gm, world):
    from waiwera_b200 import mesh as wmesh
    parts = wmesh.default_parts(world)
    return wmesh.minc_owner(gm, parts) if kind == "minc" else wmesh.box_owner(gm, parts)


NT = 2


def tracer_step(sim, flow, gm, y, nat):
    """one auxiliary tracer solve (liquid tracer with diffusion, decaying liquid tracer) at the state y: the
    partitioned system (ghost columns through the halo inside the SpMV) must give the single-GPU solution"""
    err, _ = sim.lhs(y)                    # unperturbed evaluation: the state the tracer system is built from
    assert err == 0
    assert sim.set_tracers([1, 1], diffusion=[1.0e-6, 0.0], decay=[0.0, 1.0e-7]) == 0
    x0 = np.random.default_rng(5).uniform(0.0, 0.01, gm.ninterior * NT).reshape(-1, NT)[nat].reshape(-1)
    al = sim.tracer_balances()
    x, _, reason, its = sim.tracer_solve(DT, al, np.ascontiguousarray(x0), opts=flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-12),
                                         pc_type=flow.PC_PBJACOBI)
    assert reason > 0
    return x


def worker(rank, world, port, out, p2p, kind="we"):
    import torch.distributed as dist
    from waiwera_b200 import flow, mesh as wmesh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        gm, gy, gregion = problem(kind)
        npv = 3 if kind == "wce" else 2
        owner = owner_for(kind, gm, world)
        m = wmesh.partition(gm, owner, rank, world)
        nat = m.natural[:m.nowned]
        y = np.ascontiguousarray(gy.reshape(-1, npv)[nat].reshape(-1))
        region = np.ascontiguousarray(gregion[nat])
        sim = flow.FlowSimulation(params_for(kind), m, device=rank)
        uid = torch.from_numpy(flow.FlowSimulation.unique_id()).cuda() if rank == 0 else torch.zeros(128, dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        sim.comm_init(rank, world, uid.cpu().numpy())
        if p2p:
            assert sim.p2p_setup(dist, torch.device("cuda", rank))
        assert sim.fluid_init(y, region) == 0
        err, L0 = sim.lhs(y)
        assert err == 0
        y1 = y * (1.0 + 1e-5)
        err, lhs, rhs, r = sim.residual(y1, L0, DT)
        assert err == 0
        mv, ml = sim.max_scaled(r, L0, 1.0)
        assert sim.jacobian(y1, L0, DT) == 0
        J = sim.jacobian_mat()
        x = np.random.default_rng(3).uniform(-1, 1, gm.ninterior * npv).reshape(-1, npv)[nat].reshape(-1)
        ax = np.zeros_like(x)
        J.mult(x, ax)
        # the benchmark's solver configuration: GMRES + block Jacobi over ILU(0) cube sub-domains (with the NVLink
        # path: the persistent kernel, halo / dots / norm exchanged inside it)
        # (the Newton system of a short time step, DT_KSP: converges in tens of iterations, so iteration counts are
        # comparable between partitions; at DT restarted GMRES stagnates and the count moves with rounding)
        err, _, _, rk = sim.residual(y1, L0, DT_KSP)
        assert err == 0 and sim.jacobian(y1, L0, DT_KSP) == 0
        bor = wmesh.minc_cube_blocks(m, 4) if kind == "minc" else wmesh.cube_blocks(m, 4)
        pck = flow.PC(J, flow.PC_BJACOBI_ILU0, 1, bor)
        xk = np.zeros_like(x)
        kreason, kits, _ = flow.ksp_solve(J, pck, rk, xk, flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-10))
        xk2 = np.zeros_like(x)
        k2 = flow.ksp_solve(J, pck, rk, xk2, flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-10, restart=7))
        xk3 = np.zeros_like(x)   # the reference's default Krylov method, also inside the persistent kernel
        k3 = flow.ksp_solve(J, pck, rk, xk3, flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-10))
        pck.destroy()
        assert sim.jacobian(y1, L0, DT) == 0
        y2 = y.copy()
        res = sim.newton_solve(y2, L0, DT, flow.newton_opts(max_iterations=4, pc_type=flow.PC_PBJACOBI,
                                                            ksp=flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-10)))
        xt = tracer_step(sim, flow, gm, y1, nat)
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(nat=nat, r=r, ax=ax, y2=y2, mv=mv, ml=ml, reason=res.reason,
                                              its=res.iterations, lits=res.linear_iterations, xt=xt, xk=xk,
                                              kreason=kreason, kits=kits, xk2=xk2, k2=k2[:2], xk3=xk3, k3=k3[:2]))
        if rank == 0:
            out.put(gathered)
        sim.destroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p2p,kind", [(2, False, "we"), (2, True, "we"), (4, False, "we"), (4, True, "we"),
                                            (2, True, "wce"), (2, True, "minc")])
def test_partitioned_path_matches_single_gpu(world, p2p, kind):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d CUDA devices" % world)
    import torch.multiprocessing as mp
    from waiwera_b200 import flow
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, out, p2p, kind)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single-GPU reference on the whole mesh
    gm, gy, gregion = problem(kind)
    npv = 3 if kind == "wce" else 2
    sim = flow.FlowSimulation(params_for(kind), gm, device=0)
    assert sim.fluid_init(gy, gregion) == 0
    err, L0 = sim.lhs(gy)
    y1 = gy * (1.0 + 1e-5)
    err, lhs, rhs, r = sim.residual(y1, L0, DT)
    mv, ml = sim.max_scaled(r, L0, 1.0)
    assert sim.jacobian(y1, L0, DT) == 0
    J = sim.jacobian_mat()
    x = np.random.default_rng(3).uniform(-1, 1, gm.ninterior * npv)
    ax = np.zeros_like(x)
    J.mult(x, ax)
    # the same sub-domains on one GPU: cubes cut at the partition boundaries
    from waiwera_b200 import mesh as wmesh
    owner = owner_for(kind, gm, world)
    cube = wmesh.minc_cube_blocks(gm, 4) if kind == "minc" else wmesh.cube_blocks(gm, 4)
    _, bor = np.unique(owner.astype(np.int64) * (cube.max() + 1) + cube, return_inverse=True)
    err, _, _, rk = sim.residual(y1, L0, DT_KSP)
    assert err == 0 and sim.jacobian(y1, L0, DT_KSP) == 0
    pck = flow.PC(J, flow.PC_BJACOBI_ILU0, 1, bor.astype(np.int32))
    xk = np.zeros_like(x)
    kreason, kits, _ = flow.ksp_solve(J, pck, rk, xk, flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-10))
    xk2 = np.zeros_like(x)
    k2 = flow.ksp_solve(J, pck, rk, xk2, flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-10, restart=7))
    xk3 = np.zeros_like(x)
    k3 = flow.ksp_solve(J, pck, rk, xk3, flow.ksp_opts(type=flow.KSP_BCGS, rtol=1e-10))
    pck.destroy()
    assert kits < 200, kits
    assert sim.jacobian(y1, L0, DT) == 0
    gxk, gxk2, gxk3 = np.zeros_like(xk).reshape(-1, npv), np.zeros_like(xk).reshape(-1, npv), np.zeros_like(xk).reshape(-1, npv)
    for g in gathered:
        gxk[g["nat"]] = g["xk"].reshape(-1, npv)
        gxk2[g["nat"]] = g["xk2"].reshape(-1, npv)
        gxk3[g["nat"]] = g["xk3"].reshape(-1, npv)
        assert g["k3"][0] == k3[0] > 0 and abs(g["k3"][1] - k3[1]) <= 3, (g["k3"], k3[:2])
        assert g["kreason"] == kreason > 0 and abs(g["kits"] - kits) <= 3, (g["kreason"], g["kits"], kreason, kits)
        assert g["k2"][0] == k2[0] > 0 and abs(g["k2"][1] - k2[1]) <= 6, (g["k2"], k2[:2])
    assert relerr_(gxk.reshape(-1), xk) < 1e-7 and relerr_(gxk2.reshape(-1), xk2) < 1e-7
    assert relerr_(gxk3.reshape(-1), xk3) < 1e-7
    y2 = gy.copy()
    res = sim.newton_solve(y2, L0, DT, flow.newton_opts(max_iterations=4, pc_type=flow.PC_PBJACOBI,
                                                        ksp=flow.ksp_opts(type=flow.KSP_GMRES, rtol=1e-10)))
    xt = tracer_step(sim, flow, gm, y1, np.arange(gm.ninterior))
    gxt = np.zeros_like(xt).reshape(-1, NT)
    for g in gathered:
        gxt[g["nat"]] = g["xt"].reshape(-1, NT)
    assert np.abs(gxt.reshape(-1) - xt).max() <= 1e-9 * np.abs(xt).max()
    gr, gax, gy2 = np.zeros_like(r).reshape(-1, npv), np.zeros_like(ax).reshape(-1, npv), np.zeros_like(y2).reshape(-1, npv)
    for g in gathered:
        gr[g["nat"]] = g["r"].reshape(-1, npv)
        gax[g["nat"]] = g["ax"].reshape(-1, npv)
        gy2[g["nat"]] = g["y2"].reshape(-1, npv)
    # residual: same faces, same summation order -> bit exact; SpMV: local column order differs -> rounding
    assert np.array_equal(gr.reshape(-1), r)
    assert np.abs(gax.reshape(-1) - ax).max() <= 1e-13 * np.abs(ax).max()
    assert all(abs(g["mv"] - mv) <= 1e-14 * abs(mv) for g in gathered)
    assert all(g["reason"] == res.reason and g["its"] == res.iterations for g in gathered)
    assert np.abs(gy2.reshape(-1) - y2).max() <= 1e-8 * np.abs(y2).max()
    sim.destroy()
