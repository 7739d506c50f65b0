"""N>1 path on the host: rank THREADS drive one emulated ctx each (tests/cuda_emu; NCCL replaced by an in-process rendezvous).

TEST INFRASTRUCTURE (see tests/emu_support.py): checks the logic of the partitioned path — RCB partition, interface maps,
sub-assembled K, owner-masked dots, single-reduction CG recurrence, gathers back to the reference's DOF order — against the
single-ctx result, with 0xFF-filled allocations so that any read-before-write shows.  The real gate is tests/test_gpu_y_dist.py on
2/4 B200s."""
import os
import sys
import threading

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_support  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    pkg, lib = emu_support.load_emu()
    with emu_support.emulated(pkg, lib):
        yield pkg, lib
    assert lib.emu_check_all_guards() == 0


def _problem(pkg, dims, simp):
    pts, cells = pkg.meshgen.cantilever(*dims)
    rho = pkg.meshgen.simp_like_density(cells.shape[0]) if simp else None
    fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
    load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
    return pts, cells, rho, fixed, load


def _run(pkg, ctx, prob, distributed, mf, tol=1e-10, two_level=False):
    pts, cells, rho, fixed, load = prob
    lam, mu = pkg.create_material_model(1.0, 0.3)
    ctx.set_mesh(pts, cells, distributed=distributed)
    ctx.build_dofs(); ctx.build_pattern()
    if rho is not None:
        (ctx.set_material_simp if mf else ctx.assemble_simp)(1.0, 0.3, 1e-8, 3.0, rho)
    else:
        (ctx.set_material_lame if mf else ctx.assemble_lame)(lam, mu)
    ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
    if rho is not None:
        ctx.add_volume_force([0.0, 0.0, -0.01], density=rho, skip_below=1e-6)
    nfd = ctx.node_dofs()
    pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
    m = ctx.apply_dirichlet(pres)
    st = ctx.solve_pcg(tol, tol, 20000, matrix_free=mf, graph=False, two_level=two_level)
    u = ctx.solution()
    e, c, ee = ctx.energy(per_element=True)
    _, vm, mx, arg = ctx.stresses(False, True)
    # the free-function form: another field (global dof order on every rank), another material, ctx state untouched
    w = np.cos(np.arange(u.size) * 0.37)
    if rho is not None:
        sg2, vm2, mx2, arg2 = ctx.calculate_stresses(w, simp=(2.0, 0.25, 1e-6, 2.0, rho), want_sigma=True, want_vm=True)
    else:
        sg2, vm2, mx2, arg2 = ctx.calculate_stresses(w, lame=(3.0, 1.5), want_sigma=True, want_vm=True)
    assert np.array_equal(ctx.solution(), u)
    return dict(u=u, e=e, c=c, ee=ee, it=st["niter"], conv=st["converged"], brk=st["breakdown"], restarts=st["restarts"], m=m, f=ctx.rhs(),
                nfd=nfd, mx=mx, arg=arg, vm=vm, diag=ctx.diagonal(), sg2=sg2, vm2=vm2, mx2=mx2, arg2=arg2)


def _run_ranks(pkg, world, prob, mf, repeats=1, tol=1e-10, two_level=False):
    uid = pkg.Context.comm_unique_id()
    out = [None] * world
    err = [None] * world

    def worker(rank):
        try:
            ctx = pkg.Context(rank)
            ctx.comm_init(world, rank, uid)
            res = []
            for _ in range(repeats):
                res.append(_run(pkg, ctx, prob, True, mf, tol, two_level))
            res[-1]["part"] = ctx.partition()
            res[-1]["sizes"] = ctx.local_sizes()
            res[-1]["transport"] = ctx.comm_info()["transport"]
            x = np.random.default_rng(5).standard_normal(ctx.ndofs)
            res[-1]["spmv"] = ctx.spmv(x, matrix_free=mf)
            ctx.close()
            out[rank] = res
        except BaseException as ex:  # noqa: BLE001
            err[rank] = ex

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    for r, ex in enumerate(err):
        if ex is not None:
            raise AssertionError("rank %d failed: %r" % (r, ex))
    assert all(o is not None for o in out), "a rank thread did not finish"
    return out


@pytest.mark.parametrize("world,dims,simp,mf", [(2, (10, 4, 3), True, True), (4, (16, 4, 2), False, True), (4, (12, 5, 3), True, False),
                                                (8, (16, 4, 2), False, False)])
def test_partitioned_equals_single_ctx(emu, world, dims, simp, mf):
    pkg, lib = emu
    prob = _problem(pkg, dims, simp)
    single = pkg.Context(0)
    ref = _run(pkg, single, prob, False, mf)
    x = np.random.default_rng(5).standard_normal(single.ndofs)
    y_single = single.spmv(x, matrix_free=mf)
    single.close()
    assert ref["conv"] == 1
    ranks = _run_ranks(pkg, world, prob, mf, repeats=2 if world <= 4 else 1)      # re-set-up reproducibility at N = 2 and 4
    r0 = ranks[0][-1]
    counts = np.bincount(r0["part"], minlength=world)
    assert counts.max() - counts.min() <= 1, counts
    assert r0["transport"] == ("peer-memory" if world >= 8 else "nccl-allgather")      # the default transport by world size
    for rk in range(world):
        first, r = ranks[rk][0], ranks[rk][-1]
        # every rank returns the same global answer, re-setups are bit-reproducible, and it equals the unpartitioned solve
        assert np.array_equal(r["u"], r0["u"]) and r["it"] == r0["it"] and r["e"] == r0["e"]
        assert np.array_equal(first["u"], r["u"]) and first["it"] == r["it"], "second set-up on the same ctx gave a different solve"
        assert r["conv"] == 1 and r["brk"] == 0 and r["restarts"] == 0
        assert np.array_equal(r["nfd"], ref["nfd"])
        assert np.linalg.norm(r["u"] - ref["u"]) <= 1e-8 * np.linalg.norm(ref["u"])
        assert abs(r["e"] - ref["e"]) <= 1e-8 * abs(ref["e"]) and abs(r["c"] - ref["c"]) <= 1e-8 * abs(ref["c"])
        assert np.max(np.abs(r["ee"] - ref["ee"])) <= 1e-8 * np.max(np.abs(ref["ee"]))
        assert np.max(np.abs(r["f"] - ref["f"])) <= 1e-12 * np.max(np.abs(ref["f"]))
        assert np.max(np.abs(r["diag"] - ref["diag"]) / np.abs(ref["diag"])) <= 1e-12
        assert abs(r["m"] - ref["m"]) <= 1e-12 * ref["m"]
        assert abs(r["it"] - ref["it"]) <= max(5, ref["it"] // 50)
        assert r["arg"] == ref["arg"] and np.max(np.abs(r["vm"] - ref["vm"])) <= 1e-7 * ref["mx"]
        assert np.max(np.abs(r["spmv"] - y_single)) <= 1e-12 * np.max(np.abs(y_single))
        # per-cell outputs of the free-function stress recovery: every cell is computed by exactly one rank with the single-ctx arithmetic
        assert np.array_equal(r["sg2"], ref["sg2"]) and np.array_equal(r["vm2"], ref["vm2"]) and (r["mx2"], r["arg2"]) == (ref["mx2"], ref["arg2"])


@pytest.mark.parametrize("world,dims", [(4, (16, 4, 2)), (8, (16, 4, 2))])
def test_peer_memory_exchange_protocol(emu, world, dims, monkeypatch):
    """The fused peer-memory exchange kernel (mailboxes + flags + parity buffers + software grid barrier) under emulation:
    rank threads share one address space, the kernel's CTAs run co-resident, EMU_JITTER perturbs the interleaving.
    Results must be bit-identical to the NCCL transport (same arithmetic order) and re-set-ups must reproduce."""
    pkg, lib = emu
    monkeypatch.setenv("EMU_JITTER", "1")
    prob = _problem(pkg, dims, False)
    monkeypatch.setenv("TOE_DIST_XCHG", "allgather")
    nccl = _run_ranks(pkg, world, prob, False, repeats=1, tol=1e-9)
    monkeypatch.setenv("TOE_DIST_XCHG", "p2p")
    p2p = _run_ranks(pkg, world, prob, False, repeats=3, tol=1e-9)
    assert p2p[0][-1]["transport"] == "peer-memory" and nccl[0][-1]["transport"] == "nccl-allgather"
    for rk in range(world):
        for rep in range(3):
            r = p2p[rk][rep]
            assert r["conv"] == 1 and r["restarts"] == 0
            assert r["it"] == nccl[rk][0]["it"] and np.array_equal(r["u"], nccl[rk][0]["u"]), (rk, rep, r["it"], nccl[rk][0]["it"])


@pytest.mark.parametrize("world,dims,simp,mf", [(2, (10, 4, 3), False, True), (4, (10, 4, 2), True, False), (8, (16, 4, 2), False, False)])
def test_allgather_exchange_transport(emu, world, dims, simp, mf, monkeypatch):
    """The default transport — one ncclAllGather per exchange carries every rank's packed interface values and its partial scalars —
    against the opt-in send/recv + allreduce transport (TOE_DIST_XCHG=sendrecv): same arithmetic and summation order → bit-identical
    iterates, loads, diagonals and per-cell outputs; re-set-ups on the same ctx reproduce."""
    pkg, lib = emu
    prob = _problem(pkg, dims, simp)
    monkeypatch.setenv("TOE_DIST_XCHG", "sendrecv")
    ref = _run_ranks(pkg, world, prob, mf, repeats=1, tol=1e-9)
    monkeypatch.setenv("TOE_DIST_XCHG", "allgather")
    ag = _run_ranks(pkg, world, prob, mf, repeats=2, tol=1e-9)
    assert ag[0][-1]["transport"] == "nccl-allgather" and ref[0][-1]["transport"] == "nccl"
    for rk in range(world):
        a = ref[rk][0]
        for rep in range(2):
            r = ag[rk][rep]
            assert r["conv"] == 1 and r["restarts"] == 0 and r["it"] == a["it"], (rk, rep, r["it"], a["it"])
            for key in ("u", "f", "diag", "ee", "vm", "vm2"):
                assert np.array_equal(r[key], a[key]), (rk, rep, key)
            assert r["e"] == a["e"] and r["c"] == a["c"] and r["m"] == a["m"] and (r["mx"], r["arg"]) == (a["mx"], a["arg"])
        assert np.array_equal(ag[rk][-1]["spmv"], ref[rk][-1]["spmv"])


@pytest.mark.parametrize("world,dims,simp,mf", [(4, (12, 4, 2), True, False), (2, (8, 4, 2), False, True)])
def test_two_level_preconditioner_on_partitions(emu, world, dims, simp, mf, monkeypatch):
    """Jacobi + rigid-body coarse space on a partitioned ctx: per-box sums over OWNED nodes + allreduce, coarse operator probed
    through the interface-summed operator; must reproduce the single-ctx two-level solve (same boxes, same iteration count ±1)."""
    pkg, lib = emu
    monkeypatch.setenv("TOE_TL_BOXES", "4,2,1")
    prob = _problem(pkg, dims, simp)
    single = pkg.Context(0)
    ref_j = _run(pkg, single, prob, False, mf)
    ref = _run(pkg, single, prob, False, mf, two_level=True)
    single.close()
    assert ref["conv"] == 1 and ref["it"] < ref_j["it"]
    ranks = _run_ranks(pkg, world, prob, mf, repeats=2, two_level=True)
    for rk in range(world):
        first, r = ranks[rk][0], ranks[rk][-1]
        assert r["conv"] == 1 and r["brk"] == 0 and r["restarts"] == 0
        assert np.array_equal(r["u"], ranks[0][-1]["u"]) and np.array_equal(first["u"], r["u"]) and first["it"] == r["it"]
        assert abs(r["it"] - ref["it"]) <= 1, (r["it"], ref["it"])
        assert np.linalg.norm(r["u"] - ref["u"]) <= 1e-8 * np.linalg.norm(ref["u"])
        assert np.linalg.norm(r["u"] - ref_j["u"]) <= 1e-8 * np.linalg.norm(ref_j["u"])
        assert abs(r["e"] - ref["e"]) <= 1e-8 * abs(ref["e"])


def test_partitioned_results_do_not_depend_on_block_order(emu):
    """the partitioned path (single-reduction CG, pack / unpack, partition set-up kernels) under first-to-last, last-to-first and
    shuffled block execution: bit-identical solves — no kernel leans on the order in which the blocks of a launch happen to run"""
    pkg, lib = emu
    prob = _problem(pkg, (10, 4, 2), True)
    try:
        lib.emu_set_block_order(0)
        ref = _run_ranks(pkg, 2, prob, False, repeats=1, tol=1e-9)
        for mode in (2, 1)[:1 if os.environ.get("TOE_EMU_FULL") != "1" else 2]:      # shuffled (and, on request, last-to-first)
            lib.emu_set_block_order(mode)
            got = _run_ranks(pkg, 2, prob, False, repeats=1, tol=1e-9)
            for rk in range(2):
                a, b = ref[rk][0], got[rk][0]
                assert a["it"] == b["it"] and b["conv"] == 1 and b["restarts"] == 0, (mode, rk, a["it"], b["it"])
                for key in ("u", "f", "diag", "ee", "vm", "nfd"):
                    assert np.array_equal(a[key], b[key]), (mode, rk, key)
                assert np.array_equal(ref[rk][-1]["part"], got[rk][-1]["part"])
    finally:
        lib.emu_set_block_order(0)


def test_cg_scalar_trace_is_consistent(emu, monkeypatch):
    """toe_debug_cg_trace (TOE_CG_TRACE=1, the diagnostic that located the round-1 SpMV pipeline bug): per iteration the summed γ and δ are
    identical on every rank and equal the rank-ordered sum of the ranks' partials, and √γ is the residual history."""
    pkg, lib = emu
    monkeypatch.setenv("TOE_CG_TRACE", "1")
    prob = _problem(pkg, (10, 4, 2), False)
    world = 2
    uid = pkg.Context.comm_unique_id()
    out, err = [None] * world, [None] * world

    def worker(rank):
        try:
            ctx = pkg.Context(rank)
            ctx.comm_init(world, rank, uid)
            pts, cells, rho, fixed, load = prob
            lam, mu = pkg.create_material_model(1.0, 0.3)
            ctx.set_mesh(pts, cells, distributed=True); ctx.build_dofs(); ctx.build_pattern()
            ctx.assemble_lame(lam, mu); ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
            nfd = ctx.node_dofs()
            ctx.apply_dirichlet(np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1)))
            st = ctx.solve_pcg(1e-9, 1e-9, 5000, graph=False, history=True)
            out[rank] = (st, ctx.cg_trace(st["niter"] + 1))
            ctx.close()
        except BaseException as ex:  # noqa: BLE001
            err[rank] = ex

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    assert err == [None, None], err
    (st0, t0), (st1, t1) = out
    n = st0["niter"]
    assert st0["converged"] == 1 and st1["niter"] == n and t0.shape == (n + 1, 4)
    assert np.array_equal(t0[:, :2], t1[:, :2])                                   # summed scalars: same bits on both ranks
    assert np.array_equal(t0[1:, 0], t0[1:, 2] + t1[1:, 2])                       # γ_j = rank 0's partial + rank 1's, in rank order
    assert np.array_equal(t0[1:, 1], t0[1:, 3] + t1[1:, 3])                       # δ_j likewise
    assert np.array_equal(np.sqrt(t0[:, 0]), st0["residuals"])                   # √γ_j is the history Krylov.jl would report
