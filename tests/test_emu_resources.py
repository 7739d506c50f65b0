"""Resource lifetime of a toe_ctx, checked on the EMULATED build (tests/cuda_emu, test infrastructure — see tests/emu_support.py):
every device allocation the library makes must be gone after toe_destroy — after a full pass through every feature, after calls
that returned errors, and after re-set-ups that grow and shrink the problem — and no allocation may have been written out of bounds."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_support  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    pkg, lib = emu_support.load_emu()
    with emu_support.emulated(pkg, lib):
        yield pkg, lib


def _full_pass(pkg, ctx, dims, hexm, two_level=True):
    A = pkg._lib
    pts, cells = pkg.meshgen.cantilever(*dims, hex=hexm)
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
    rho = pkg.meshgen.simp_like_density(cells.shape[0])
    lam, mu = pkg.create_material_model(1.0, 0.3)
    fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
    nfd = ctx.node_dofs()
    pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
    for variant in (A.ASM_GATHER, A.ASM_ATOMIC, A.ASM_ROWS):
        ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho, variant)
    ctx.ke_batch(1, 3); ctx.pattern(); ctx.values(); ctx.cell_dofs(1, 2)
    ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
    ctx.add_volume_force([0.0, 0.0, -0.01], density=rho, skip_below=1e-6)
    facets = ctx.boundary_facets(ctx.select_nodes_by_plane([60.0, 0.0, 0.0], [1.0, 0.0, 0.0], 1e-6))
    ctx.boundary_area(facets); ctx.facet_quadrature(facets)
    ctx.add_surface_traction(facets, traction_uniform=[0.0, 0.0, -0.5])
    ctx.select_nodes_by_circle([0.0, 10.0, 2.0], [1.0, 0.0, 0.0], 3.0, 1e-6); ctx.surface_nodes()
    ctx.apply_dirichlet(pres)
    for mf, graph, tl in ((False, True, False), (False, False, True), (True, True, False), (True, True, True)):
        if tl and not two_level:
            continue
        st = ctx.solve_pcg(1e-8, 1e-8, 20000, matrix_free=mf, graph=graph, history=True, two_level=tl)
        assert st["converged"] == 1
    ctx.energy(per_element=True); ctx.energy_assembled(); ctx.stresses(True, True)
    ctx.calculate_stresses(ctx.solution(), lame=(lam, mu), want_sigma=True, want_vm=True)
    ctx.calculate_stresses(None, simp=(1.0, 0.3, 1e-8, 3.0, rho), want_vm=True)
    ctx.spmv(np.ones(ctx.ndofs)); ctx.spmv(np.ones(ctx.ndofs), matrix_free=True); ctx.time_spmv(reps=2); ctx.diagonal(); ctx.rhs()
    os.environ["TOE_EBE_PIPE"] = "1"
    try:
        ctx.spmv(np.ones(ctx.ndofs), matrix_free=True)
    finally:
        del os.environ["TOE_EBE_PIPE"]


def test_no_device_allocation_survives_destroy(emu):
    pkg, lib = emu
    base = lib.emu_live_allocations()
    ctx = pkg.Context(0)
    _full_pass(pkg, ctx, (6, 2, 2), False)
    _full_pass(pkg, ctx, (8, 3, 2), False, two_level=False)   # larger mesh on the same ctx: buffers grow
    _full_pass(pkg, ctx, (3, 2, 2), True, two_level=False)     # Hex8, smaller: buffers are reused
    assert lib.emu_live_allocations() > base
    assert lib.emu_check_all_guards() == 0, "a device allocation was written out of bounds"
    ctx.close()
    assert lib.emu_live_allocations() == base, "%d device allocations leaked" % (lib.emu_live_allocations() - base)


def test_error_returns_do_not_leak(emu):
    """every call below fails (bad arguments / wrong state / inverted cell); temporaries of the failed call and everything the ctx
    owns must still be released"""
    pkg, lib = emu
    base = lib.emu_live_allocations()
    ctx = pkg.Context(0)
    pts, cells = pkg.meshgen.cantilever(6, 2, 2)
    E = pkg.TopOptError
    with pytest.raises(E):
        ctx.build_pattern()                                          # no mesh yet
    with pytest.raises(E):
        ctx.set_mesh(pts, np.full_like(cells, pts.shape[0] + 7))     # node ids out of range
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
    with pytest.raises(E):
        ctx.solve_pcg(1e-8, 1e-8, 100)                               # K not assembled
    bad = cells.copy(); bad[3, [0, 1]] = bad[3, [1, 0]]
    ctx.set_mesh(pts, bad); ctx.build_dofs(); ctx.build_pattern()
    for variant in (pkg._lib.ASM_GATHER, pkg._lib.ASM_ROWS, pkg._lib.ASM_ATOMIC):
        with pytest.raises(E):
            ctx.assemble_lame(0.5, 0.4, variant)                     # det(J) <= 0
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
    ctx.assemble_lame(0.5, 0.4)
    with pytest.raises(E):
        ctx.add_nodal_force(np.array([], dtype=np.int64), [0, 0, -1.0])
    with pytest.raises(E):
        ctx.add_nodal_force(np.array([10 ** 9], dtype=np.int64), [0, 0, -1.0])
    with pytest.raises(E):
        ctx.apply_dirichlet(np.array([ctx.ndofs + 1], dtype=np.int64))
    with pytest.raises(E):
        ctx.ke_batch(0, 4)
    with pytest.raises(E):
        ctx.calculate_stresses(np.zeros(ctx.ndofs), simp=(1.0, 0.3, 1e-8, 3.0, np.ones(3)))
    with pytest.raises(E):
        ctx.boundary_facets(np.array([-5], dtype=np.int64))
    assert lib.emu_check_all_guards() == 0
    ctx.close()
    assert lib.emu_live_allocations() == base, "%d device allocations leaked" % (lib.emu_live_allocations() - base)


def test_many_contexts_come_and_go(emu):
    pkg, lib = emu
    base = lib.emu_live_allocations()
    pts, cells = pkg.meshgen.cantilever(4, 2, 1)
    for _ in range(5):
        ctx = pkg.Context(0)
        ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
        ctx.assemble_lame(0.5, 0.4)
        ctx.close()
    assert lib.emu_live_allocations() == base


def test_results_do_not_depend_on_block_order(emu):
    """CUDA promises no order among the blocks of a launch.  The emulated build runs them first-to-last, last-to-first and in a
    shuffled order: everything the library calls deterministic (K, loads, PCG iterates of both operator forms and of the two-level
    preconditioner, energies, stresses) must come out bit-identical — a kernel in which a block leans on another block of the same
    launch having run already (or not yet) fails here."""
    pkg, lib = emu
    A = pkg._lib
    pts, cells = pkg.meshgen.cantilever(7, 3, 2)
    rho = pkg.meshgen.simp_like_density(cells.shape[0])
    fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)

    def run():
        ctx = pkg.Context(0)
        out = {}
        ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
        nfd = ctx.node_dofs()
        pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
        out["pattern"] = np.concatenate(ctx.pattern())
        for name, variant in (("gather", A.ASM_GATHER), ("rows", A.ASM_ROWS)):
            ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho, variant)
            out["K_" + name] = ctx.values()
        ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho)
        ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
        ctx.add_volume_force([0.0, 0.0, -0.01], density=rho, skip_below=1e-6)
        facets = ctx.boundary_facets(ctx.surface_nodes())                      # the whole skin: every surface node collects several facets
        ctx.add_surface_traction(facets, traction_uniform=[0.3, -0.2, -0.5])
        out["facets"] = np.asarray(facets)
        out["m"] = ctx.apply_dirichlet(pres)
        out["f"] = ctx.rhs()
        for key, kw in (("asm", {}), ("mf", {"matrix_free": True}), ("tl", {"two_level": True})):
            st = ctx.solve_pcg(1e-9, 1e-9, 20000, **kw)
            assert st["converged"] == 1
            out["u_" + key] = ctx.solution(); out["it_" + key] = st["niter"]
        e, c, ee = ctx.energy(per_element=True)
        sig, vm, mx, arg = ctx.stresses(True, True)
        out.update(e=e, c=c, ee=ee, sig=sig, vm=vm, mx=mx, arg=arg)
        ctx.close()
        return out

    lib.emu_set_block_order(0)
    try:
        ref = run()
        for mode in (2, 1)[:1 if os.environ.get("TOE_EMU_FULL") != "1" else 2]:      # shuffled (and, on request, last-to-first)
            lib.emu_set_block_order(mode)
            r = run()
            for k, v in ref.items():
                same = np.array_equal(v, r[k]) if isinstance(v, np.ndarray) else v == r[k]
                assert same, "block order %d changes %s" % (mode, k)
    finally:
        lib.emu_set_block_order(0)


def test_maximum_sizes_are_refused_before_anything_is_read(emu):
    """device indices are 32-bit: meshes beyond that (nn > (2^31-1)/3 nodes, ne·npc > 2^31-1 connectivity entries, more cells than the
    packed (cell, a, b) contribution lists hold) must come back as an error — checked with the size arguments alone, before the
    library touches the (tiny) arrays behind the pointers — and leave the ctx usable"""
    import ctypes as C
    pkg, lib = emu
    ctx = pkg.Context(0)
    pts, cells = pkg.meshgen.cantilever(2, 1, 1)
    xyz = np.ascontiguousarray(pts, dtype=np.float64); conn = np.ascontiguousarray(cells, dtype=np.int64)
    dp = xyz.ctypes.data_as(C.POINTER(C.c_double)); ip = conn.ctypes.data_as(C.POINTER(C.c_int64))
    for nn, ne, npc in ((2 ** 31, 1, 4), ((2 ** 31 - 1) // 3 + 1, 1, 4), (8, 2 ** 29, 4), (8, 2 ** 28, 8), (0, 1, 4), (8, 0, 4), (8, 1, 5)):
        st = lib.toe_set_mesh(ctx.h, nn, dp, ne, npc, ip)
        assert st < 0, (nn, ne, npc)
        msg = lib.toe_last_error(ctx.h).decode()
        assert ("too large" in msg) or ("empty mesh" in msg) or ("unsupported cell type" in msg), msg
    st = lib.toe_set_mesh_distributed(ctx.h, 2 ** 31, dp, 1, 4, ip)
    assert st < 0                                                   # no communicator, and too large anyway
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()       # the ctx survived all of it
    ctx.assemble_lame(0.5, 0.4)
    assert np.all(ctx.diagonal() > 0)
    ctx.close()
