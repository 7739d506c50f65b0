"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against the CPU oracle
on the same inputs and against the committed golden fixtures.

Bars (BASELINE.json north_star): DOF maps and sparsity pattern bit-exact; Ke entries 1e-12 relative (fp64);
displacement field, per-element energies and compliance 1e-8 relative.
"""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_KE = 1e-12
TOL_U = 1e-8


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _setup(ctx, pts, cells):
    ctx.set_mesh(pts, cells)
    ctx.build_dofs()
    ctx.build_pattern()


def _row_scale(prob):
    K = prob.K().tocsr()
    rowmax = np.maximum.reduceat(np.abs(K.data), K.indptr[:-1])
    K2 = prob.K()
    return rowmax[K2.indices]        # per stored entry (CSC order): scale of its row


# ----------------------------------------------------------------------------------------------------------
# setup_problem: DOF map + pattern, bit-exact
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["c1", "c2", "syn"])
def test_dofs_and_pattern_bit_exact(ctx, pkg, fo, golden_c1, golden_c2, golden_syn, case):
    if case == "c1":
        pts, cells, g = golden_c1["points"], golden_c1["cells"].astype(np.int64), golden_c1
    elif case == "c2":
        pts, cells, g = golden_c2["points"], golden_c2["cells"].astype(np.int64), golden_c2
    else:
        pts, cells = pkg.meshgen.cantilever(24, 8, 4)
        g = {k.replace("24x8x4_", ""): v for k, v in golden_syn.items() if k.startswith("24x8x4_")}
    _setup(ctx, pts, cells)
    prob = fo.setup_problem(pts, cells)
    assert ctx.ndofs == prob.ndofs == int(g["ndofs"])
    assert ctx.nnz == prob.nnz == int(g["nnz"])
    assert np.array_equal(ctx.node_dofs(), prob.node_first_dof)
    assert np.array_equal(ctx.cell_dofs(), prob.cell_dofs)
    colptr, rowval = ctx.pattern()
    assert np.array_equal(colptr, prob.colptr)
    assert np.array_equal(rowval, prob.rowval)
    assert sha(colptr) == str(g["colptr_sha"]) and sha(rowval) == str(g["rowval_sha"])


def test_dofs_permuted_cells_and_unreferenced_nodes(ctx, pkg, fo):
    """first-touch numbering must follow the cell walk, not the node ids; nodes in no cell get no DOFs."""
    pts, cells = pkg.meshgen.cantilever(6, 3, 2)
    rng = np.random.default_rng(7)
    cells = cells[rng.permutation(cells.shape[0])]
    pts = np.vstack([pts, [[100.0, 100.0, 100.0], [101.0, 100.0, 100.0]]])     # two orphan nodes
    _setup(ctx, pts, cells)
    prob = fo.setup_problem(pts, cells)
    nfd = ctx.node_dofs()
    assert np.array_equal(nfd, prob.node_first_dof) and nfd[-1] == 0 and nfd[-2] == 0
    lit = fo.first_touch_dofs_literal(cells, pts.shape[0])
    assert np.array_equal(nfd, lit[0]) and np.array_equal(ctx.cell_dofs(), lit[1])
    colptr, rowval = ctx.pattern()
    assert np.array_equal(colptr, prob.colptr) and np.array_equal(rowval, prob.rowval)


# ----------------------------------------------------------------------------------------------------------
# element stiffness
# ----------------------------------------------------------------------------------------------------------
def test_ke_tet_fixture(ctx, pkg, fo, golden_c1):
    pts, cells = golden_c1["points"], golden_c1["cells"].astype(np.int64)
    _setup(ctx, pts, cells)
    lam, mu = fo.create_material_model(1.0, 0.3)
    ctx.set_material_lame(lam, mu)
    ke = ctx.ke_batch(1, ctx.ne)
    ref = fo.element_stiffness(pts, cells, lam, mu)
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(ke - ref) / scale) <= TOL_KE
    ids = golden_c1["ke_sample_ids"]
    assert np.max(np.abs(ke[ids - 1] - golden_c1["ke_sample"]) / scale[ids - 1]) <= TOL_KE
    assert np.array_equal(ke, ke.transpose(0, 2, 1)), "closed-form Ke is exactly symmetric"


def test_ke_hex_simp_fixture(ctx, pkg, fo, golden_c2):
    pts, cells, rho = golden_c2["points"], golden_c2["cells"].astype(np.int64), golden_c2["density"]
    _setup(ctx, pts, cells)
    ctx.set_material_simp(1.0, 0.3, 1e-8, 3.0, rho)
    ke = ctx.ke_batch(1, ctx.ne)
    lam, mu = fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)(rho)
    ref = fo.element_stiffness(pts, cells, lam, mu)
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(ke - ref) / scale) <= TOL_KE
    ids = golden_c2["ke_sample_ids"]
    assert np.max(np.abs(ke[ids - 1] - golden_c2["ke_sample"]) / scale[ids - 1]) <= TOL_KE


def test_ke_partial_range_and_errors(ctx, pkg, fo):
    pts, cells = pkg.meshgen.cantilever(4, 2, 2)
    _setup(ctx, pts, cells)
    ctx.set_material_lame(0.5, 0.4)
    a = ctx.ke_batch(1, ctx.ne)
    b = ctx.ke_batch(17, 5)
    assert np.array_equal(a[16:21], b)
    with pytest.raises(pkg.TopOptError):
        ctx.ke_batch(ctx.ne, 2)
    # inverted cell → det(J) <= 0 is an error, like Ferrite's reinit!
    bad = cells.copy(); bad[3, [0, 1]] = bad[3, [1, 0]]
    _setup(ctx, pts, bad)
    with pytest.raises(pkg.TopOptError, match="det"):
        ctx.assemble_lame(0.5, 0.4)


# ----------------------------------------------------------------------------------------------------------
# assembly, loads, Dirichlet
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["gather", "atomic"])
def test_assembled_K_tet(ctx, pkg, fo, golden_c1, variant):
    pts, cells = golden_c1["points"], golden_c1["cells"].astype(np.int64)
    _setup(ctx, pts, cells)
    lam, mu = fo.create_material_model(1.0, 0.3)
    ctx.assemble_lame(lam, mu, pkg._lib.ASM_GATHER if variant == "gather" else pkg._lib.ASM_ATOMIC)
    prob = fo.setup_problem(pts, cells)
    fo.assemble_stiffness_matrix(prob, lam, mu)
    nz = ctx.values()
    assert np.max(np.abs(nz - prob.nzval) / _row_scale(prob)) <= TOL_KE
    assert abs(nz.sum() - float(golden_c1["K_unconstrained_sum"])) <= 1e-9 * float(golden_c1["K_unconstrained_abs_sum"])
    assert np.all(ctx.rhs() == 0.0)
    d = ctx.diagonal()
    assert np.max(np.abs(d - prob.K().diagonal()) / np.abs(d)) <= TOL_KE


@pytest.mark.parametrize("variant", ["gather", "atomic"])
def test_assembled_K_hex_simp(ctx, pkg, fo, golden_c2, variant):
    pts, cells, rho = golden_c2["points"], golden_c2["cells"].astype(np.int64), golden_c2["density"]
    _setup(ctx, pts, cells)
    ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho, pkg._lib.ASM_GATHER if variant == "gather" else pkg._lib.ASM_ATOMIC)
    prob = fo.setup_problem(pts, cells)
    fo.assemble_stiffness_matrix_simp(prob, fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0), rho)
    nz = ctx.values()
    assert np.max(np.abs(nz - prob.nzval) / _row_scale(prob)) <= TOL_KE


def test_gather_assembly_is_deterministic_and_symmetric(ctx, pkg):
    pts, cells = pkg.meshgen.cantilever(12, 4, 2)
    _setup(ctx, pts, cells)
    ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, pkg.meshgen.simp_like_density(cells.shape[0]))
    a = ctx.values()
    ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, pkg.meshgen.simp_like_density(cells.shape[0]))
    assert np.array_equal(a, ctx.values())
    import scipy.sparse as sp
    colptr, rowval = ctx.pattern()
    K = sp.csc_matrix((a, rowval - 1, colptr - 1), shape=(ctx.ndofs, ctx.ndofs))
    assert (K - K.T).nnz == 0 or abs(K - K.T).max() == 0.0


def test_per_cell_lame_matches_simp(ctx, pkg, fo):
    pts, cells = pkg.meshgen.cantilever(6, 3, 2)
    rho = pkg.meshgen.simp_like_density(cells.shape[0])
    _setup(ctx, pts, cells)
    ctx.assemble_simp(2.0, 0.25, 1e-6, 1.0, rho)
    a = ctx.values()
    lam, mu = fo.create_simp_material_model(2.0, 0.25)(rho)       # code defaults Emin=1e-6, p=1.0
    ctx.assemble_lame_per_cell(lam, mu)
    b = ctx.values()
    assert np.max(np.abs(a - b)) <= 1e-14 * np.abs(a).max()


def test_loads(ctx, pkg, fo, golden_c2):
    pts, cells, rho = golden_c2["points"], golden_c2["cells"].astype(np.int64), golden_c2["density"]
    _setup(ctx, pts, cells)
    ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho)
    prob = fo.setup_problem(pts, cells)
    # nodal force (apply_force!)
    load = golden_c2["load_nodes"]
    ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
    fo.apply_force(prob, load, [0.0, 0.0, -1.0])
    assert np.max(np.abs(ctx.rhs() - prob.f)) <= 1e-15
    assert np.allclose(ctx.rhs(), golden_c2["f_loaded"], rtol=0, atol=1e-15)
    # variable-density volume force on top (loads accumulate, like the reference's += into f)
    tot = ctx.add_volume_force([0.0, 0.0, -1.0], density=rho, skip_below=1e-6)
    tot_ref = fo.apply_variable_density_volume_force(prob, [0.0, 0.0, -1.0], rho)
    assert np.max(np.abs(ctx.rhs() - prob.f)) <= 1e-12 * np.abs(prob.f).max()
    assert abs(tot[2] - tot_ref[2]) <= 1e-10 * abs(tot_ref[2]) and abs(tot[2] + 1923.3236661882) < 1e-6
    # uniform volume force / gravity recipe: b is divided by density and multiplied back
    ctx.set_rhs(np.zeros(ctx.ndofs)); prob.f[:] = 0
    ctx.add_volume_force([0.0, 2.0, -3.0], rho_uniform=7850.0)
    fo.apply_volume_force(prob, [0.0, 2.0, -3.0], 7850.0)
    assert np.max(np.abs(ctx.rhs() - prob.f)) <= 1e-12 * np.abs(prob.f).max()
    with pytest.raises(pkg.TopOptError, match="No nodes"):
        ctx.add_nodal_force(np.array([], dtype=np.int64), [0.0, 0.0, -1.0])


def test_volume_force_tet(ctx, pkg, fo, golden_c1):
    pts, cells = golden_c1["points"], golden_c1["cells"].astype(np.int64)
    _setup(ctx, pts, cells)
    ctx.assemble_lame(1.0, 1.0)
    prob = fo.setup_problem(pts, cells)
    tot = ctx.add_volume_force([1.0, 0.0, -9.81], rho_uniform=1.0)
    tot_ref, vol = fo.apply_volume_force(prob, [1.0, 0.0, -9.81], 1.0)
    assert np.max(np.abs(ctx.rhs() - prob.f)) <= 1e-12 * np.abs(prob.f).max()
    assert abs(vol - 1928.3685972) < 1e-6 and np.allclose(tot, tot_ref, rtol=1e-11)


def test_dirichlet_ferrite_semantics(ctx, pkg, fo, golden_c1):
    pts, cells = golden_c1["points"], golden_c1["cells"].astype(np.int64)
    _setup(ctx, pts, cells)
    lam, mu = fo.create_material_model(1.0, 0.3)
    ctx.assemble_lame(lam, mu)
    prob = fo.setup_problem(pts, cells)
    fo.assemble_stiffness_matrix(prob, lam, mu)
    ctx.add_nodal_force(golden_c1["load_nodes"], [0.0, 0.0, -1.0]); fo.apply_force(prob, golden_c1["load_nodes"], [0.0, 0.0, -1.0])
    pres = golden_c1["prescribed"]
    # two handlers applied in sequence: m is recomputed on the already-modified K
    m1 = ctx.apply_dirichlet(pres[:60]); m1r = fo.apply_dirichlet(prob, pres[:60])
    m2 = ctx.apply_dirichlet(pres[60:]); m2r = fo.apply_dirichlet(prob, pres[60:])
    assert abs(m1 - m1r) <= 1e-13 * m1r and abs(m2 - m2r) <= 1e-13 * m2r
    assert abs(m1 - float(golden_c1["mean_diag"])) <= 1e-12 * m1
    nz = ctx.values()
    assert np.max(np.abs(nz - prob.nzval)) <= TOL_KE * np.abs(prob.nzval).max()
    # zeroed entries stay in the pattern as explicit zeros; the prescribed diagonal carries its handler's m
    col_of = np.repeat(np.arange(prob.ndofs), np.diff(prob.colptr)); row_of = prob.rowval - 1
    flag = np.zeros(prob.ndofs, dtype=bool); flag[pres - 1] = True
    hit = flag[col_of] | flag[row_of]
    diag = col_of == row_of
    assert np.all(nz[hit & ~diag] == 0.0) and np.allclose(nz[hit], prob.nzval[hit], rtol=1e-13, atol=0)
    assert set(np.unique(nz[hit & diag])) == {m1, m2} and abs(m1 - m2) > 0
    assert np.max(np.abs(ctx.rhs() - prob.f)) <= 1e-15


# ----------------------------------------------------------------------------------------------------------
# operator, solve, energy
# ----------------------------------------------------------------------------------------------------------
def test_spmv_assembled_and_matrix_free(ctx, pkg, fo, golden_c1, golden_c2):
    rng = np.random.default_rng(3)
    for g, simp in ((golden_c1, False), (golden_c2, True)):
        pts, cells = g["points"], g["cells"].astype(np.int64)
        _setup(ctx, pts, cells)
        prob = fo.setup_problem(pts, cells)
        if simp:
            ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, g["density"])
            fo.assemble_stiffness_matrix_simp(prob, fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0), g["density"])
        else:
            lam, mu = fo.create_material_model(1.0, 0.3)
            ctx.assemble_lame(lam, mu); fo.assemble_stiffness_matrix(prob, lam, mu)
        x = rng.standard_normal(ctx.ndofs)
        y_ref = prob.K() @ x
        s = np.abs(prob.K()) @ np.abs(x)
        assert np.max(np.abs(ctx.spmv(x) - y_ref) / s) <= 1e-13
        assert np.max(np.abs(ctx.spmv(x, matrix_free=True) - y_ref) / s) <= 1e-12
        # constrained operator: both forms agree with the oracle's constrained K
        pres = g["prescribed"]
        ctx.apply_dirichlet(pres); fo.apply_dirichlet(prob, pres)
        y_ref = prob.K() @ x
        s = np.abs(prob.K()) @ np.abs(x)
        assert np.max(np.abs(ctx.spmv(x) - y_ref) / s) <= 1e-13
        assert np.max(np.abs(ctx.spmv(x, matrix_free=True) - y_ref) / s) <= 1e-12


def _solve_case(ctx, pkg, g, *, simp, load, matrix_free, tol=1e-10, itmax=200000):
    pts, cells = g["points"], g["cells"].astype(np.int64)
    _setup(ctx, pts, cells)
    if simp:
        (ctx.set_material_simp if matrix_free else ctx.assemble_simp)(1.0, 0.3, 1e-8, 3.0, g["density"])
    else:
        lam, mu = 1.0 * 0.3 / (1.3 * 0.4), 1.0 / 2.6
        (ctx.set_material_lame if matrix_free else ctx.assemble_lame)(lam, mu)
    if load == "tip":
        ctx.add_nodal_force(g["load_nodes"], [0.0, 0.0, -1.0])
    else:
        ctx.add_volume_force([0.0, 0.0, -1.0], density=g["density"], skip_below=1e-6)
    ctx.apply_dirichlet(g["prescribed"])
    st = ctx.solve_pcg(tol, tol, itmax, matrix_free=matrix_free, history=True)
    return st


@pytest.mark.parametrize("matrix_free", [False, True])
def test_solve_c1_tet_beam(ctx, pkg, golden_c1, matrix_free):
    """test/runtests.jl:21-49 recipe; parity target = the oracle's direct solve."""
    g = golden_c1
    st = _solve_case(ctx, pkg, g, simp=False, load="tip", matrix_free=matrix_free)
    assert st["converged"] == 1 and st["breakdown"] == 0
    u = ctx.solution()
    assert rel(u, g["u"]) <= TOL_U
    e, c, ee = ctx.energy(per_element=True)
    assert abs(e - float(g["energy"])) <= TOL_U * float(g["energy"])
    assert abs(c - float(g["compliance"])) <= TOL_U * float(g["compliance"])
    assert np.max(np.abs(ee - g["elem_energy"])) <= TOL_U * np.abs(g["elem_energy"]).max()
    assert abs(ee.sum() - e) <= 1e-12 * e
    if not matrix_free:
        assert abs(ctx.energy_assembled() - float(g["energy"])) <= TOL_U * float(g["energy"])
    assert np.all(u[g["prescribed"] - 1] == 0.0)
    assert st["rel_res_l2"] < 1e-7
    assert len(st["residuals"]) == st["niter"] + 1 and st["residuals"][0] == st["res0_M"]


def test_pcg_krylov_semantics_and_iteration_count(ctx, pkg, golden_c1):
    """atol = rtol = 1e-8 on the M-norm (RobustSolver.jl:294-305): iteration count must track the oracle's PCG."""
    g = golden_c1
    st = _solve_case(ctx, pkg, g, simp=False, load="tip", matrix_free=False, tol=1e-8)
    ref_it = int(g["pcg_niter"])
    assert st["converged"] == 1
    assert abs(st["niter"] - ref_it) <= max(5, ref_it // 50), (st["niter"], ref_it)
    assert st["res_M"] <= 1e-8 + 1e-8 * st["res0_M"]
    assert rel(ctx.solution(), g["u"]) <= 1e-8
    # itmax is honoured exactly; not converged is reported, not hidden
    st2 = _solve_case(ctx, pkg, g, simp=False, load="tip", matrix_free=False, tol=1e-8, itmax=137)
    assert st2["niter"] == 137 and st2["converged"] == 0
    # no-graph path gives the same iterates
    pts = g["points"]
    st3 = ctx.solve_pcg(1e-8, 1e-8, 200000, graph=False)
    assert st3["niter"] == st["niter"]


def test_l2_norm_criterion_and_true_residual(ctx, pkg, fo, golden_syn):
    """TOE_PCG_L2_NORM (SURVEY §8(b) norm_kind): stop on ||r||_2 <= atol + rtol*||r0||_2 instead of Krylov.jl's M-norm rule, on a
    synthetic cantilever (24x8x4 cubes on the GPU), assembled and matrix-free; checked against an l2-stopped PCG of the oracle's K (same recurrence in
    numpy) and against the true residual the library recomputes after the solve (toe_pcg_stats.true_res)."""
    mg = pkg.meshgen
    emulated = hasattr(ctx.lib, "emu_check_all_guards")              # the host-side emulation exports this symbol, the product library does not
    pts, cells = mg.cantilever(*((12, 4, 2) if emulated else (24, 8, 4)))          # emulated PCG solves are slow: smaller box there
    _setup(ctx, pts, cells)
    lam, mu = fo.create_material_model(1.0, 0.3)
    prob = fo.setup_problem(pts, cells)
    fo.assemble_stiffness_matrix(prob, lam, mu)
    load = mg.nodes_at_plane(pts, 0, 60.0); fixed = mg.nodes_at_plane(pts, 0, 0.0)
    fo.apply_force(prob, load, [0.0, 0.0, -1.0])
    pres = fo.fixed_boundary_dofs(prob, fixed)
    fo.apply_dirichlet(prob, pres)
    K = prob.K().tocsr(); f = prob.f.copy()
    # numpy PCG with the l2 rule (Jacobi, x0 = 0)
    Minv = 1.0 / np.where(np.abs(K.diagonal()) < 1e-12, 1.0, K.diagonal())
    x = np.zeros_like(f); r = f.copy(); z = Minv * r; p = z.copy(); gamma = r @ z
    eps = 1e-9 + 1e-9 * np.linalg.norm(r); it_ref = 0
    while np.linalg.norm(r) > eps and it_ref < 100000:
        Ap = K @ p; alpha = gamma / (p @ Ap); x += alpha * p; r -= alpha * Ap; z = Minv * r; g2 = r @ z; p = z + (g2 / gamma) * p; gamma = g2; it_ref += 1
    for mf in (False, True):
        (ctx.set_material_lame if mf else ctx.assemble_lame)(lam, mu)
        ctx.add_nodal_force(load, [0.0, 0.0, -1.0]); ctx.apply_dirichlet(pres)
        st_m = ctx.solve_pcg(1e-9, 1e-9, 100000, matrix_free=mf)
        st = ctx.solve_pcg(1e-9, 1e-9, 100000, matrix_free=mf, l2_norm=True, history=True)
        assert st["converged"] == 1 and st_m["converged"] == 1
        assert abs(st["niter"] - it_ref) <= max(5, it_ref // 50), (mf, st["niter"], it_ref)          # l2 rule: the oracle's count
        assert st["niter"] != st_m["niter"]                                                            # and not the M-norm one
        nf = np.linalg.norm(f)
        assert abs(st["res0_M"] - nf) <= 1e-12 * nf and st["res_M"] <= 1e-9 + 1e-9 * nf                # history / stats hold l2 norms
        assert st["residuals"][0] == st["res0_M"] and st["residuals"][-1] == st["res_M"]
        # the recomputed true residual agrees with the recurrence to rounding, in the norm of the test
        assert st["true_res"] <= 2.0 * (1e-9 + 1e-9 * nf) and abs(st["rel_res_l2"] - st["true_res"] / nf) <= 1e-12
        assert st_m["true_res"] <= 2.0 * (1e-9 + 1e-9 * st_m["res0_M"])                                # M-norm run: true_res in the M-norm
        assert rel(ctx.solution(), x) <= 1e-7
    with pytest.raises(pkg.TopOptError):
        ctx.solve_pcg(1e-9, 1e-9, 100, two_level=True, l2_norm=True)                                   # Jacobi only


@pytest.mark.parametrize("load", ["tip", "volume"])
@pytest.mark.parametrize("matrix_free", [False, True])
def test_solve_c2_hex_simp(ctx, pkg, golden_c2, load, matrix_free):
    """test/runtests.jl:51-89 recipe (tip load) and BASELINE config 2 (variable-density volume force)."""
    g = golden_c2
    pre = "" if load == "tip" else "vf_"
    st = _solve_case(ctx, pkg, g, simp=True, load=load, matrix_free=matrix_free)
    assert st["converged"] == 1
    u = ctx.solution()
    assert rel(u, g[pre + "u"]) <= TOL_U
    e, c, ee = ctx.energy(per_element=True)
    assert abs(e - float(g[pre + "energy"])) <= TOL_U * float(g[pre + "energy"])
    assert abs(c - float(g[pre + "compliance"])) <= TOL_U * float(g[pre + "compliance"])
    assert np.max(np.abs(ee - g[pre + "elem_energy"])) <= TOL_U * np.abs(g[pre + "elem_energy"]).max()


def test_stresses(ctx, pkg, fo, golden_c1, golden_c2):
    for g, simp in ((golden_c1, False), (golden_c2, True)):
        pts, cells = g["points"], g["cells"].astype(np.int64)
        _setup(ctx, pts, cells)
        prob = fo.setup_problem(pts, cells)
        if simp:
            ctx.set_material_simp(1.0, 0.3, 1e-8, 3.0, g["density"])
            lam, mu = fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)(g["density"])
        else:
            lam, mu = fo.create_material_model(1.0, 0.3)
            ctx.set_material_lame(lam, mu)
        ctx.set_solution(g["u"])
        sig, vm, mx, arg = ctx.stresses(True, True)
        sref, vmref, mxref, argref = fo.calculate_stresses(prob, g["u"], lam, mu)
        s6 = np.stack([sref[..., 0, 0], sref[..., 1, 1], sref[..., 2, 2], sref[..., 0, 1], sref[..., 1, 2], sref[..., 0, 2]], axis=-1)
        assert np.max(np.abs(sig - s6)) <= 1e-10 * np.abs(s6).max()
        assert np.max(np.abs(vm - vmref)) <= 1e-10 * vmref.max()
        assert arg == argref == int(g["max_stress_cell"]) and abs(mx - float(g["max_von_mises"])) <= 1e-9 * mx


def check_calculate_stresses_free_functions(ctx, pkg, fo, golden_c1, golden_c2):      # run by test_gpu_zz_optin_variants.py (child process) and by the emulated suite
    """calculate_stresses(u, dh, cv, λ, μ) / calculate_stresses_simp(u, dh, cv, model, ρ) are free functions in the reference
    (FiniteElementAnalysis.jl:440 / :730): any u, any material — and the ctx's own K, material and solution are not touched."""
    def s6_of(sref):
        return np.stack([sref[..., 0, 0], sref[..., 1, 1], sref[..., 2, 2], sref[..., 0, 1], sref[..., 1, 2], sref[..., 0, 2]], axis=-1)

    # Tet4: the ctx holds K(λ₀, μ₀) and a stored field v; stresses are asked for another field under another material
    g = golden_c1
    pts, cells = g["points"], g["cells"].astype(np.int64)
    _setup(ctx, pts, cells)
    prob = fo.setup_problem(pts, cells)
    lam0, mu0 = fo.create_material_model(1.0, 0.3)
    ctx.assemble_lame(lam0, mu0)
    vals0 = ctx.values()
    v = np.random.default_rng(5).standard_normal(ctx.ndofs)
    ctx.set_solution(v)
    own = ctx.stresses(False, True)
    lam1, mu1 = fo.create_material_model(210.0, 0.25)
    sig, vm, mx, arg = ctx.calculate_stresses(g["u"], lame=(lam1, mu1), want_sigma=True, want_vm=True)
    sref, vmref, mxref, argref = fo.calculate_stresses(prob, g["u"], lam1, mu1)
    assert np.max(np.abs(sig - s6_of(sref))) <= 1e-10 * np.abs(sref).max()
    assert np.max(np.abs(vm - vmref)) <= 1e-10 * vmref.max() and arg == argref and abs(mx - mxref) <= 1e-10 * mxref
    assert np.array_equal(ctx.solution(), v) and np.array_equal(ctx.values(), vals0)            # ctx state untouched
    again = ctx.stresses(False, True)
    assert np.array_equal(again[1], own[1]) and again[2:] == own[2:]
    # u = None → the stored field; same numbers as toe_stresses when the material is the ctx's
    _, vm_n, mx_n, arg_n = ctx.calculate_stresses(None, lame=(lam0, mu0), want_vm=True)
    assert np.array_equal(vm_n, own[1]) and (mx_n, arg_n) == own[2:]
    with pytest.raises(pkg.TopOptError):
        ctx.calculate_stresses(np.zeros(ctx.ndofs + 3), lame=(lam0, mu0))

    # the API mirror: 3-tuple like the reference, lazy stress field indexed by 1-based cell id
    grid = pkg.Grid(pts, cells, 10)
    dh, cv, K, f = pkg.setup_problem(grid)
    sf, mx_a, cell_a = pkg.calculate_stresses(g["u"], dh, cv, lam1, mu1)
    assert cell_a == argref and abs(mx_a - mxref) <= 1e-10 * mxref and len(sf) == cells.shape[0]
    assert np.max(np.abs(sf[argref] - s6_of(sref)[argref - 1])) <= 1e-10 * np.abs(sref).max()

    # Hex8 + SIMP: parametrised model, and the same model hidden in an opaque callable (host-evaluated per cell)
    g = golden_c2
    pts, cells = g["points"], g["cells"].astype(np.int64)
    grid = pkg.Grid(pts, cells, 12)
    dh, cv, K, f = pkg.setup_problem(grid)
    prob = fo.setup_problem(pts, cells)
    mm = pkg.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)
    lam_e, mu_e = fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)(g["density"])
    sref, vmref, mxref, argref = fo.calculate_stresses(prob, g["u"], lam_e, mu_e)
    for model in (mm, lambda rho: mm(rho)):
        sf, mx_s, cell_s = pkg.calculate_stresses_simp(g["u"], dh, cv, model, g["density"])
        assert cell_s == argref == int(g["max_stress_cell"]) and abs(mx_s - mxref) <= 1e-10 * mxref
        assert np.max(np.abs(sf.von_mises - vmref)) <= 1e-10 * vmref.max()
        assert np.max(np.abs(sf.sigma - s6_of(sref))) <= 1e-10 * np.abs(sref).max()
    with pytest.raises(pkg.TopOptError):
        pkg.calculate_stresses_simp(g["u"], dh, cv, mm, g["density"][:-1])
    dh.ctx.close()


# ----------------------------------------------------------------------------------------------------------
# the reference's own test recipes through the API mirror
# ----------------------------------------------------------------------------------------------------------
def test_runtests_recipe_linear_beam(pkg, golden_c1, tmp_path):
    g = golden_c1
    grid = pkg.Grid(g["points"], g["cells"].astype(np.int64), 10)
    assert pkg.calculate_volume(grid) > 0.0
    lam, mu = pkg.create_material_model(1.0, 0.3)
    dh, cv, K, f = pkg.setup_problem(grid)
    pkg.assemble_stiffness_matrix(K, f, dh, cv, lam, mu)
    fixed = pkg.meshgen.nodes_at_plane(grid.nodes, 0, 0.0)
    load = pkg.meshgen.nodes_at_plane(grid.nodes, 0, 60.0)
    assert len(fixed) == 40 and len(load) == 25
    ch = pkg.apply_fixed_boundary(K, f, dh, set(fixed.tolist()))
    assert np.array_equal(ch.prescribed_dofs, g["prescribed"])
    pkg.apply_force(f, dh, list(load), [0.0, 0.0, -1.0])
    u, energy, stress_field, max_vm, max_cell = pkg.solve_system(K, f, dh, cv, lam, mu, ch)
    assert energy > 0.0 and max_vm > 0.0 and np.all(np.isfinite(u))          # the reference's assertions
    assert abs(energy - float(g["energy"])) <= TOL_U * energy and rel(u, g["u"]) <= TOL_U
    assert max_cell == int(g["max_stress_cell"])
    out = pkg.export_results(u, dh, str(tmp_path / "cantilever_beam-linear_u"))
    back = pkg.vtu.read_vtu(out)
    nfd = dh.node_first_dof
    assert np.array_equal(back.point_data["u"][:, 2], u[nfd + 1])
    pkg.export_results(stress_field, dh, str(tmp_path / "cantilever_beam-linear_stress"))
    dh.ctx.close()


def test_runtests_recipe_simp_beam(pkg, golden_c2):
    g = golden_c2
    grid = pkg.Grid(g["points"], g["cells"].astype(np.int64), 12)
    rho = g["density"]
    assert len(rho) == grid.getncells()
    assert abs(pkg.calculate_volume(grid, rho) - rho.sum()) < 1e-9
    mm = pkg.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)
    dh, cv, K, f = pkg.setup_problem(grid)
    pkg.assemble_stiffness_matrix_simp(K, f, dh, cv, mm, rho)
    fixed = pkg.meshgen.nodes_at_plane(grid.nodes, 0, 0.0)
    load = pkg.meshgen.nodes_at_plane(grid.nodes, 0, 60.0)
    ch = pkg.apply_fixed_boundary(K, f, dh, fixed)
    pkg.apply_force(f, dh, load, [0.0, 0.0, -1.0])
    cfg = pkg.SolverConfig(method="cg", tolerance=1e-10, max_iterations=100000, verbose=False)
    u, energy, sf, max_vm, max_cell = pkg.solve_system_robust_simp(K, f, dh, cv, mm, rho, ch, config=cfg)
    assert energy > 0.0 and max_vm > 0.0 and np.all(np.isfinite(u))
    assert abs(energy - float(g["energy"])) <= TOL_U * energy and rel(u, g["u"]) <= TOL_U
    assert sf[1].shape == (8, 6)
    dh.ctx.close()


def test_gravity_cantilever_known_answer(pkg):
    """test/VolumeForces/testVolumeForces.jl:6-60,159-168: 40x8x8 hex cantilever under gravity, tip deflection
    within 10 % of ρgL⁴/(8EI) (clamped with an explicit tolerance)."""
    L, h = 10.0, 1.0
    pts, cells = pkg.meshgen.cantilever(40, 8, 8, L=(L, h, h), hex=True)
    grid = pkg.Grid(pts, cells, 12)
    E, nu, rho, gacc = 200e9, 0.3, 7850.0, 9.81
    lam, mu = pkg.create_material_model(E, nu)
    dh, cv, K, f = pkg.setup_problem(grid)
    pkg.assemble_stiffness_matrix(K, f, dh, cv, lam, mu)
    ch = pkg.apply_fixed_boundary(K, f, dh, pkg.meshgen.nodes_at_plane(pts, 0, 0.0))
    pkg.apply_gravity(f, dh, cv, rho, gacc, [0.0, 0.0, -1.0])
    cfg = pkg.SolverConfig(method="cg", tolerance=1e-10, max_iterations=200000, verbose=False)
    u, energy, *_ = pkg.solve_system_robust(K, f, dh, cv, lam, mu, ch, config=cfg)
    nfd = dh.node_first_dof
    tip = pkg.meshgen.nodes_at_plane(pts, 0, L)
    tip_defl = np.abs(u[nfd[tip - 1] + 1]).max()
    analytic = rho * gacc * L ** 4 / (8 * E * (h * h ** 3 / 12)) * h * h   # q = ρ g A
    assert abs(tip_defl - analytic) / analytic < 0.10
    assert energy > 0
    dh.ctx.close()


def test_synthetic_cantilever_energies(ctx, pkg, golden_syn):
    for dims in ((12, 4, 2), (24, 8, 4)):
        tag = "x".join(map(str, dims))
        pts, cells = pkg.meshgen.cantilever(*dims)
        _setup(ctx, pts, cells)
        lam, mu = pkg.create_material_model(1.0, 0.3)
        for mf in (False, True):
            (ctx.set_material_lame if mf else ctx.assemble_lame)(lam, mu)
            ctx.add_nodal_force(pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0.0, 0.0, -1.0])
            nfd = ctx.node_dofs()
            fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
            pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
            ctx.apply_dirichlet(pres)
            st = ctx.solve_pcg(1e-10, 1e-10, 100000, matrix_free=mf)
            assert st["converged"] == 1
            assert rel(ctx.solution(), golden_syn[tag + "_u"]) <= TOL_U
            e, c, _ = ctx.energy()
            assert abs(e - float(golden_syn[tag + "_energy"])) <= TOL_U * e


def test_error_behaviour(pkg):
    c = pkg.Context(0)
    pts, cells = pkg.meshgen.cantilever(2, 2, 2)
    with pytest.raises(pkg.TopOptError):
        c.build_dofs()                                   # no mesh yet
    with pytest.raises(pkg.TopOptError, match="cell type"):
        c.set_mesh(pts, cells[:, :3])
    bad = cells.copy(); bad[0, 0] = pts.shape[0] + 5
    with pytest.raises(pkg.TopOptError, match="node id"):
        c.set_mesh(pts, bad)
    c.set_mesh(pts, cells); c.build_dofs(); c.build_pattern()
    with pytest.raises(pkg.TopOptError):
        c.solve_pcg()                                    # K not assembled
    c.assemble_lame(1.0, 1.0)
    with pytest.raises(pkg.TopOptError):
        c.apply_dirichlet(np.array([0], dtype=np.int64))
    c.close()


# ----------------------------------------------------------------------------------------------------------
# edge cases
# ----------------------------------------------------------------------------------------------------------
def test_edge_single_cell_and_trivial_solves(ctx, pkg, fo):
    """smallest meshes (one tet, one hex), zero right-hand side, itmax = 0, every free dof loaded."""
    for npc in (4, 8):
        if npc == 4:
            pts = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]); cells = np.array([[1, 2, 3, 4]], dtype=np.int64)
        else:
            pts, cells = pkg.meshgen.cantilever(1, 1, 1, L=(1.0, 1.0, 1.0), hex=True)
        _setup(ctx, pts, cells)
        prob = fo.setup_problem(pts, cells)
        assert ctx.ndofs == 3 * npc and ctx.nnz == (3 * npc) ** 2
        ctx.assemble_lame(0.6, 0.4); fo.assemble_stiffness_matrix(prob, 0.6, 0.4)
        assert np.max(np.abs(ctx.values() - prob.nzval)) <= 1e-13 * np.abs(prob.nzval).max()
        # zero load: converged at iteration 0 with u = 0 (Krylov: sqrt(r0'Mr0) = 0 <= atol)
        st = ctx.solve_pcg(1e-8, 1e-8, 100)
        assert st["niter"] == 0 and st["converged"] == 1 and np.all(ctx.solution() == 0.0)
        # clamp three nodes, pull the others
        nfd = ctx.node_dofs()
        pres = np.sort((nfd[:3][:, None] + np.arange(3)[None, :]).reshape(-1))
        ctx.add_nodal_force(np.arange(4, npc + 1), [0.3, -0.2, 0.1]); fo.apply_force(prob, np.arange(4, npc + 1), [0.3, -0.2, 0.1])
        ctx.apply_dirichlet(pres); fo.apply_dirichlet(prob, pres)
        st0 = ctx.solve_pcg(1e-8, 1e-8, 0)
        assert st0["niter"] == 0 and st0["converged"] == 0
        for mf in (False, True):
            st = ctx.solve_pcg(1e-12, 1e-12, 1000, matrix_free=mf)
            assert st["converged"] == 1 and rel(ctx.solution(), fo.solve_direct(prob)) <= TOL_U


def test_edge_duplicate_load_nodes_and_repeated_solves(ctx, pkg, fo):
    """`apply_force!` with a Vector holding a node twice adds twice (and divides by the list length); loads accumulate
    over calls; a second solve on the same operator (graph replay) with a new right-hand side is independent of the first."""
    pts, cells = pkg.meshgen.cantilever(8, 3, 2)
    _setup(ctx, pts, cells)
    prob = fo.setup_problem(pts, cells)
    lam, mu = fo.create_material_model(2.0, 0.25)
    ctx.assemble_lame(lam, mu); fo.assemble_stiffness_matrix(prob, lam, mu)
    nodes = np.array([5, 9, 9, 14], dtype=np.int64)
    ctx.add_nodal_force(nodes, [1.0, 2.0, 3.0]); fo.apply_force(prob, nodes, [1.0, 2.0, 3.0])
    ctx.add_nodal_force(nodes[:1], [0.0, 0.0, -1.0]); fo.apply_force(prob, nodes[:1], [0.0, 0.0, -1.0])
    assert np.max(np.abs(ctx.rhs() - prob.f)) <= 1e-15
    fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
    pres = fo.fixed_boundary_dofs(prob, fixed)
    ctx.apply_dirichlet(pres); fo.apply_dirichlet(prob, pres)
    st1 = ctx.solve_pcg(1e-11, 1e-11, 50000)
    u1 = ctx.solution()
    assert rel(u1, fo.solve_direct(prob)) <= TOL_U
    f2 = np.random.default_rng(11).standard_normal(ctx.ndofs); f2[pres - 1] = 0.0
    ctx.set_rhs(f2); prob.f[:] = f2
    st2 = ctx.solve_pcg(1e-11, 1e-11, 50000)
    assert st2["converged"] == 1 and rel(ctx.solution(), fo.solve_direct(prob)) <= TOL_U
    ctx.set_rhs(np.zeros(ctx.ndofs))
    assert ctx.solve_pcg(1e-8, 1e-8, 10)["niter"] == 0


def test_edge_sliding_boundary_and_void_material(ctx, pkg, fo):
    """apply_sliding_boundary! (a subset of components fixed) together with a clamp; densities of exactly 0 and 1."""
    pts, cells = pkg.meshgen.cantilever(6, 3, 2)
    rho = np.where(np.arange(cells.shape[0]) % 5 == 0, 0.0, 1.0)
    grid = pkg.Grid(pts, cells, 10)
    mm = pkg.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)
    dh, cv, K, f = pkg.setup_problem(grid, ctx=ctx)
    pkg.assemble_stiffness_matrix_simp(K, f, dh, cv, mm, rho)
    ch1 = pkg.apply_fixed_boundary(K, f, dh, pkg.meshgen.nodes_at_plane(pts, 0, 0.0))
    ch2 = pkg.apply_sliding_boundary(K, f, dh, pkg.meshgen.nodes_at_plane(pts, 2, 0.0), [3])
    pkg.apply_force(f, dh, pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0.0, -1.0, 0.0])
    cfg = pkg.SolverConfig(method="cg", tolerance=1e-11, max_iterations=200000, verbose=False)
    u, energy, *_ = pkg.solve_system_robust_simp(K, f, dh, cv, mm, rho, ch1, ch2, config=cfg)
    prob = fo.setup_problem(pts, cells)
    fo.assemble_stiffness_matrix_simp(prob, fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0), rho)
    fo.apply_force(prob, pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0.0, -1.0, 0.0])
    p1 = fo.fixed_boundary_dofs(prob, pkg.meshgen.nodes_at_plane(pts, 0, 0.0))
    p2 = fo.fixed_boundary_dofs(prob, pkg.meshgen.nodes_at_plane(pts, 2, 0.0), (3,))
    assert np.array_equal(ch1.prescribed_dofs, p1) and np.array_equal(ch2.prescribed_dofs, p2)
    fo.apply_dirichlet(prob, p1); fo.apply_dirichlet(prob, p2)
    uref = fo.solve_direct(prob)
    assert rel(u, uref) <= TOL_U and abs(energy - fo.deformation_energy(prob, uref)) <= TOL_U * energy


def test_edge_arbitrary_material_callable(ctx, pkg, fo):
    """assemble_stiffness_matrix_simp! with a plain callable ρ ↦ (λ, μ) (host-evaluated, per cell)."""
    pts, cells = pkg.meshgen.cantilever(4, 2, 2, hex=True)
    rho = np.linspace(0.2, 1.0, cells.shape[0])
    grid = pkg.Grid(pts, cells, 12)
    model = lambda r: (0.5 + 2.0 * r, 0.25 + r * r)
    dh, cv, K, f = pkg.setup_problem(grid, ctx=ctx)
    pkg.assemble_stiffness_matrix_simp(K, f, dh, cv, model, rho)
    prob = fo.setup_problem(pts, cells)
    fo.assemble_stiffness_matrix(prob, 0.5 + 2.0 * rho, 0.25 + rho * rho)
    assert np.max(np.abs(K.nzval() - prob.nzval)) <= 1e-12 * np.abs(prob.nzval).max()
    Ks = K.to_scipy()
    assert abs(Ks - prob.K()).max() <= 1e-12 * np.abs(prob.nzval).max()


# ----------------------------------------------------------------------------------------------------------
# ragged input: unstructured Delaunay tets (irregular valence, arbitrary cell order, SIMP densities)
# ----------------------------------------------------------------------------------------------------------
def _delaunay_mesh(seed, npts):
    """Random points in the 6x2x1 box → Delaunay tets, positively oriented, slivers dropped.  Node valences range from a handful to
    several dozen cells, nothing like the 24 of the structured split."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((npts, 3)) * np.array([6.0, 2.0, 1.0])
    corners = np.array([(x, y, z) for x in (0.0, 6.0) for y in (0.0, 2.0) for z in (0.0, 1.0)])
    pts = np.vstack([corners, pts])
    pts[8:8 + npts // 6, 0] = 0.0                                      # a populated clamp face ...
    pts[8 + npts // 6:8 + npts // 3, 0] = 6.0                          # ... and a populated load face
    tet = Delaunay(pts).simplices.astype(np.int64)
    X = pts[tet]
    vol = np.linalg.det(np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=2)) / 6.0
    flip = vol < 0
    tet[flip] = tet[flip][:, [0, 2, 1, 3]]
    keep = np.abs(vol) > 1e-5
    return pts, np.ascontiguousarray(tet[keep] + 1)


@pytest.mark.parametrize("seed,npts", [(1, 120), (2, 260)])
def test_unstructured_delaunay_mesh(ctx, pkg, fo, seed, npts):
    pts, cells = _delaunay_mesh(seed, npts)
    ne = cells.shape[0]
    rho = np.random.default_rng(seed + 100).uniform(0.05, 1.0, ne)
    _setup(ctx, pts, cells)
    prob = fo.setup_problem(pts, cells)
    assert np.array_equal(ctx.node_dofs(), prob.node_first_dof)
    colptr, rowval = ctx.pattern()
    assert np.array_equal(colptr, prob.colptr) and np.array_equal(rowval, prob.rowval)
    fo.assemble_stiffness_matrix_simp(prob, fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0), rho)
    scale = _row_scale(prob)
    A = pkg._lib
    vals = {}
    for name, var in (("gather", A.ASM_GATHER), ("atomic", A.ASM_ATOMIC), ("rows", A.ASM_ROWS)):
        ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho, var)
        vals[name] = ctx.values()
        assert np.max(np.abs(vals[name] - prob.nzval) / scale) <= TOL_KE, name
    # loads, constraints, solve against the oracle's direct solve; assembled and matrix-free
    fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 6.0)
    assert fixed.size >= 3 and load.size >= 3
    fo.apply_force(prob, load, [0.0, 0.0, -1.0])
    pres = fo.fixed_boundary_dofs(prob, fixed)
    fo.apply_dirichlet(prob, pres)
    uref = fo.solve_direct(prob)
    eref = fo.deformation_energy(prob, uref)
    x = np.random.default_rng(seed).standard_normal(ctx.ndofs)
    for mf in (False, True):
        (ctx.set_material_simp if mf else ctx.assemble_simp)(1.0, 0.3, 1e-8, 3.0, rho)
        ctx.add_nodal_force(load, [0.0, 0.0, -1.0]); ctx.apply_dirichlet(pres)
        y = ctx.spmv(x, matrix_free=mf)
        assert np.max(np.abs(y - prob.K() @ x)) <= 1e-12 * np.max(np.abs(prob.K()).sum(axis=1)) * np.max(np.abs(x))
        st = ctx.solve_pcg(1e-12, 1e-12, 200000, matrix_free=mf)
        assert st["converged"] == 1 and st["breakdown"] == 0
        assert rel(ctx.solution(), uref) <= TOL_U
        e, c, ee = ctx.energy(per_element=True)
        assert abs(e - eref) <= TOL_U * abs(eref) and abs(ee.sum() - e) <= 1e-11 * abs(e)
