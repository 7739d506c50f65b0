"""Checks of the ROWS assembly variant (TOE_ASM_ROWS), shared by the GPU test and the emulated run."""
import numpy as np


def check_rows_variant(pkg, fo, ctx, golden_c1):
    A = pkg._lib
    cases = []
    pts, cells = golden_c1["points"], golden_c1["cells"].astype(np.int64)             # unstructured fixture: rows up to 34 blocks (G = 64)
    cases.append(("c1", pts, cells, None))
    p2, c2 = pkg.meshgen.cantilever(8, 3, 2)                                            # structured: rows up to 15 blocks (G = 16)
    cases.append(("syn", p2, c2, None))
    cases.append(("syn-simp", p2, c2, pkg.meshgen.simp_like_density(c2.shape[0])))
    rng = np.random.default_rng(3)
    p3 = p2.copy(); c3 = c2[rng.permutation(c2.shape[0])]                               # permuted cells: rows whose cells are far apart in cell order
    cases.append(("syn-perm", p3, c3, None))
    for name, pts, cells, rho in cases:
        ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
        prob = fo.setup_problem(pts, cells)
        if rho is None:
            lam, mu = fo.create_material_model(1.0, 0.3)
            fo.assemble_stiffness_matrix(prob, lam, mu)
            ctx.assemble_lame(lam, mu, A.ASM_GATHER); vg = ctx.values()
            ctx.assemble_lame(lam, mu, A.ASM_ROWS); vr = ctx.values()
            ctx.assemble_lame(lam, mu, A.ASM_ROWS); vr2 = ctx.values()
        else:
            fo.assemble_stiffness_matrix_simp(prob, fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0), rho)
            ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho, A.ASM_GATHER); vg = ctx.values()
            ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho, A.ASM_ROWS); vr = ctx.values()
            vr2 = vr
        K = prob.K().tocsr()
        rowmax = np.maximum.reduceat(np.abs(K.data), K.indptr[:-1])
        scale = rowmax[prob.K().indices]
        assert np.max(np.abs(vr - prob.nzval) / scale) <= 1e-12, name            # against the oracle: 1e-12 of the row scale
        assert np.max(np.abs(vr - vg) / scale) <= 1e-14, name                      # against GATHER: last bits only
        assert np.array_equal(vr, vr2), name                                        # bit-reproducible
        colptr, rowval = ctx.pattern()
        import scipy.sparse as sp
        Kr = sp.csc_matrix((vr, rowval - 1, colptr - 1), shape=(ctx.ndofs, ctx.ndofs))
        assert (Kr - Kr.T).nnz == 0 or abs(Kr - Kr.T).max() == 0.0, name           # exactly symmetric
    # inverted cell: det(J) <= 0 is an error in this variant too
    bad = c2.copy(); bad[5, [0, 1]] = bad[5, [1, 0]]
    ctx.set_mesh(p2, bad); ctx.build_dofs(); ctx.build_pattern()
    try:
        ctx.assemble_lame(0.5, 0.4, A.ASM_ROWS)
    except pkg.TopOptError as ex:
        assert "det" in str(ex)
    else:
        raise AssertionError("inverted cell not rejected")
    # a full solve on the ROWS-assembled K reproduces the oracle's direct solve
    ctx.set_mesh(p2, c2); ctx.build_dofs(); ctx.build_pattern()
    prob = fo.setup_problem(p2, c2)
    lam, mu = fo.create_material_model(1.0, 0.3)
    fo.assemble_stiffness_matrix(prob, lam, mu)
    ctx.assemble_lame(lam, mu, A.ASM_ROWS)
    load = pkg.meshgen.nodes_at_plane(p2, 0, 60.0); fixed = pkg.meshgen.nodes_at_plane(p2, 0, 0.0)
    ctx.add_nodal_force(load, [0.0, 0.0, -1.0]); fo.apply_force(prob, load, [0.0, 0.0, -1.0])
    pres = fo.fixed_boundary_dofs(prob, fixed)
    ctx.apply_dirichlet(pres); fo.apply_dirichlet(prob, pres)
    st = ctx.solve_pcg(1e-11, 1e-11, 50000)
    uref = fo.solve_direct(prob)
    assert st["converged"] == 1 and np.linalg.norm(ctx.solution() - uref) <= 1e-8 * np.linalg.norm(uref)
