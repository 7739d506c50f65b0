"""Checks of the two-level preconditioner (TOE_PCG_TWO_LEVEL), shared by the GPU test and the emulated run."""
import os

import numpy as np


def numpy_two_level_iterations(fo, pkg, pts, cells, boxes, tol):
    """The same preconditioner in numpy/scipy on the oracle's K: M⁻¹ = D⁻¹ + Z (ZᵀKZ)⁻¹ Zᵀ, Z = rigid-body modes of the boxes
    (rotations about the box centres), prescribed DOFs masked; Krylov.jl stopping rule."""
    import scipy.sparse as sp
    prob = fo.setup_problem(pts, cells)
    lam, mu = fo.create_material_model(1.0, 0.3)
    fo.assemble_stiffness_matrix(prob, lam, mu)
    fo.apply_force(prob, pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0.0, 0.0, -1.0])
    pres = fo.fixed_boundary_dofs(prob, pkg.meshgen.nodes_at_plane(pts, 0, 0.0))
    fo.apply_dirichlet(prob, pres)
    K = prob.K().tocsr(); b = prob.f
    used = prob.node_first_dof > 0
    P = pts[used]; nfd = prob.node_first_dof[used]
    lo, hi = P.min(0), P.max(0)
    bx = np.array(boxes)
    h = (hi - lo) / bx
    idx = np.clip(np.floor((P - lo) / h).astype(int), 0, bx - 1)
    agg = idx[:, 0] + bx[0] * (idx[:, 1] + bx[1] * idx[:, 2])
    cen = lo + (idx + 0.5) * h
    d = P - cen
    rows, cols, vals = [], [], []
    for c in range(3):
        rows.append(nfd - 1 + c); cols.append(6 * agg + c); vals.append(np.ones(P.shape[0]))
    for comp, k, v in ((1, 0, -d[:, 2]), (2, 0, d[:, 1]), (0, 1, d[:, 2]), (2, 1, -d[:, 0]), (0, 2, -d[:, 1]), (1, 2, d[:, 0])):
        rows.append(nfd - 1 + comp); cols.append(6 * agg + 3 + k); vals.append(v)
    Z = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(prob.ndofs, 6 * int(np.prod(bx))))
    mask = np.ones(prob.ndofs); mask[pres - 1] = 0.0
    Z = sp.diags(mask) @ Z
    Ac = (Z.T @ K @ Z).toarray()
    keep = np.abs(np.diag(Ac)) > 1e-13 * np.abs(np.diag(Ac)).max()
    Z = Z[:, keep]; Ak = Ac[np.ix_(keep, keep)]
    Aci = np.linalg.pinv(0.5 * (Ak + Ak.T), rcond=1e-12, hermitian=True)      # boxes with few free nodes make the modes dependent: drop the null space
    Dinv = fo.jacobi_preconditioner(prob)
    M = lambda r: Dinv * r + Z @ (Aci @ (Z.T @ r))
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); gam = r @ z; eps = tol + tol * np.sqrt(gam); k = 0
    while np.sqrt(gam) > eps and k < 100000:
        Ap = K @ p; a = gam / (p @ Ap); x += a * p; r -= a * Ap; z = M(r); g2 = r @ z; p = z + (g2 / gam) * p; gam = g2; k += 1
    return k, x, fo.solve_direct(prob)


def check_two_level(pkg, fo, ctx, cases, tol=1e-10, auto_dims=(24, 8, 4), light_after_first=False):
    """light_after_first (emulated runs, where a two-level iteration costs 25 ms): graph replay and the probing cross-check only on the first case"""
    try:
        for icase, (dims, hexm, boxes, simp) in enumerate(cases):
            light = light_after_first and icase > 0
            os.environ["TOE_TL_BOXES"] = ",".join(map(str, boxes))
            pts, cells = pkg.meshgen.cantilever(*dims, hex=hexm)
            ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
            lam, mu = pkg.create_material_model(1.0, 0.3)
            rho = pkg.meshgen.simp_like_density(cells.shape[0])
            fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
            nfd = ctx.node_dofs()
            pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
            ref = None
            if not simp and not hexm:
                ref = numpy_two_level_iterations(fo, pkg, pts, cells, boxes, tol)
            for mf in (False, True):
                if simp:
                    (ctx.set_material_simp if mf else ctx.assemble_simp)(1.0, 0.3, 1e-8, 3.0, rho)
                else:
                    (ctx.set_material_lame if mf else ctx.assemble_lame)(lam, mu)
                ctx.add_nodal_force(load, [0.0, 0.0, -1.0]); ctx.apply_dirichlet(pres)
                s0 = ctx.solve_pcg(tol, tol, 100000, matrix_free=mf); u0 = ctx.solution()
                runs = []
                for graph in ((False,) if light else (False, True)):
                    s1 = ctx.solve_pcg(tol, tol, 100000, matrix_free=mf, two_level=True, graph=graph, history=True)
                    u1 = ctx.solution()
                    assert s1["converged"] == 1 and s1["breakdown"] == 0 and s1["coarse_dofs"] == 6 * int(np.prod(boxes))
                    assert s1["niter"] < s0["niter"], (s1["niter"], s0["niter"])
                    assert np.linalg.norm(u1 - u0) <= 1e-8 * np.linalg.norm(u0)
                    assert np.all(u1[pres - 1] == 0.0)
                    res = s1["residuals"]
                    assert res.size == s1["niter"] + 1 and res[-1] <= tol + tol * res[0]
                    runs.append((s1["niter"], u1))
                assert light or (runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1]))          # graph replay = direct launches, bit for bit
                if not mf and not light:      # assembled K: the coarse operator comes straight from the blocks; the probing path must agree with it
                    os.environ["TOE_TL_PROBE"] = "1"
                    try:
                        if simp:
                            ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho)
                        else:
                            ctx.assemble_lame(lam, mu)
                        ctx.add_nodal_force(load, [0.0, 0.0, -1.0]); ctx.apply_dirichlet(pres)
                        sp_ = ctx.solve_pcg(tol, tol, 100000, two_level=True)
                        assert sp_["converged"] == 1 and abs(sp_["niter"] - runs[0][0]) <= 2
                        assert np.linalg.norm(ctx.solution() - runs[0][1]) <= 1e-8 * np.linalg.norm(runs[0][1])
                    finally:
                        os.environ.pop("TOE_TL_PROBE", None)
                e, c, _ = ctx.energy()
                assert abs(c - 2.0 * e) <= 1e-6 * c
                if ref is not None:
                    k_ref, x_ref, u_direct = ref
                    assert abs(runs[0][0] - k_ref) <= max(3, k_ref // 20), (runs[0][0], k_ref)    # same preconditioner as the numpy restatement
                    assert np.linalg.norm(runs[0][1] - u_direct) <= 1e-8 * np.linalg.norm(u_direct)
        # automatic box choice (no TOE_TL_BOXES): halve the longest box edge until ≥ target boxes
        os.environ.pop("TOE_TL_BOXES", None)
        os.environ["TOE_TL_BOXES_TARGET"] = "12"
        pts, cells = pkg.meshgen.cantilever(*auto_dims)
        ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
        ctx.assemble_lame(*pkg.create_material_model(1.0, 0.3))
        fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
        nfd = ctx.node_dofs(); pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
        ctx.add_nodal_force(load, [0.0, 0.0, -1.0]); ctx.apply_dirichlet(pres)
        s = ctx.solve_pcg(1e-9, 1e-9, 100000, two_level=True)
        assert s["converged"] == 1 and s["coarse_dofs"] == 6 * 16                       # 60x20x4 → (8,2,1) boxes
    finally:
        os.environ.pop("TOE_TL_BOXES", None); os.environ.pop("TOE_TL_BOXES_TARGET", None)
