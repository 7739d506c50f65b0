"""Boundary-node selection and surface traction against the oracle restatement (shared by the GPU test and the emulated run)."""
import numpy as np


def check_surface(pkg, fo, golden_c1):
    cases = [(pkg.meshgen.cantilever(6, 3, 2), 10), (pkg.meshgen.cantilever(5, 3, 2, hex=True), 12),
             ((golden_c1["points"], golden_c1["cells"].astype(np.int64)), 10)]
    for ci, ((pts, cells), ct) in enumerate(cases):
        grid = pkg.Grid(pts, cells, ct)
        dh, cv, K, f = pkg.setup_problem(grid)
        ctx = dh.ctx
        prob = fo.setup_problem(pts, cells)
        # surface nodes: Dict-counting restatement vs the device search
        surf = ctx.surface_nodes()
        assert np.array_equal(surf, fo.extract_surface_nodes(cells))
        # plane / circle selection (reference defaults: tolerance 1.0)
        for point, normal, tol in (([60.0, 0, 0], [1.0, 0, 0], 1e-6), ([0.0, 0, 0], [2.0, 0, 0], 1.0), ([30.0, 10, 2], [0.3, -0.2, 1.0], 1.5)):
            got = pkg.select_nodes_by_plane(grid, point, normal, tol)
            assert got == set(fo.select_nodes_by_plane(pts, cells, point, normal, tol).tolist())
        assert pkg.select_nodes_by_plane(grid, [0.0, 0, 0], [1.0, 0, 0]) == set(fo.select_nodes_by_plane(pts, cells, [0.0, 0, 0], [1.0, 0, 0]).tolist())
        for center, normal, radius, tol in (([60.0, 10, 2], [1.0, 0, 0], 5.0, 1e-6), ([60.0, 10, 2], [1.0, 0, 0], 2.0, 1.0), ([30.0, 20, 2], [0, 1.0, 0], 12.0, 0.5)):
            got = pkg.select_nodes_by_circle(grid, center, normal, radius, tol)
            assert got == set(fo.select_nodes_by_circle(pts, cells, center, normal, radius, tol).tolist())
        # facets of the tip face, area, uniform traction
        tip = pkg.select_nodes_by_plane(grid, [60.0, 0, 0], [1.0, 0, 0], 1e-6)
        facets = pkg.get_boundary_facets(grid, tip)
        ref_facets = fo.get_boundary_facets(cells, sorted(tip))
        assert np.array_equal(facets, ref_facets) and facets.shape[0] > 0
        area = pkg.compute_boundary_area(grid, dh, facets)
        assert abs(area - fo.compute_boundary_area(pts, cells, ref_facets)) <= 1e-12 * area and (ci == 2 or abs(area - 80.0) <= 1e-9)   # box meshes: the whole 20 x 4 tip face
        lam, mu = pkg.create_material_model(1.0, 0.3)
        pkg.assemble_stiffness_matrix(K, f, dh, cv, lam, mu)
        a2, tot = pkg.apply_uniform_surface_traction(f, dh, grid, facets, [0.0, 0.5, -1.0])
        ra, rtot = fo.apply_uniform_surface_traction(prob, ref_facets, [0.0, 0.5, -1.0])
        assert abs(a2 - ra) <= 1e-12 * ra and np.max(np.abs(tot - rtot)) <= 1e-12
        assert np.max(np.abs(f.to_numpy() - prob.f)) <= 1e-13 * np.abs(prob.f).max()
        assert abs(f.to_numpy().reshape(-1, 3)[:, 2].sum() + 1.0) <= 1e-12
        # position-dependent traction through the host callback, on every surface facet
        all_facets = pkg.get_boundary_facets(grid, set(surf.tolist()))
        ref_all = fo.get_boundary_facets(cells, surf)
        assert np.array_equal(all_facets, ref_all)
        fn = lambda x, y, z: [0.01 * x, -0.02 * y * z, 0.3 + z]
        f0 = f.to_numpy().copy(); p0 = prob.f.copy()
        a3, t3 = pkg.apply_surface_traction(f, dh, grid, all_facets, fn)
        ra3, rt3 = fo.apply_surface_traction(prob, ref_all, fn)
        assert abs(a3 - ra3) <= 1e-12 * ra3 and np.max(np.abs(t3 - rt3)) <= 1e-11 * np.abs(rt3).max()
        assert np.max(np.abs((f.to_numpy() - f0) - (prob.f - p0))) <= 1e-12 * np.abs(prob.f - p0).max()
        # quadrature points as sets per facet (the order inside a facet is a convention)
        xq, dg = ctx.facet_quadrature(facets)
        rxq, rdg, _ = fo.facet_quadrature(pts, cells, ref_facets)
        assert np.max(np.abs(np.sort(xq, axis=1) - np.sort(rxq, axis=1))) <= 1e-12 * 60.0 and np.max(np.abs(dg.sum(axis=1) - rdg.sum(axis=1))) <= 1e-12 * area
        # errors: empty selection gives zero area -> the reference's error; bad ids
        none = pkg.get_boundary_facets(grid, set())
        assert none.shape == (0, 2)
        try:
            pkg.apply_uniform_surface_traction(f, dh, grid, none, [0.0, 0.0, -1.0])
        except pkg.TopOptError as ex:
            assert "zero" in str(ex)
        else:
            raise AssertionError("zero-area traction not rejected")
        for bad in (np.array([[cells.shape[0] + 1, 1]]), np.array([[1, 9]])):
            try:
                ctx.boundary_area(bad)
            except pkg.TopOptError:
                pass
            else:
                raise AssertionError("bad facet accepted")
        nd = pkg.get_node_dofs(dh)
        g = int(cells[0, 0])
        assert g in nd and nd[g] == [int(dh.node_first_dof[g - 1]) + k for k in range(3)] and len(nd) == np.count_nonzero(dh.node_first_dof)
        ctx.close()
