"""Checks of the pipelined matrix-free operator (TOE_EBE_PIPE=1), shared by the GPU test and the emulated run: it must give
the same BITS as the tile kernel (same arithmetic, same summation order), whatever the number of tiles a CTA walks."""
import os

import numpy as np


def check_pipe_equals_tile(pkg, ctx, cases, grids=(None,), solve=True):
    for dims, hexm in cases:
        pts, cells = pkg.meshgen.cantilever(*dims, hex=hexm)
        ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
        rho = pkg.meshgen.simp_like_density(cells.shape[0])
        ctx.set_material_simp(1.0, 0.3, 1e-8, 3.0, rho)
        x = np.random.default_rng(1).standard_normal(ctx.ndofs)
        os.environ.pop("TOE_EBE_PIPE", None)
        y_tile = ctx.spmv(x, matrix_free=True)
        try:
            os.environ["TOE_EBE_PIPE"] = "1"
            for g in grids:
                if g is not None:
                    os.environ["TOE_EBE_PIPE_GRID"] = str(g)
                y_pipe = ctx.spmv(x, matrix_free=True)
                assert np.array_equal(y_tile, y_pipe), (dims, hexm, g)
            os.environ.pop("TOE_EBE_PIPE_GRID", None)
            if solve:
                fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
                nfd = ctx.node_dofs()
                pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
                res = {}
                for pipe in (False, True):
                    if pipe:
                        os.environ["TOE_EBE_PIPE"] = "1"
                    else:
                        os.environ.pop("TOE_EBE_PIPE", None)
                    ctx.set_material_simp(1.0, 0.3, 1e-8, 3.0, rho)
                    ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
                    ctx.apply_dirichlet(pres)
                    for graph in (False, True):
                        st = ctx.solve_pcg(1e-9, 1e-9, 100000, matrix_free=True, graph=graph)
                        assert st["converged"] == 1
                        res[(pipe, graph)] = (st["niter"], ctx.solution())
                base = res[(False, False)]
                for k, v in res.items():
                    assert v[0] == base[0] and np.array_equal(v[1], base[1]), (dims, hexm, k)
        finally:
            os.environ.pop("TOE_EBE_PIPE", None)
            os.environ.pop("TOE_EBE_PIPE_GRID", None)
