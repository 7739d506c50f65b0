"""The full-size property checks (tests/fullsize_properties.py) at toy sizes: against the EMULATED build of the CUDA
sources (tests/cuda_emu — test infrastructure, see tests/emu_support.py) and with the oracle's solve as the golden value,
so that the checks themselves are known to be right before they run at 1M / 10M tets on a B200."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_support  # noqa: E402
import fullsize_properties as fp  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    pkg, lib = emu_support.load_emu()
    with emu_support.emulated(pkg, lib):
        yield pkg, lib
    assert lib.emu_check_all_guards() == 0


def test_structured_counts_match_the_oracle_pattern(fo, pkg):
    for dims in ((3, 2, 2), (5, 4, 3)):
        pts, cells = pkg.meshgen.cantilever(*dims)
        prob = fo.setup_problem(pts, cells)
        nn, ne, nnz = fp.structured_counts(dims)
        assert (nn, ne, nnz) == (pts.shape[0], cells.shape[0], prob.nnz)
        assert np.array_equal(fp.first_touch_node_dofs(cells, nn), prob.node_first_dof)


@pytest.mark.parametrize("dims,simp", [((8, 3, 2), False), ((6, 4, 3), True)])
def test_properties_hold_on_the_emulated_build(emu, fo, dims, simp):
    pkg, _ = emu
    golden = None
    if not simp:
        pts, cells = pkg.meshgen.cantilever(*dims)
        prob = fo.setup_problem(pts, cells)
        lam, mu = fo.create_material_model(1.0, 0.3)
        fo.assemble_stiffness_matrix(prob, lam, mu)
        fo.apply_force(prob, pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0.0, 0.0, -1.0])
        f = prob.f.copy()
        fo.apply_dirichlet(prob, fo.fixed_boundary_dofs(prob, pkg.meshgen.nodes_at_plane(pts, 0, 0.0)))
        u, stats = fo.solve_pcg(prob, 1e-10, 100000)
        golden = {"energy": fo.deformation_energy(prob, u), "compliance": float(f @ u), "niter": int(stats["niter"])}
    ctx = pkg.Context(0)
    try:
        out = fp.run_properties(pkg, ctx, dims, simp=simp, golden=golden, check_pattern=True, tol_solve=1e-10, itmax=100000)
    finally:
        ctx.close()
    assert out["niter"] > 0
