"""Opt-in kernel variants on a B200 (ROWS assembly, pipelined matrix-free operator).  The file name sorts last on purpose: these
variants were written after the round-1 GPU budget was spent (logic verified on the emulated build) and are not defaults yet."""
import pytest

import rows_variant_checks as rc

pytestmark = pytest.mark.gpu


def test_rows_variant(pkg, fo, golden_c1):
    ctx = pkg.Context(0)
    try:
        rc.check_rows_variant(pkg, fo, ctx, golden_c1)
    finally:
        ctx.close()


def test_pipelined_matrix_free_operator_equals_tile_kernel(pkg):
    """221k tets = 864 tiles on ≤ 444 persistent CTAs (several tiles per CTA), plus a Hex8 mesh and a forced 37-CTA grid."""
    import ebe_pipe_checks as pc
    ctx = pkg.Context(0)
    try:
        pc.check_pipe_equals_tile(pkg, ctx, [((96, 32, 12), False), ((40, 16, 8), True)], grids=(None, 37), solve=False)
        pc.check_pipe_equals_tile(pkg, ctx, [((24, 8, 4), False)], grids=(None, 3), solve=True)
    finally:
        ctx.close()


def test_boundary_selection_and_surface_traction(pkg, fo, golden_c1):
    """SURVEY §8(f) next-row 3 on the GPU: surface extraction, plane / circle selection, facets, area, uniform and callback
    traction — against the oracle's literal restatement of SelectNodesForBC.jl / SurfaceTraction.jl."""
    import surface_checks as sc
    sc.check_surface(pkg, fo, golden_c1)
