"""The GPU parity tests, re-run on the host against the EMULATED build of the same CUDA sources (tests/cuda_emu).

TEST INFRASTRUCTURE: checks kernel logic (indexing, initialisation — emulated cudaMalloc returns 0xFF-filled memory —
barriers, the mbarrier pipeline protocol, CG recurrences) where there is no GPU.  It proves nothing about speed and is not
a product path: the package itself only loads libtopopt_b200.so (see tests/emu_support.py).  The real parity gate is
tests/test_gpu_parity.py on a B200; the tests below are the very same functions with the `ctx` / `pkg` fixtures swapped."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_support  # noqa: E402
import test_gpu_parity as gp  # noqa: E402

FAST = os.environ.get("TOE_EMU_FULL") != "1"        # the long solves (thousands of PCG iterations on the fixtures) only on request


@pytest.fixture(scope="module")
def emu():
    pkg, lib = emu_support.load_emu()
    with emu_support.emulated(pkg, lib):
        yield pkg, lib
    assert lib.emu_check_all_guards() == 0, "an emulated device allocation was written out of bounds"


@pytest.fixture(scope="module")
def pkg(emu):
    return emu[0]


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def _adopt(name, slow=False):
    fn = getattr(gp, name)
    if slow and FAST:
        fn = pytest.mark.skip(reason="long solve under emulation; set TOE_EMU_FULL=1")(fn)
    globals()[name] = fn


for _n in ["test_dofs_and_pattern_bit_exact", "test_dofs_permuted_cells_and_unreferenced_nodes", "test_ke_tet_fixture", "test_ke_hex_simp_fixture",
           "test_ke_partial_range_and_errors", "test_assembled_K_tet", "test_assembled_K_hex_simp", "test_gather_assembly_is_deterministic_and_symmetric",
           "test_per_cell_lame_matches_simp", "test_loads", "test_volume_force_tet", "test_dirichlet_ferrite_semantics", "test_spmv_assembled_and_matrix_free",
           "test_stresses", "test_error_behaviour", "test_edge_single_cell_and_trivial_solves", "test_edge_duplicate_load_nodes_and_repeated_solves",
           "test_edge_sliding_boundary_and_void_material", "test_edge_arbitrary_material_callable"]:
    _adopt(_n)
for _n in ["test_synthetic_cantilever_energies", "test_solve_c1_tet_beam", "test_pcg_krylov_semantics_and_iteration_count", "test_solve_c2_hex_simp", "test_runtests_recipe_linear_beam",
           "test_runtests_recipe_simp_beam", "test_gravity_cantilever_known_answer"]:
    _adopt(_n, slow=True)


def test_calculate_stresses_free_functions(ctx, pkg, fo, golden_c1, golden_c2):
    gp.check_calculate_stresses_free_functions(ctx, pkg, fo, golden_c1, golden_c2)


def test_rows_assembly_variant(ctx, pkg, fo, golden_c1):
    import rows_variant_checks as rc
    rc.check_rows_variant(pkg, fo, ctx, golden_c1)
    for g in ("32", "64"):                            # the wider thread groups on the structured mesh as well
        os.environ["TOE_ASM_ROWS_G"] = g
        try:
            rc.check_rows_variant(pkg, fo, ctx, golden_c1)
        finally:
            del os.environ["TOE_ASM_ROWS_G"]


def test_pipelined_matrix_free_operator(ctx, pkg):
    import ebe_pipe_checks as pc
    pc.check_pipe_equals_tile(pkg, ctx, [((30, 3, 2), False), ((40, 4, 3), False), ((33, 5, 4), True)], grids=(None, 1, 2, 3, 5), solve=False)
    pc.check_pipe_equals_tile(pkg, ctx, [((8, 3, 2), False), ((6, 3, 2), True)], grids=(None, 1), solve=True)


def test_boundary_selection_and_surface_traction(pkg, fo, golden_c1):
    import surface_checks as sc
    sc.check_surface(pkg, fo, golden_c1)


def test_two_level_preconditioner(ctx, pkg, fo):
    import two_level_checks as tc
    tc.check_two_level(pkg, fo, ctx, [((8, 3, 2), False, (4, 2, 1), False), ((5, 2, 2), False, (3, 2, 2), True), ((5, 3, 2), True, (3, 1, 1), False)], auto_dims=(12, 4, 2), light_after_first=True)
