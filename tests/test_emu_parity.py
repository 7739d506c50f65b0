"""The GPU parity tests, re-run on the host against the EMULATED build of the same CUDA sources (tests/cuda_emu).

TEST INFRASTRUCTURE: checks kernel logic (indexing, initialisation — emulated cudaMalloc returns 0xFF-filled memory —
barriers, the mbarrier pipeline protocol, CG recurrences) where there is no GPU.  It proves nothing about speed and is not
a product path: the package itself only loads libtopopt_b200.so (see tests/emu_support.py).  The real parity gate is
tests/test_gpu_parity.py on a B200; the tests below are the very same functions with the `ctx` / `pkg` fixtures swapped."""
import os
import sys
import types

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
    """Re-exports test_gpu_parity.<name> here.  NEVER mark the imported function itself: pytest.mark.* mutates the function object in
    place, so a skip mark set here would also skip the original in test_gpu_parity.py on the GPU box (round-1 bug).  A slow test gets
    a COPY of the function (own __dict__, hence own marks; same code, same signature for fixture lookup) and only the copy is marked."""
    fn = getattr(gp, name)
    if slow and FAST:
        clone = types.FunctionType(fn.__code__, fn.__globals__, name, fn.__defaults__, fn.__closure__)
        clone.__kwdefaults__ = fn.__kwdefaults__
        clone.__doc__ = fn.__doc__
        clone.__module__ = __name__
        # parametrize marks of the original are kept (so the fixture closure stays valid), the gpu mark is irrelevant for a skipped test
        clone.pytestmark = list(getattr(fn, "pytestmark", [])) + [pytest.mark.skip(reason="long solve under emulation; set TOE_EMU_FULL=1").mark]
        fn = clone
    globals()[name] = fn


for _n in ["test_dofs_and_pattern_bit_exact", "test_dofs_permuted_cells_and_unreferenced_nodes", "test_ke_tet_fixture", "test_ke_hex_simp_fixture",
           "test_ke_partial_range_and_errors", "test_assembled_K_tet", "test_assembled_K_hex_simp", "test_gather_assembly_is_deterministic_and_symmetric",
           "test_per_cell_lame_matches_simp", "test_loads", "test_volume_force_tet", "test_dirichlet_ferrite_semantics", "test_spmv_assembled_and_matrix_free",
           "test_stresses", "test_error_behaviour", "test_edge_single_cell_and_trivial_solves", "test_edge_duplicate_load_nodes_and_repeated_solves",
           "test_edge_sliding_boundary_and_void_material", "test_edge_arbitrary_material_callable", "test_l2_norm_criterion_and_true_residual", "test_unstructured_delaunay_mesh"]:
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
    tc.check_two_level(pkg, fo, ctx, [((6, 3, 2), False, (4, 2, 1), False), ((5, 2, 2), False, (3, 2, 2), True), ((5, 3, 2), True, (3, 1, 1), False)], auto_dims=(12, 4, 2), light_after_first=True)


def test_c_example_against_the_oracle(emu, fo, tmp_path):
    """examples/cantilever.c (plain C99 host, C ABI only) linked against the EMULATED library: the energy it prints equals the oracle's
    direct solve of the same structured mesh — the C call sequence, its 1-based ids and its mesh generator are right."""
    import re
    import subprocess
    pkg, lib = emu
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(emu_support.EMU_LIB)
    exe = tmp_path / "cantilever_emu"
    r = subprocess.run(["gcc", "-std=c99", "-O1", os.path.join(root, "examples", "cantilever.c"), "-I" + os.path.join(root, "include"),
                        "-L" + libdir, "-ltopopt_emu", "-Wl,-rpath," + libdir, "-lm", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"} if not emu_support.ASAN else dict(os.environ)
    r = subprocess.run([str(exe), "8", "3", "2"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"(\d+) tets, (\d+) DOFs, nnz (\d+) .* deformation energy ([0-9.eE+-]+), compliance ([0-9.eE+-]+), max von Mises ([0-9.eE+-]+) in cell (\d+)", r.stdout)
    assert m, r.stdout
    pts, cells = pkg.meshgen.cantilever(8, 3, 2)
    prob = fo.setup_problem(pts, cells)
    lam, mu = fo.create_material_model(1.0, 0.3)
    fo.assemble_stiffness_matrix(prob, lam, mu)
    fo.apply_force(prob, pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0.0, 0.0, -1.0])
    fo.apply_dirichlet(prob, fo.fixed_boundary_dofs(prob, pkg.meshgen.nodes_at_plane(pts, 0, 0.0)))
    u = fo.solve_direct(prob)
    e_ref = fo.deformation_energy(prob, u)
    _, _, mx_ref, arg_ref = fo.calculate_stresses(prob, u, lam, mu)
    assert (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (cells.shape[0], prob.ndofs, prob.nnz)
    assert abs(float(m.group(4)) - e_ref) <= 1e-6 * e_ref and abs(float(m.group(5)) - 2 * e_ref) <= 1e-6 * 2 * e_ref
    assert abs(float(m.group(6)) - mx_ref) <= 1e-5 * mx_ref and int(m.group(7)) == arg_ref
