"""CPU tests of the host-side logic: VTU reader/writer dialect, mesh generator, API-mirror helpers."""
import os

import numpy as np
import pytest


def test_meshgen_structured_tets(pkg, fo):
    pts, cells = pkg.meshgen.cantilever(6, 4, 2)
    assert pts.shape == (7 * 5 * 3, 3) and cells.shape == (6 * 4 * 2 * 6, 4)
    X = pts[cells - 1]
    J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=2)
    assert np.all(np.linalg.det(J) > 0)
    assert abs(np.linalg.det(J).sum() / 6 - 60 * 20 * 4) < 1e-9
    prob = fo.setup_problem(pts, cells)
    deg = np.diff(prob.colptr)
    assert deg.max() == 45                                   # 15 block neighbours per interior node
    assert pkg.meshgen.SIZES["C4_10M"] == (260, 110, 58) and 260 * 110 * 58 * 6 == 9952800


def test_meshgen_hex(pkg):
    pts, cells = pkg.meshgen.cantilever(3, 2, 2, hex=True)
    assert cells.shape == (12, 8)
    grid = pkg.Grid(pts, cells, 12)
    assert abs(pkg.calculate_volume(grid) - 4800.0) < 1e-9
    p2, c2 = pkg.meshgen.cantilever(3, 2, 2)
    assert abs(pkg.calculate_volume(pkg.Grid(p2, c2, 10)) - 4800.0) < 1e-9


def test_vtu_roundtrip(pkg, tmp_path):
    pts, cells = pkg.meshgen.cantilever(3, 2, 1)
    rho = np.linspace(0, 1, cells.shape[0])
    u = np.random.default_rng(0).standard_normal((pts.shape[0], 3))
    path = pkg.vtu.write_vtu(str(tmp_path / "m"), pts, cells, 10, point_data={"u": u}, cell_data={"density": rho})
    assert path.endswith(".vtu")
    m = pkg.vtu.read_vtu(path)
    assert np.array_equal(m.points, pts) and np.array_equal(m.cells, cells) and m.cell_type == 10
    assert np.array_equal(m.point_data["u"], u)
    assert np.array_equal(pkg.vtu.extract_cell_density(path), rho)
    g = pkg.import_mesh(path)
    assert g.getncells() == cells.shape[0] and g.getnnodes() == pts.shape[0]
    with pytest.raises(pkg.TopOptError):
        pkg.import_mesh(str(tmp_path / "m.msh"))


@pytest.mark.skipif(not os.path.exists("/root/reference/data"), reason="reference fixtures only exist in the build container")
def test_vtu_reads_reference_fixtures(pkg, golden_c1, golden_c2):
    m1 = pkg.vtu.read_vtu("/root/reference/data/beam_linear_volume_mesh.vtu")
    assert np.array_equal(m1.points, golden_c1["points"]) and np.array_equal(m1.cells, golden_c1["cells"]) and m1.cell_type == 10
    m2 = pkg.vtu.read_vtu("/root/reference/data/beam_vfrac_04_Raw.vtu")
    assert np.array_equal(m2.cells, golden_c2["cells"]) and m2.cell_type == 12
    assert np.array_equal(pkg.vtu.extract_cell_density("/root/reference/data/beam_vfrac_04_Raw.vtu"), golden_c2["density"])


def test_material_models(pkg, fo):
    assert pkg.create_material_model(1.0, 0.3) == fo.create_material_model(1.0, 0.3)
    mm = pkg.create_simp_material_model(1.0, 0.3)
    assert (mm.Emin, mm.p) == (1e-6, 1.0)                     # code defaults, FiniteElementAnalysis.jl:619-620
    ref = fo.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)
    mm = pkg.create_simp_material_model(1.0, 0.3, 1e-8, 3.0)
    for rho in (0.0, 0.3, 1.0):
        assert mm(rho) == ref(rho)


def test_solver_config(pkg):
    c = pkg.SolverConfig()
    assert c.tolerance == 1e-8 and c.max_iterations == 10000 and c.preconditioner == "diagonal"
    with pytest.raises(pkg.TopOptError):
        pkg.SolverConfig(method="gmres")
    with pytest.raises(pkg.TopOptError):
        pkg.SolverConfig(preconditioner="ilu")
