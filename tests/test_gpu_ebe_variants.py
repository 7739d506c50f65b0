"""Both matrix-free forms (tile, default; node-gather via TOE_EBE_GATHER=1) must agree with the assembled operator."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
import numpy as np
sys.path.insert(0, %(root)r)
import __graft_entry__ as graft
pkg = graft.load_package()
rng = np.random.default_rng(1)
for hexmesh in (False, True):
    pts, cells = pkg.meshgen.cantilever(13, 5, 3, hex=hexmesh)
    cells = cells[rng.permutation(cells.shape[0])]               # scrambled cell order: tiles are not spatially compact
    rho = pkg.meshgen.simp_like_density(cells.shape[0])
    ctx = pkg.Context(0)
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
    ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho)
    nfd = ctx.node_dofs(); fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
    pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
    ctx.apply_dirichlet(pres)
    x = rng.standard_normal(ctx.ndofs)
    ya = ctx.spmv(x); ym = ctx.spmv(x, matrix_free=True)
    scale = np.abs(ya).max()
    assert np.max(np.abs(ya - ym)) <= 1e-12 * scale, np.max(np.abs(ya - ym)) / scale
    ctx.add_nodal_force(pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0.0, 0.0, -1.0])
    ctx.apply_dirichlet(pres)
    sa = ctx.solve_pcg(1e-10, 1e-10, 100000); ua = ctx.solution()
    sm = ctx.solve_pcg(1e-10, 1e-10, 100000, matrix_free=True); um = ctx.solution()
    assert sa["converged"] and sm["converged"] and np.linalg.norm(ua - um) <= 1e-8 * np.linalg.norm(ua)
    ctx.close()
print("EBE VARIANT OK")
'''


@pytest.mark.parametrize("gather", [False, True])
def test_matrix_free_forms_match_assembled(tmp_path, gather):
    script = tmp_path / "ebe_variant.py"
    script.write_text(SCRIPT % {"root": ROOT})
    env = dict(os.environ)
    if gather:
        env["TOE_EBE_GATHER"] = "1"
    r = subprocess.run([sys.executable, str(script)], cwd=ROOT, capture_output=True, text=True, timeout=280, env=env)
    assert r.returncode == 0 and "EBE VARIANT OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
