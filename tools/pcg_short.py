"""Short PCG run (fixed iteration budget) for ncu captures: python tools/pcg_short.py 10M 20 [mf]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
pkg = graft.load_package()
which = sys.argv[1]; iters = int(sys.argv[2]); mf = "mf" in sys.argv[3:]
dims = {"small": (48, 16, 6), "200k": (96, 32, 12), "1M": (120, 50, 28), "10M": (260, 110, 58)}[which]
pts, cells = pkg.meshgen.cantilever(*dims)
ctx = pkg.Context(0)
ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
lam, mu = pkg.create_material_model(1.0, 0.3)
ctx.set_material_lame(lam, mu) if mf else ctx.assemble_lame(lam, mu)
ctx.add_nodal_force(pkg.meshgen.nodes_at_plane(pts, 0, 60.0), [0, 0, -1.0])
nfd = ctx.node_dofs(); fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
ctx.apply_dirichlet(np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1)))
st = ctx.solve_pcg(1e-8, 1e-8, iters, matrix_free=mf, graph=False)
print(st)
print(ctx.energy()[:2])
