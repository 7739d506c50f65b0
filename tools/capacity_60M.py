"""BASELINE config 5 capacity check: 60M-tet cantilever on one B200, assembled and matrix-free operators, fixed iteration budget."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
pkg = graft.load_package()
dims = (480, 200, 104)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
t0 = time.perf_counter(); pts, cells = pkg.meshgen.cantilever(*dims); tgen = time.perf_counter() - t0
ctx = pkg.Context(0)
t0 = time.perf_counter(); ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern(); tsetup = time.perf_counter() - t0
lam, mu = pkg.create_material_model(1.0, 0.3)
out = {"ne": ctx.ne, "ndofs": ctx.ndofs, "nnz": ctx.nnz, "gen_s": tgen, "setup_s": tsetup}
load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0); fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
nfd = ctx.node_dofs(); pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
for mf in (True, False):
    t0 = time.perf_counter()
    ctx.set_material_lame(lam, mu) if mf else ctx.assemble_lame(lam, mu)
    out["assemble_s_mf%d" % mf] = time.perf_counter() - t0
    ctx.add_nodal_force(load, [0, 0, -1.0]); ctx.apply_dirichlet(pres)
    s, b = ctx.time_spmv(matrix_free=mf, reps=5)
    st = ctx.solve_pcg(1e-8, 1e-8, iters, matrix_free=mf)
    e, c, _ = ctx.energy()
    out["mf%d" % mf] = {"op_ms": s * 1e3, "op_GBs": b / s / 1e9, "iters": st["niter"], "solve_s": st["solve_seconds"], "res_M": st["res_M"], "energy_partial": e}
out["timings"] = ctx.timings()
print(json.dumps(out, indent=1, default=float))
