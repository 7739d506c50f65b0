"""Stage-by-stage timing probe on one GPU (development aid; bench.py is the contract)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as graft

pkg = graft.load_package()


def run(dims, itmax=30000, matrix_free=False, variant=0, simp=False, solve=True):
    t0 = time.perf_counter()
    pts, cells = pkg.meshgen.cantilever(*dims)
    tgen = time.perf_counter() - t0
    ctx = pkg.Context(0)
    t0 = time.perf_counter()
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
    tsetup = time.perf_counter() - t0
    lam, mu = pkg.create_material_model(1.0, 0.3)
    out = {"dims": dims, "ne": ctx.ne, "ndofs": ctx.ndofs, "nnz": ctx.nnz, "gen_s": tgen, "setup_wall_s": tsetup}
    rho = pkg.meshgen.simp_like_density(ctx.ne) if simp else None
    for rep in range(2):
        t0 = time.perf_counter()
        if matrix_free:
            ctx.set_material_simp(1.0, 0.3, 1e-8, 3.0, rho) if simp else ctx.set_material_lame(lam, mu)
        else:
            ctx.assemble_simp(1.0, 0.3, 1e-8, 3.0, rho, variant) if simp else ctx.assemble_lame(lam, mu, variant)
        out["assemble_wall_s_%d" % rep] = time.perf_counter() - t0
    load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
    fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
    ctx.add_nodal_force(load, [0, 0, -1.0])
    nfd = ctx.node_dofs()
    pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
    ctx.apply_dirichlet(pres)
    for mf in ([False, True] if not matrix_free else [True]):
        s, b = ctx.time_spmv(matrix_free=mf, reps=20)
        out["spmv_%s" % ("ebe" if mf else "bsr")] = {"ms": s * 1e3, "GBs": b / s / 1e9, "bytes": b}
    if solve:
        t0 = time.perf_counter()
        st = ctx.solve_pcg(1e-8, 1e-8, itmax, matrix_free=matrix_free)
        out["solve_wall_s"] = time.perf_counter() - t0
        out["pcg"] = st
        e, c, _ = ctx.energy()
        out["energy"] = e; out["compliance"] = c
    out["timings"] = ctx.timings()
    ctx.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "1M"
    dims = {"small": (48, 16, 6), "200k": (96, 32, 12), "1M": (120, 50, 28), "10M": (260, 110, 58)}[which]
    mf = "mf" in sys.argv[2:]
    var = 1 if "atomic" in sys.argv[2:] else 0
    simp = "simp" in sys.argv[2:]
    nosolve = "nosolve" in sys.argv[2:]
    r = run(dims, matrix_free=mf, variant=var, simp=simp, solve=not nosolve)
    print(json.dumps(r, indent=1, default=float))
