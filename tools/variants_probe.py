"""One-GPU measurement of every opt-in variant next to its default (development aid for the first GPU call of a round):
    python tools/variants_probe.py 10M            # sizes: small 200k 1M 10M
Prints one JSON object: assembly GATHER vs ROWS (ms, values agree), matrix-free tile vs pipelined operator (ms, bit-identical),
Jacobi vs two-level PCG (iterations, seconds, energies)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as graft

pkg = graft.load_package()
A = pkg._lib


def main():
    argv = [a for a in sys.argv[1:]]
    only = None                                    # --only asm|ebe|pc: one section (bench.py runs each in its own child process)
    if "--only" in argv:
        i = argv.index("--only"); only = argv[i + 1]; del argv[i:i + 2]
    which = argv[0] if argv else "1M"
    dims = {"toy": (8, 3, 2), "small": (48, 16, 6), "200k": (96, 32, 12), "1M": (120, 50, 28), "10M": (260, 110, 58)}[which]
    pts, cells = pkg.meshgen.cantilever(*dims)
    ctx = pkg.Context(0)
    ctx.set_mesh(pts, cells); ctx.build_dofs(); ctx.build_pattern()
    lam, mu = pkg.create_material_model(1.0, 0.3)
    out = {"dims": dims, "ne": int(ctx.ne), "ndofs": int(ctx.ndofs)}
    # ---- assembly variants ------------------------------------------------------------------------------------------
    asm = {}
    ref_diag = None
    for name, var in (("gather", A.ASM_GATHER), ("rows", A.ASM_ROWS)) if only in (None, "asm") else ():
        ts = []
        for _ in range(4):
            ctx.assemble_lame(lam, mu, var)
            ts.append(ctx.timings()["assemble"] * 1e3)
        d = ctx.diagonal()
        x = np.random.default_rng(0).standard_normal(ctx.ndofs)
        y = ctx.spmv(x)
        if ref_diag is None:
            ref_diag, ref_y = d, y
        asm[name] = {"ms_min": min(ts[1:]), "ms_all": ts, "elements_per_s": ctx.ne / (min(ts[1:]) * 1e-3),
                     "max_rel_diff_diag_vs_gather": float(np.max(np.abs(d - ref_diag) / np.abs(ref_diag))),
                     "max_rel_diff_Kx_vs_gather": float(np.max(np.abs(y - ref_y)) / np.max(np.abs(ref_y)))}
    if asm:
        out["assembly"] = asm
    # ---- matrix-free operator variants ----------------------------------------------------------------------------
    ctx.assemble_lame(lam, mu)
    # bit-identity of the pipelined operator is checked BEFORE constraints exist: with prescribed DOFs toe_spmv masks the input
    # columns, and the masking form always takes the tile kernel (the pipelined kernel serves the PCG loop, whose vectors are masked)
    pipe_identical = None
    if only in (None, "ebe"):
        xf = np.random.default_rng(2).standard_normal(ctx.ndofs)
        os.environ.pop("TOE_EBE_PIPE", None)
        y0 = ctx.spmv(xf, matrix_free=True)
        os.environ["TOE_EBE_PIPE"] = "1"
        y1 = ctx.spmv(xf, matrix_free=True)
        os.environ.pop("TOE_EBE_PIPE", None)
        pipe_identical = bool(np.array_equal(y0, y1))
        del xf, y0, y1
    load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0); fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
    ctx.add_nodal_force(load, [0, 0, -1.0])
    nfd = ctx.node_dofs()
    pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
    ctx.apply_dirichlet(pres)
    x = np.random.default_rng(1).standard_normal(ctx.ndofs); x[pres - 1] = 0.0
    ebe = {}
    y_tile = None
    for name, env in (("tile", None), ("pipe", "1")) if only in (None, "ebe") else ():
        if env:
            os.environ["TOE_EBE_PIPE"] = env
        else:
            os.environ.pop("TOE_EBE_PIPE", None)
        s, b = ctx.time_spmv(matrix_free=True, reps=20)
        ebe[name] = {"ms": s * 1e3, "GBs_algorithmic": b / s / 1e9, "bit_identical_to_tile": True if name == "tile" else pipe_identical}
    os.environ.pop("TOE_EBE_PIPE", None)
    if ebe:
        s, b = ctx.time_spmv(matrix_free=False, reps=20)
        ebe["assembled_spmv_ms"] = s * 1e3
        out["matrix_free_operator"] = ebe
    # ---- preconditioners ----------------------------------------------------------------------------------------------
    pc = {}
    if only not in (None, "pc"):
        ctx.close()
        print(json.dumps(out, default=float))
        return
    for name, tl in (("jacobi", False), ("two_level", True)):
        t0 = time.perf_counter()
        st = ctx.solve_pcg(1e-8, 1e-8, 40000, two_level=tl)
        wall = time.perf_counter() - t0
        e, c, _ = ctx.energy()
        pc[name] = {"iterations": int(st["niter"]), "converged": bool(st["converged"]), "solve_seconds": st["solve_seconds"], "wall_seconds": wall,
                    "coarse_dofs": int(st["coarse_dofs"]), "coarse_operator_seconds": st["precond_seconds"], "rel_res_l2": st["rel_res_l2"], "energy": e}
    st = ctx.solve_pcg(1e-8, 1e-8, 40000, two_level=True)          # second two-level solve: coarse operator cached, graph captured
    pc["two_level_repeat"] = {"iterations": int(st["niter"]), "solve_seconds": st["solve_seconds"], "coarse_operator_seconds": st["precond_seconds"]}
    out["preconditioner"] = pc
    ctx.close()
    print(json.dumps(out, indent=1, default=float))


if __name__ == "__main__":
    main()
