"""Multi-GPU parity worker: run under torchrun with N ranks (one per GPU).  Every rank solves its part of the same
global problem; rank 0 also solves it on a single, unpartitioned ctx and compares (partition invariance)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    pkg = graft.load_package()
    rank, local_rank, world = pkg.parallel.env_rank()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "48,16,6").split(","))
    simp = "simp" in sys.argv[2:]
    pts, cells = pkg.meshgen.cantilever(*dims)
    rho = pkg.meshgen.simp_like_density(cells.shape[0]) if simp else None
    fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
    load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
    lam, mu = pkg.create_material_model(1.0, 0.3)
    results = {}

    verbose = os.environ.get("TOE_DIST_VERBOSE") == "1"
    use_graph = os.environ.get("TOE_DIST_NOGRAPH") != "1"

    def say(*a):
        if verbose:
            print("[rank %d]" % rank, *a, flush=True)

    def run(ctx, distributed, mf):
        say("set_mesh", distributed, mf)
        ctx.set_mesh(pts, cells, distributed=distributed)
        say("build")
        ctx.build_dofs(); ctx.build_pattern()
        say("material")
        if simp:
            (ctx.set_material_simp if mf else ctx.assemble_simp)(1.0, 0.3, 1e-8, 3.0, rho)
        else:
            (ctx.set_material_lame if mf else ctx.assemble_lame)(lam, mu)
        say("loads")
        ctx.add_nodal_force(load, [0.0, 0.0, -1.0])
        if simp:
            ctx.add_volume_force([0.0, 0.0, -0.01], density=rho, skip_below=1e-6)
        nfd = ctx.node_dofs()
        pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
        say("dirichlet")
        m = ctx.apply_dirichlet(pres)
        say("solve")
        st = ctx.solve_pcg(1e-10, 1e-10, 200000, matrix_free=mf, graph=use_graph)
        say("solved", st["niter"], st["converged"])
        u = ctx.solution()
        e, c, ee = ctx.energy(per_element=True)
        _, vm, mx, arg = ctx.stresses(False, True)
        return dict(u=u, e=e, c=c, ee=ee, it=st["niter"], conv=st["converged"], restarts=st.get("restarts", 0), m=m, f=ctx.rhs(), nfd=nfd, mx=mx, arg=arg, vm=vm,
                    diag=ctx.diagonal())

    ctx = pkg.parallel.create_distributed_context(dist, local_rank)
    repeat = 3 if "repeat" in sys.argv[2:] else 1
    for mf in (False, True):
        for rep in range(repeat):
            r = run(ctx, True, mf)
            if rep and not (np.array_equal(r["u"], results[("dist", mf)]["u"]) and r["it"] == results[("dist", mf)]["it"]):
                print("[rank %d] NON-REPRODUCIBLE partitioned solve (mf=%d, repeat %d): iters %d vs %d" % (rank, mf, rep, r["it"], results[("dist", mf)]["it"]), flush=True)
                r["conv"] = 0
            results[("dist", mf)] = r
    info = ctx.comm_info()
    part = ctx.partition()
    sizes = ctx.local_sizes()
    x = np.random.default_rng(5).standard_normal(ctx.ndofs)
    y_dist = ctx.spmv(x, matrix_free=True)
    ctx.close()
    ok = True
    if rank == 0:
        single = pkg.Context(local_rank)
        ref = run(single, False, False)
        y_single = single.spmv(x)
        single.close()
        counts = np.bincount(part, minlength=world)
        print("partition sizes", counts.tolist(), "local sizes rank0", sizes, "transport", info["transport"])
        want = os.environ.get("TOE_EXPECT_TRANSPORT")
        assert want is None or info["transport"] == want, info
        assert counts.max() - counts.min() <= 1, counts
        for mf in (False, True):
            r = results[("dist", mf)]
            rel_u = np.linalg.norm(r["u"] - ref["u"]) / np.linalg.norm(ref["u"])
            rel_e = abs(r["e"] - ref["e"]) / abs(ref["e"])
            rel_c = abs(r["c"] - ref["c"]) / abs(ref["c"])
            rel_ee = np.max(np.abs(r["ee"] - ref["ee"])) / np.max(np.abs(ref["ee"]))
            rel_f = np.max(np.abs(r["f"] - ref["f"])) / np.max(np.abs(ref["f"]))
            rel_d = np.max(np.abs(r["diag"] - ref["diag"]) / np.abs(ref["diag"]))
            print("mf=%d iters %d (single %d) rel_u %.2e rel_e %.2e rel_c %.2e rel_ee %.2e rel_f %.2e rel_diag %.2e m %.12g/%.12g argmax %d/%d restarts %d"
                  % (mf, r["it"], ref["it"], rel_u, rel_e, rel_c, rel_ee, rel_f, rel_d, r["m"], ref["m"], r["arg"], ref["arg"], r["restarts"]))
            good = (r["conv"] == 1 and rel_u < 1e-8 and rel_e < 1e-8 and rel_c < 1e-8 and rel_ee < 1e-8 and rel_f < 1e-12 and rel_d < 1e-12
                    and abs(r["m"] - ref["m"]) < 1e-12 * ref["m"] and np.array_equal(r["nfd"], ref["nfd"])
                    and abs(r["it"] - ref["it"]) <= max(5, ref["it"] // 50) and r["arg"] == ref["arg"]
                    and np.max(np.abs(r["vm"] - ref["vm"])) <= 1e-7 * ref["mx"])
            ok = ok and good
        rel_y = np.max(np.abs(y_dist - y_single)) / np.max(np.abs(y_single))
        print("spmv rel", rel_y)
        ok = ok and rel_y < 1e-12
        print("DIST PARITY", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
