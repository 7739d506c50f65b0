"""N-GPU diagnostic of the partitioned solve (round-1 open item: a partitioned solve at 10M tets sometimes ended in a CG breakdown).
    torchrun --nproc-per-node 2 tools/dist_diag.py 260,110,58 [reps] [solves_per_setup]
Every rep re-sets-up the same mesh on the same ctx; per set-up it runs `solves_per_setup` full steps (assemble, load, constrain, solve)
and one more solve without re-assembly.  EVERY rank prints, per solve: iteration count, flags, fingerprints of the state the solve
starts from (diagonal, f, K·x for two fixed x — must be identical in every rep) and the first iteration at which the residual history
leaves the first converged history seen — so a wrong set-up (fingerprints differ / history differs from iteration 0) is told apart from
a transient fault in the iteration (history leaves the reference somewhere in the middle) and from a deterministic failure."""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
os.environ.setdefault("TOE_DIST_NO_RETRY", "1")
import __graft_entry__ as graft  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

pkg = graft.load_package()
rank, local_rank, world = pkg.parallel.env_rank()
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dims = tuple(int(x) for x in sys.argv[1].split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
per_setup = int(sys.argv[3]) if len(sys.argv) > 3 else 2
itmax = int(os.environ.get("DIAG_ITMAX", "20000"))
pts, cells = pkg.meshgen.cantilever(*dims)
fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
lam, mu = pkg.create_material_model(1.0, 0.3)
ctx = pkg.parallel.create_distributed_context(dist, local_rank) if world > 1 else pkg.Context(local_rank)
pres = None
x1 = x2 = None
ref_hist = None
ref_trace = None
TRACE = os.environ.get("TOE_CG_TRACE") == "1" and world > 1


def fp(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]


def report(tag, st, t0, state):
    global ref_hist
    h = np.asarray(st["residuals"])
    div = ""
    if ref_hist is None and st["converged"]:
        ref_hist = h.copy()
    if ref_hist is not None:
        m = min(len(h), len(ref_hist))
        d = np.nonzero(h[:m] != ref_hist[:m])[0]
        div = " leaves_ref_at %s" % (int(d[0]) if len(d) else ("never" if len(h) == len(ref_hist) else "len %d/%d" % (len(h), len(ref_hist))))
    if TRACE:
        global ref_trace
        tr = ctx.cg_trace(st["niter"] + 1)
        if ref_trace is None and st["converged"]:
            ref_trace = tr.copy()
        elif ref_trace is not None:
            m = min(len(tr), len(ref_trace))
            bad = np.nonzero(np.any(tr[:m] != ref_trace[:m], axis=1))[0]
            if len(bad):
                j = int(bad[0])
                cols = ["gamma_sum", "delta_sum", "gamma_partial", "delta_partial"]
                np.set_printoptions(precision=17)
                print("[r%d] TRACE first mismatch at iteration %d, columns %s" % (rank, j, [cols[k] for k in range(4) if tr[j, k] != ref_trace[j, k]]), flush=True)
                for jj in range(max(0, j - 1), min(m, j + 3)):
                    print("[r%d]   it %d got %s" % (rank, jj, np.array2string(tr[jj], floatmode="unique")), flush=True)
                    print("[r%d]   it %d ref %s" % (rank, jj, np.array2string(ref_trace[jj], floatmode="unique")), flush=True)
                nb = np.nonzero(np.any(tr[:m, 2:] != ref_trace[:m, 2:], axis=1))[0]
                print("[r%d]   first mismatching LOCAL partial at iteration %s" % (rank, int(nb[0]) if len(nb) else None), flush=True)
    e, c, _ = ctx.energy()
    print("[r%d] %s niter %d conv %d brk %d rst %d solve_s %.3f relres %.3e energy %.10f hist %s%s%s wall %.2f" % (
        rank, tag, st["niter"], st["converged"], st["breakdown"], st.get("restarts", 0), st["solve_seconds"], st["rel_res_l2"], e, fp(h), div, state,
        time.perf_counter() - t0), flush=True)


for rep in range(reps):
    t0 = time.perf_counter()
    ctx.set_mesh(pts, cells, distributed=world > 1); ctx.build_dofs(); ctx.build_pattern()
    if pres is None:
        nfd = ctx.node_dofs(); pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
        x1 = np.cos(0.001 * np.arange(ctx.ndofs)); x1[pres - 1] = 0.0
        x2 = np.random.default_rng(7).standard_normal(ctx.ndofs); x2[pres - 1] = 0.0
    for k in range(per_setup):
        ctx.assemble_lame(lam, mu); ctx.add_nodal_force(load, [0, 0, -1.0]); ctx.apply_dirichlet(pres)
        state = ""
        if k == 0:
            state = " diag %s f %s Kx1 %s Kx2 %s" % (fp(ctx.diagonal()), fp(ctx.rhs()), fp(ctx.spmv(x1)), fp(ctx.spmv(x2)))
        st = ctx.solve_pcg(1e-8, 1e-8, itmax, history=True)
        report("rep %d step %d" % (rep, k), st, t0, state)
    st = ctx.solve_pcg(1e-8, 1e-8, itmax, history=True)
    report("rep %d resolve" % rep, st, t0, " Kx1 %s" % fp(ctx.spmv(x1)))
soak = int(os.environ.get("DIAG_SOAK", "0"))
for what in ((1, 2, 3) if world > 1 else (1,)) if soak else ():
    t0 = time.perf_counter()
    bad_batches, bad_entries = ctx.spmv_soak(soak, what)
    print("[r%d] operator soak (%s): %d applications, %d batches of 256 with a mismatch, %d mismatching entries, %.1f s" % (
        rank, {1: "local product", 2: "exchange", 3: "product + exchange"}[what], soak, bad_batches, bad_entries, time.perf_counter() - t0), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
