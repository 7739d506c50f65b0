"""N-GPU diagnostic for the open item of DESIGN.md §6 (first partitioned solve after a set-up breaking down at 10M tets):
    torchrun --nproc-per-node 2 tools/dist_resetup_check.py 260,110,58 [reps] [probe]
Every rep re-sets-up the same mesh on the same ctx and solves with the restart net OFF (TOE_DIST_NO_RETRY=1), so a breakdown is
visible; with `probe` the state the solve starts from is fingerprinted first (diagonal, f, K·x of a fixed x — all three must be
identical in every rep, so a corrupted set-up is told apart from a solve that goes wrong on good data).  A/B switches:
TOE_DIST_NO_ALIGN=1 (no stream rendezvous before the first exchange), TOE_DIST_P2P=1 (peer-memory transport)."""
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
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dims = tuple(int(x) for x in sys.argv[1].split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
probe = "probe" in sys.argv[2:]
pts, cells = pkg.meshgen.cantilever(*dims)
fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
lam, mu = pkg.create_material_model(1.0, 0.3)
ctx = pkg.parallel.create_distributed_context(dist, local_rank)
pres = None
x = None


def fp(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


for rep in range(reps):
    t0 = time.perf_counter()
    ctx.set_mesh(pts, cells, distributed=True); ctx.build_dofs(); ctx.build_pattern()
    if pres is None:
        nfd = ctx.node_dofs(); pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
        x = np.cos(0.001 * np.arange(ctx.ndofs)); x[pres - 1] = 0.0
    ctx.assemble_lame(lam, mu); ctx.add_nodal_force(load, [0, 0, -1.0]); ctx.apply_dirichlet(pres)
    state = ""
    if probe:
        state = " diag %s f %s Kx %s" % (fp(ctx.diagonal()), fp(ctx.rhs()), fp(ctx.spmv(x)))
    st = ctx.solve_pcg(1e-8, 1e-8, 40000, history=True)
    e, c, _ = ctx.energy()
    hist = st.get("residuals")
    tail = ""
    if hist is not None and len(hist) > 1:
        h = np.asarray(hist)
        k = int(np.argmin(h))
        tail = " res[0] %.3e min %.3e at %d last %.3e" % (h[0], h[k], k, h[-1])
    if rank == 0:
        print("rep", rep, "wall %.3f" % (time.perf_counter() - t0), "niter", st["niter"], "conv", st["converged"], "brk", st["breakdown"],
              "solve_s %.3f" % st["solve_seconds"], "energy %.10f" % e, ctx.comm_info()["transport"] + state + tail, flush=True)
dist.destroy_process_group()
