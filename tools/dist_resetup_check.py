"""N-GPU diagnostic: solve, re-setup the same mesh on the same ctx, solve again; print both outcomes (run under torchrun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
import torch, torch.distributed as dist
pkg = graft.load_package()
rank, local_rank, world = pkg.parallel.env_rank()
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dims = tuple(int(x) for x in sys.argv[1].split(","))
pts, cells = pkg.meshgen.cantilever(*dims)
fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
lam, mu = pkg.create_material_model(1.0, 0.3)
ctx = pkg.parallel.create_distributed_context(dist, local_rank)
pres = None
for rep in range(3):
    t0 = time.perf_counter()
    ctx.set_mesh(pts, cells, distributed=True); ctx.build_dofs(); ctx.build_pattern()
    if pres is None:
        nfd = ctx.node_dofs(); pres = np.sort((nfd[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1))
    ctx.assemble_lame(lam, mu); ctx.add_nodal_force(load, [0, 0, -1.0]); ctx.apply_dirichlet(pres)
    st = ctx.solve_pcg(1e-8, 1e-8, 40000)
    e, c, _ = ctx.energy()
    if rank == 0:
        print("rep", rep, "wall %.3f" % (time.perf_counter() - t0), "niter", st["niter"], "conv", st["converged"], "brk", st["breakdown"],
              "solve_s %.3f" % st["solve_seconds"], "energy %.10f" % e, ctx.comm_info()["transport"], flush=True)
dist.destroy_process_group()
