"""Freezes the C oracle's answer for the C3 (1M-tet) synthetic cantilever of BASELINE.json: energy, compliance, PCG
iteration count at atol=rtol=1e-8 (Krylov.jl criterion).  ~10 minutes of single-core CPU.  Output: fullsize_c3.json.
The 10M-tet values in the same file come from the round-1 B200 runs (N=1, 2 and 8 GPUs agree to 9 digits) and are a
regression anchor, not an oracle result: the CPU oracle needs hours at that size."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from oracle import c_oracle  # noqa: E402

pkg = graft.load_package()
dims = (120, 50, 28)
pts, cells = pkg.meshgen.cantilever(*dims)
fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0)
load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
lam, mu = pkg.create_material_model(1.0, 0.3)
t0 = time.time()
cp = c_oracle.CProblem(pts, cells)
cp.assemble(lam_mu=(lam, mu))
cp.apply_force(load, [0.0, 0.0, -1.0])
pres0 = (cp.node_first_dof[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1)
m = cp.apply_dirichlet(pres0)


def mean_abs_diag_fsum():
    """mean(abs(diag K)) summed exactly (math.fsum): the C oracle adds 536 877 terms in a naive running sum and lands 2.3e-12 off, which
    is above the 1e-12 bar the GPU test holds this value to (round-1 verdict).  diag(K) from the per-cell closed form."""
    import math
    X = pts[cells - 1]
    J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=1)
    g123 = np.transpose(np.linalg.inv(J), (0, 2, 1))
    G = np.concatenate([-g123.sum(axis=1, keepdims=True), g123], axis=1)
    d = (np.linalg.det(J) / 6)[:, None, None] * (lam * G * G + mu * ((G * G).sum(axis=2, keepdims=True) + G * G))
    idx = (3 * (cells - 1))[:, :, None] + np.arange(3)[None, None, :]
    diag = np.bincount(idx.reshape(-1), weights=d.reshape(-1), minlength=3 * pts.shape[0])
    return math.fsum(np.abs(diag)) / diag.size


m_naive, m = m, mean_abs_diag_fsum()
assert abs(m - m_naive) <= 1e-11 * m
u, niter, solved, _ = cp.pcg(1e-8, 40000)
e = cp.energy(u)
out = {"C3_1M": {"dims": dims, "ne": int(cp.ne), "ndofs": int(cp.n), "nnz": int(cp.nnz), "mean_diag": float(m), "niter": int(niter), "solved": bool(solved),
                 "energy": float(e), "compliance": float(cp.f @ u), "max_abs_u": float(np.abs(u).max()), "seconds": time.time() - t0,
                 "source": "oracle/oracle.c (C restatement), Jacobi-PCG atol=rtol=1e-8"},
       "C4_10M": {"dims": (260, 110, 58), "ne": 9952800, "ndofs": 5127867, "nnz": 227103345, "niter": 13689, "energy": 328.5807407, "compliance": 657.1614464,
                  "source": "round-1 B200 runs of this library (1, 2 and 8 GPUs agree to 9 digits): regression anchor, not an oracle value"}}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_c3.json"), "w"), indent=1)
print(out)
