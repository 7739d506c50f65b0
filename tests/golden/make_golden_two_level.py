"""Iteration count of the two-level preconditioner (512 boxes = 32x8x2, rotations about the box centres) on the C3 (1M-tet)
cantilever, computed with the C oracle's K and a numpy/scipy restatement of the preconditioner (tests/two_level_checks.py has
the same restatement for small meshes).  Adds the entry "two_level" to fullsize_c3.json["C3_1M"].  A few minutes of CPU."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from oracle import c_oracle  # noqa: E402

pkg = graft.load_package()
dims, boxes, tol = (120, 50, 28), (32, 8, 2), 1e-8
pts, cells = pkg.meshgen.cantilever(*dims)
fixed = pkg.meshgen.nodes_at_plane(pts, 0, 0.0); load = pkg.meshgen.nodes_at_plane(pts, 0, 60.0)
lam, mu = pkg.create_material_model(1.0, 0.3)
t0 = time.time()
cp = c_oracle.CProblem(pts, cells)
cp.assemble(lam_mu=(lam, mu)); cp.apply_force(load, [0.0, 0.0, -1.0])
pres0 = (cp.node_first_dof[fixed - 1][:, None] + np.arange(3)[None, :]).reshape(-1)
cp.apply_dirichlet(pres0)
n = cp.n
K = sp.csc_matrix((cp.nzval, cp.rowval, cp.colptr), shape=(n, n)).tocsr()
b = cp.f.copy()
nfd = cp.node_first_dof                                   # 0-based first dof here
lo, hi = pts.min(0), pts.max(0); bx = np.array(boxes); h = (hi - lo) / bx
idx = np.clip(np.floor((pts - lo) / h).astype(int), 0, bx - 1)
agg = idx[:, 0] + bx[0] * (idx[:, 1] + bx[1] * idx[:, 2])
d = pts - (lo + (idx + 0.5) * h)
rows, cols, vals = [], [], []
for c in range(3):
    rows.append(nfd + c); cols.append(6 * agg + c); vals.append(np.ones(pts.shape[0]))
for comp, k, v in ((1, 0, -d[:, 2]), (2, 0, d[:, 1]), (0, 1, d[:, 2]), (2, 1, -d[:, 0]), (0, 2, -d[:, 1]), (1, 2, d[:, 0])):
    rows.append(nfd + comp); cols.append(6 * agg + 3 + k); vals.append(v)
Z = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, 6 * int(np.prod(bx))))
mask = np.ones(n); mask[pres0] = 0.0
Z = (sp.diags(mask) @ Z).tocsr()
Ac = (Z.T @ (K @ Z)).toarray()
Aci = np.linalg.pinv(0.5 * (Ac + Ac.T), rcond=1e-12, hermitian=True)
D = K.diagonal().copy(); D[np.abs(D) < 1e-12] = 1.0; Dinv = 1.0 / D
ZT = Z.T.tocsr()
M = lambda r: Dinv * r + Z @ (Aci @ (ZT @ r))
x = np.zeros(n); r = b.copy(); z = M(r); p = z.copy(); gam = r @ z; eps = tol + tol * np.sqrt(gam); k = 0
while np.sqrt(gam) > eps and k < 20000:
    Ap = K @ p; a = gam / (p @ Ap); x += a * p; r -= a * Ap; z = M(r); g2 = r @ z; p = z + (g2 / gam) * p; gam = g2; k += 1
energy = 0.5 * float(x @ (K @ x))
out = {"boxes": boxes, "coarse_dofs": int(Ac.shape[0]), "niter": int(k), "energy": energy, "seconds": time.time() - t0,
       "source": "C oracle K + numpy restatement of M^-1 = D^-1 + Z pinv(Z'KZ) Z', atol = rtol = 1e-8"}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_c3.json")
g = json.load(open(path)); g["C3_1M"]["two_level"] = out
json.dump(g, open(path, "w"), indent=1)
print(out)
