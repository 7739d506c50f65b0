"""Synthetic structured cantilever meshes (SURVEY.md §8(d), configs C3–C5).

Box [0,Lx]×[0,Ly]×[0,Lz] cut into nx×ny×nz cubes; node index = i + (nx+1)(j + (ny+1)k), i fastest
(1-based on output, like everything that crosses the reference's API).  Each cube is split into
6 tetrahedra sharing the diagonal c2–c8 of its VTK-hex-ordered corners — the split Ferrite's
`generate_grid(Tetrahedron, …)` uses, all with positive Jacobian.  `hex=True` keeps the cubes as
Hex8 cells in VTK order (the layout of the reference's SIMP fixture).
"""
from __future__ import annotations

import numpy as np

# cube corners (VTK hexahedron order, 1-based) of the 6 tets
_TETS = np.array([(1, 2, 4, 8), (1, 5, 2, 8), (2, 3, 4, 8), (2, 7, 3, 8), (2, 5, 6, 8), (2, 6, 7, 8)]) - 1
_CORNER = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)])

# named sizes (cubes per axis) of BASELINE.json's synthetic configs
SIZES = {"C3_1M": (120, 50, 28), "C4_10M": (260, 110, 58), "C5_60M": (480, 200, 104)}


def box_points(nx, ny, nz, L=(60.0, 20.0, 4.0)):
    x = np.linspace(0.0, L[0], nx + 1)
    y = np.linspace(0.0, L[1], ny + 1)
    z = np.linspace(0.0, L[2], nz + 1)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    return np.ascontiguousarray(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1))


def _cube_corners(nx, ny, nz):
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # cube order: i fastest, then j, then k
    i = i.transpose(2, 1, 0).ravel(); j = j.transpose(2, 1, 0).ravel(); k = k.transpose(2, 1, 0).ravel()
    c = np.empty((i.size, 8), dtype=np.int64)
    for a, (di, dj, dk) in enumerate(_CORNER):
        c[:, a] = (i + di) + (nx + 1) * ((j + dj) + (ny + 1) * (k + dk))
    return c


def cantilever(nx, ny, nz, L=(60.0, 20.0, 4.0), hex=False):
    """→ (points (nn,3) f64, cells (ne,npc) int64 1-based)."""
    pts = box_points(nx, ny, nz, L)
    corners = _cube_corners(nx, ny, nz)
    if hex:
        return pts, np.ascontiguousarray(corners + 1)
    cells = corners[:, _TETS].reshape(-1, 4)
    return pts, np.ascontiguousarray(cells + 1)


def nodes_at_plane(points, axis, value, tol=1e-6):
    """1-based node ids with |x[axis]-value| < tol — `nodes_at_plane` of test/runtests.jl:10-18."""
    return np.nonzero(np.abs(points[:, axis] - value) < tol)[0].astype(np.int64) + 1


def simp_like_density(ne, seed=12345, lo=0.05, hi=1.0):
    """ρ_e ~ U[lo,hi], drawn in element order."""
    return np.random.default_rng(seed).uniform(lo, hi, size=ne)
