"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

numpy/scipy restatement of the reference's strain-energy evaluation path
(jezekon/TopOptEval.jl): setup → assembly → loads → Dirichlet → solve → energy.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker.

PARITY UNPINNED: the reference is pure Julia (not installed in this image), its arithmetic lives in
un-vendored packages (Ferrite 1.0, Tensors 1.16, Krylov 0.10, SparseArrays — `Project.toml:19-27`)
and its own tests assert no numeric result for this path (`test/runtests.jl:26,43-45,83-85` check only
`>0` / `isfinite`).  So this oracle is anchored on (i) the reference's call sites, cited per function,
(ii) the published algorithms of those packages, (iii) the reference's own
recipes (`test/runtests.jl:21-89`) and its one analytic check (`test/VolumeForces/testVolumeForces.jl:8-37,159-168`:
cantilever tip deflection ρgL⁴/(8EI) within 10 %), which `tests/test_oracle.py` reproduces.

Index convention: everything returned is **1-based** like the Julia objects it restates
(`dh.cell_dofs`, `K.colptr`, `K.rowval`, `ch.prescribed_dofs`).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ----------------------------------------------------------------------------------------------
# material models
# ----------------------------------------------------------------------------------------------

def create_material_model(youngs_modulus: float, poissons_ratio: float):
    """FiniteElementAnalysis.jl:103-109."""
    lam = youngs_modulus * poissons_ratio / ((1 + poissons_ratio) * (1 - 2 * poissons_ratio))
    mu = youngs_modulus / (2 * (1 + poissons_ratio))
    return lam, mu


def create_simp_material_model(E0: float, nu: float, Emin: float = 1e-6, p: float = 1.0):
    """FiniteElementAnalysis.jl:616-634 — note the *code* defaults (1e-6, 1.0), not the docstring's."""
    def material_for_density(density):
        E = Emin + (E0 - Emin) * density ** p                      # :624
        lam = E * nu / ((1 + nu) * (1 - 2 * nu))                   # :627
        mu = E / (2 * (1 + nu))                                    # :628
        return lam, mu
    return material_for_density


# ----------------------------------------------------------------------------------------------
# reference element data  (Ferrite Lagrange{Ref*,1}, QuadratureRule{Ref*}(2) — call site :162-171)
# ----------------------------------------------------------------------------------------------

_HEX_SIGNS = np.array([(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1),
                       (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)], dtype=np.float64)


def reference_element(npc: int):
    """→ (N (nq,npc), dNdxi (nq,npc,3), w (nq,)) for Tet4 (npc=4) or Hex8 (npc=8)."""
    if npc == 4:
        a = (5.0 - np.sqrt(5.0)) / 20.0
        b = (5.0 + 3.0 * np.sqrt(5.0)) / 20.0
        qp = np.array([(a, a, a), (a, a, b), (a, b, a), (b, a, a)])
        w = np.full(4, 1.0 / 24.0)
        N = np.stack([1 - qp[:, 0] - qp[:, 1] - qp[:, 2], qp[:, 0], qp[:, 1], qp[:, 2]], axis=1)
        d = np.array([(-1.0, -1.0, -1.0), (1.0, 0, 0), (0, 1.0, 0), (0, 0, 1.0)])
        dN = np.broadcast_to(d, (4, 4, 3)).copy()
        return N, dN, w
    if npc == 8:
        g = 1.0 / np.sqrt(3.0)
        qp = np.array([(sx * g, sy * g, sz * g) for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)])
        w = np.ones(8)
        s = _HEX_SIGNS
        N = 0.125 * (1 + qp[:, None, 0] * s[None, :, 0]) * (1 + qp[:, None, 1] * s[None, :, 1]) * (1 + qp[:, None, 2] * s[None, :, 2])
        dN = np.empty((8, 8, 3))
        dN[:, :, 0] = 0.125 * s[None, :, 0] * (1 + qp[:, None, 1] * s[None, :, 1]) * (1 + qp[:, None, 2] * s[None, :, 2])
        dN[:, :, 1] = 0.125 * s[None, :, 1] * (1 + qp[:, None, 0] * s[None, :, 0]) * (1 + qp[:, None, 2] * s[None, :, 2])
        dN[:, :, 2] = 0.125 * s[None, :, 2] * (1 + qp[:, None, 0] * s[None, :, 0]) * (1 + qp[:, None, 1] * s[None, :, 1])
        return N, dN, w
    raise ValueError("nodes per cell must be 4 (Tet4) or 8 (Hex8), got %d" % npc)


def cell_geometry(points, cells, q):
    """Ferrite `reinit!` (call sites FiniteElementAnalysis.jl:215,665; VolumeForce.jl:44,205) for
    quadrature point q of every cell: → (dNdx (ne,npc,3), detJdV (ne,)).  J = Σ_a x_a ⊗ ∂N_a/∂ξ,
    error if det J ≤ 0, ∇N_a = ∂N_a/∂ξ · J⁻¹, detJdV = det J · w_q."""
    npc = cells.shape[1]
    _, dN, w = reference_element(npc)
    X = points[cells - 1]                                          # (ne, npc, 3)
    J = np.einsum("eai,aj->eij", X, dN[q])                         # J[i,j] = Σ_a x_a[i] dN_a/dξ_j
    detJ = np.linalg.det(J)
    if np.any(detJ <= 0):
        bad = int(np.argmax(detJ <= 0)) + 1
        raise ValueError("det(J) is not positive: det(J) = %g in cell %d" % (detJ[bad - 1], bad))
    Jinv = np.linalg.inv(J)
    dNdx = np.einsum("aj,eji->eai", dN[q], Jinv)
    return dNdx, detJ * w[q]


# ----------------------------------------------------------------------------------------------
# setup_problem  (FiniteElementAnalysis.jl:151-185)
# ----------------------------------------------------------------------------------------------

def first_touch_dofs(cells, nn):
    """Ferrite `close!(dh)` (call site :174-176): walk cells 1..ne, vertices in
    the cell's own order; the first visit of a node hands it the next 3 DOFs.
    → (node_first_dof (nn,) 1-based, 0 = node in no cell; cell_dofs (ne,3·npc) 1-based; ndofs)."""
    ne, npc = cells.shape
    flat = (cells - 1).reshape(-1)
    _, first_pos = np.unique(flat, return_index=True)              # first position of every referenced node
    order = np.sort(first_pos)                                     # positions in walking order
    node_first_dof = np.zeros(nn, dtype=np.int64)
    node_first_dof[flat[order]] = 3 * np.arange(order.size, dtype=np.int64) + 1
    base = node_first_dof[cells - 1]                               # (ne, npc)
    cell_dofs = (base[:, :, None] + np.arange(3)[None, None, :]).reshape(ne, 3 * npc)
    return node_first_dof, cell_dofs, 3 * order.size


def first_touch_dofs_literal(cells, nn):
    """Same as `first_touch_dofs`, written as the sequential loop of Appendix A1 (small meshes)."""
    first = np.zeros(nn, dtype=np.int64)
    nxt = 1
    cd = []
    for cell in cells:
        row = []
        for g in cell:
            if first[g - 1] == 0:
                first[g - 1] = nxt
                nxt += 3
            row += [first[g - 1], first[g - 1] + 1, first[g - 1] + 2]
        cd.append(row)
    return first, np.array(cd, dtype=np.int64), nxt - 1


def sparsity_pattern(cell_dofs, ndofs):
    """Ferrite `allocate_matrix(dh)` (call site :181; Appendix A2): CSC, entry (i,j) stored iff some
    cell holds both DOFs, row indices ascending per column.  → (colptr (n+1,), rowval (nnz,)), 1-based."""
    ne, nb = cell_dofs.shape
    rows = np.repeat(cell_dofs - 1, nb, axis=1).reshape(-1)
    cols = np.tile(cell_dofs - 1, (1, nb)).reshape(-1)
    key = np.unique(cols.astype(np.int64) * ndofs + rows)
    col = key // ndofs
    row = key - col * ndofs
    colptr = np.zeros(ndofs + 1, dtype=np.int64)
    np.add.at(colptr, col + 1, 1)
    colptr = np.cumsum(colptr) + 1
    return colptr, row + 1


class Problem:
    """What `setup_problem` returns, flattened: dh (node_first_dof, cell_dofs, ndofs), the CSC
    pattern of K with its values, and f."""

    def __init__(self, points, cells):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.ne, self.npc = self.cells.shape
        self.nn = self.points.shape[0]
        self.node_first_dof, self.cell_dofs, self.ndofs = first_touch_dofs(self.cells, self.nn)
        self.colptr, self.rowval = sparsity_pattern(self.cell_dofs, self.ndofs)
        self.nzval = np.zeros(self.rowval.size)
        self.f = np.zeros(self.ndofs)

    @property
    def nnz(self):
        return self.rowval.size

    def K(self):
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.ndofs, self.ndofs))


def setup_problem(points, cells):
    return Problem(points, cells)


# ----------------------------------------------------------------------------------------------
# element stiffness + assembly  (FiniteElementAnalysis.jl:204-250, 654-707)
# ----------------------------------------------------------------------------------------------

def element_stiffness(points, cells, lam, mu):
    """Kₑ for every cell by the reference's quadrature loop (:677-699): for each q,
    ke[i,j] += (sym(∇N_i) ⊡ σ(sym(∇N_j))) dΩ with basis i = 3(a-1)+c ↦ e_c ⊗ ∇N_a.
    `lam`, `mu` scalars or (ne,) arrays.  → (ne, nb, nb), ke[e,i,j]."""
    ne, npc = cells.shape
    nb = 3 * npc
    lam = np.broadcast_to(np.asarray(lam, dtype=np.float64), (ne,))
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64), (ne,))
    nq = reference_element(npc)[2].size
    ke = np.zeros((ne, nb, nb))
    eye = np.eye(3)
    for q in range(nq):
        dNdx, dOm = cell_geometry(points, cells, q)
        G = np.zeros((ne, nb, 3, 3))                               # shape_gradient(cv,q,i) = e_c ⊗ ∇N_a
        for a in range(npc):
            for c in range(3):
                G[:, 3 * a + c, c, :] = dNdx[:, a, :]
        eps = 0.5 * (G + G.transpose(0, 1, 3, 2))                  # symmetric(∇N)        :690-691
        tr = np.trace(eps, axis1=2, axis2=3)
        sig = lam[:, None, None, None] * tr[:, :, None, None] * eye + 2 * mu[:, None, None, None] * eps   # :128
        ke += np.einsum("eiab,ejab->eij", eps, sig) * dOm[:, None, None]                                  # :697
    return ke


def element_stiffness_literal(X, lam, mu):
    """One cell, scalar loops exactly as :218-243 (pure Python; for cross-checking the batched form)."""
    npc = X.shape[0]
    nb = 3 * npc
    _, dN, w = reference_element(npc)
    ke = np.zeros((nb, nb))
    for q in range(w.size):
        J = sum(np.outer(X[a], dN[q, a]) for a in range(npc))
        detJ = np.linalg.det(J)
        assert detJ > 0
        dNdx = dN[q] @ np.linalg.inv(J)
        dOm = detJ * w[q]
        grads = []
        for i in range(nb):
            a, c = divmod(i, 3)
            g = np.zeros((3, 3)); g[c, :] = dNdx[a]
            grads.append(g)
        for i in range(nb):
            for j in range(nb):
                ei = 0.5 * (grads[i] + grads[i].T)
                ej = 0.5 * (grads[j] + grads[j].T)
                sig = lam * np.trace(ej) * np.eye(3) + 2 * mu * ej
                ke[i, j] += np.sum(ei * sig) * dOm
    return ke


def _slots(prob, rows0, cols0):
    """position in nzval of entries (rows0, cols0) (0-based) — sorted-column lookup."""
    key = prob._key if hasattr(prob, "_key") else None
    if key is None:
        col_of = np.repeat(np.arange(prob.ndofs, dtype=np.int64), np.diff(prob.colptr))
        key = col_of * prob.ndofs + (prob.rowval - 1)
        prob._key = key
    return np.searchsorted(key, cols0.astype(np.int64) * prob.ndofs + rows0)


def assemble_stiffness_matrix(prob, lam, mu):
    """`assemble_stiffness_matrix!` (:204-250): `start_assemble` zeroes K **and f** (:211), then
    K[celldofs,celldofs] += Kₑ in ascending cell order (Ferrite `assemble!`, Appendix A4)."""
    ke = element_stiffness(prob.points, prob.cells, lam, mu)
    prob.nzval[:] = 0.0
    prob.f[:] = 0.0
    nb = prob.cell_dofs.shape[1]
    rows = np.repeat(prob.cell_dofs - 1, nb, axis=1).reshape(-1)   # i index varies slowest
    cols = np.tile(prob.cell_dofs - 1, (1, nb)).reshape(-1)
    np.add.at(prob.nzval, _slots(prob, rows, cols), ke.reshape(-1))  # unbuffered, sequential ⇒ cell order
    return ke


def assemble_stiffness_matrix_simp(prob, material_model, density_data):
    """`assemble_stiffness_matrix_simp!` (:654-707): λ,μ = material_model(density[cellid]) per cell (:670-674)."""
    density_data = np.asarray(density_data, dtype=np.float64)
    lam, mu = material_model(density_data)
    return assemble_stiffness_matrix(prob, lam, mu)


# ----------------------------------------------------------------------------------------------
# loads
# ----------------------------------------------------------------------------------------------

def apply_force(prob, nodes, force_vector):
    """`apply_force!` (:392-418) with `get_node_dofs` (:265-293): f[dof(node,c)] += F[c]/length(nodes);
    nodes that belong to no cell are silently skipped (`haskey`, :402); empty set is an error (:393-395)."""
    nodes = np.asarray(list(nodes), dtype=np.int64)
    if nodes.size == 0:
        raise ValueError("No nodes provided for force application.")
    per_node = np.asarray(force_vector, dtype=np.float64) / nodes.size
    for g in nodes:
        d0 = prob.node_first_dof[g - 1]
        if d0 > 0:
            prob.f[d0 - 1:d0 + 2] += per_node


def _volume_force(prob, b, density_per_cell, active):
    N, _, w = reference_element(prob.npc)
    fe = np.zeros((prob.ne, 3 * prob.npc))
    total_volume = 0.0
    for q in range(w.size):
        _, dOm = cell_geometry(prob.points, prob.cells, q)
        total_volume += dOm.sum()
        for a in range(prob.npc):
            for c in range(3):
                fe[:, 3 * a + c] += density_per_cell * b[c] * N[q, a] * dOm
    fe[~active] = 0.0
    np.add.at(prob.f, (prob.cell_dofs - 1).reshape(-1), fe.reshape(-1))
    return fe[active].reshape(-1, prob.npc, 3).sum(axis=(0, 1)), total_volume


def apply_volume_force(prob, body_force_vector, density=1.0):
    """`apply_volume_force!` (VolumeForce.jl:26-94): divides b by `density` (:29) and multiplies it
    back (:76) — the net load is b·N·dΩ."""
    b = np.asarray(body_force_vector, dtype=np.float64) / density
    return _volume_force(prob, b, np.full(prob.ne, float(density)), np.ones(prob.ne, dtype=bool))


def apply_gravity(prob, density=1.0, g=9.81, direction=(0.0, 0.0, -1.0)):
    """`apply_gravity!` (VolumeForce.jl:112-132): b = ρ g d̂/‖d̂‖, then apply_volume_force!(…, 1.0) (:131)."""
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    return apply_volume_force(prob, density * g * d, 1.0)


def apply_variable_density_volume_force(prob, body_force_vector, density_data):
    """`apply_variable_density_volume_force!` (VolumeForce.jl:176-243): per-cell ρ, cells with ρ<1e-6 skipped (:199)."""
    rho = np.asarray(density_data, dtype=np.float64)
    return _volume_force(prob, np.asarray(body_force_vector, dtype=np.float64), rho, rho >= 1e-6)[0]


# ----------------------------------------------------------------------------------------------
# Dirichlet  (FiniteElementAnalysis.jl:314-333, 356-374;  Ferrite apply! at :540-542, :841-843)
# ----------------------------------------------------------------------------------------------

def fixed_boundary_dofs(prob, nodes, components=(1, 2, 3)):
    """`apply_fixed_boundary!` / `apply_sliding_boundary!`: the ConstraintHandler's sorted
    `prescribed_dofs` (1-based) — first_dof(node)+d-1 for nodes that belong to a cell; values 0."""
    nodes = np.asarray(sorted(set(int(g) for g in nodes)), dtype=np.int64)
    base = prob.node_first_dof[nodes - 1]
    base = base[base > 0]
    comps = np.asarray(sorted(set(components)), dtype=np.int64)
    return np.unique((base[:, None] + comps[None, :] - 1).reshape(-1))


def apply_dirichlet(prob, prescribed_dofs):
    """Ferrite `apply!(K,f,ch)` with zero-valued constraints (Appendix A5): m = mean(abs(diag K)) of the
    incoming K; stored entries of prescribed rows and columns → 0.0 (pattern kept); K[d,d]=m; f[d]=0."""
    d0 = np.asarray(prescribed_dofs, dtype=np.int64) - 1
    n = prob.ndofs
    col_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(prob.colptr))
    diag_slots = np.nonzero(col_of == prob.rowval - 1)[0]
    m = float(np.sum(np.abs(prob.nzval[diag_slots])) / n)
    flag = np.zeros(n, dtype=bool)
    flag[d0] = True
    hit = flag[col_of] | flag[prob.rowval - 1]
    prob.nzval[hit] = 0.0
    prob.nzval[diag_slots[flag[col_of[diag_slots]]]] = m
    prob.f[d0] = 0.0
    return m


# ----------------------------------------------------------------------------------------------
# solves + energy
# ----------------------------------------------------------------------------------------------

def solve_direct(prob):
    """`u = K \\ f` (:547, :848) — any accurate sparse direct solve; SuperLU here."""
    return spla.splu(prob.K()).solve(prob.f)


def jacobi_preconditioner(prob):
    """RobustSolver.jl:231-236: D = diag(K); D[abs(D) < 1e-12] = 1; M = Diagonal(1 ./ D)."""
    D = prob.K().diagonal().copy()
    D[np.abs(D) < 1e-12] = 1.0
    return 1.0 / D


def pcg_krylov(K, b, Minv, atol=1e-8, rtol=1e-8, itmax=10000):
    """Krylov.jl `cg(A,b; M, atol, rtol, itmax, history=true)` as the reference calls it
    (RobustSolver.jl:294-305, 337; Appendix A7).  x₀=0; stop on √(rᵀMr) ≤ atol + rtol·√(r₀ᵀMr₀).
    → (x, dict(niter, solved, residuals))."""
    x = np.zeros_like(b)
    r = b.copy()
    z = Minv * r
    p = z.copy()
    gamma = float(r @ z)
    rho0 = np.sqrt(gamma)
    eps = atol + rtol * rho0
    residuals = [rho0]
    k = 0
    while np.sqrt(gamma) > eps and k < itmax:
        Ap = K @ p
        pAp = float(p @ Ap)
        if pAp <= 0:
            break
        alpha = gamma / pAp
        x += alpha * p
        r -= alpha * Ap
        z = Minv * r
        gamma_new = float(r @ z)
        beta = gamma_new / gamma
        p = z + beta * p
        gamma = gamma_new
        k += 1
        residuals.append(np.sqrt(gamma))
    return x, {"niter": k, "solved": bool(np.sqrt(gamma) <= eps), "residuals": np.array(residuals)}


def solve_pcg(prob, tolerance=1e-8, itmax=10000):
    """`solve_with_krylov(K,f,:cg,config,…)` with `:diagonal` (RobustSolver.jl:279-338)."""
    return pcg_krylov(prob.K().tocsr(), prob.f, jacobi_preconditioner(prob), tolerance, tolerance, itmax)


def deformation_energy(prob, u):
    """`0.5 * dot(u, K*u)` with the constrained K (:550, :851; RobustSolver.jl:604, 717)."""
    return 0.5 * float(u @ (prob.K() @ u))


def element_energies(prob, u, ke):
    """North-star output (4): eₑ = ½ uₑᵀ Kₑ uₑ with the *unconstrained* Kₑ (the reference computes only
    the scalar; Σ eₑ equals it because u = 0 on prescribed DOFs — SURVEY F4)."""
    ue = u[prob.cell_dofs - 1]
    return 0.5 * np.einsum("ei,eij,ej->e", ue, ke, ue)


# ----------------------------------------------------------------------------------------------
# stress recovery
# ----------------------------------------------------------------------------------------------

def calculate_stresses(prob, u, lam, mu):
    """`calculate_stresses(_simp)` (:440-509, :730-801): σ_q = λ tr(ε) I + 2μ ε with ε = sym(Σ_a u_a ⊗ ∇N_a)
    per quadrature point; per-cell von Mises of the qp-averaged stress; max and 1-based argmax.
    → (sigma (ne,nq,3,3), von_mises (ne,), max_vm, argmax_cell)."""
    ne, npc = prob.cells.shape
    lam = np.broadcast_to(np.asarray(lam, dtype=np.float64), (ne,))
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64), (ne,))
    nq = reference_element(npc)[2].size
    ue = u[prob.cell_dofs - 1].reshape(ne, npc, 3)
    sig = np.zeros((ne, nq, 3, 3))
    for q in range(nq):
        dNdx, _ = cell_geometry(prob.points, prob.cells, q)
        grad = np.einsum("eac,ead->ecd", ue, dNdx)
        eps = 0.5 * (grad + grad.transpose(0, 2, 1))
        tr = np.trace(eps, axis1=1, axis2=2)
        sig[:, q] = lam[:, None, None] * tr[:, None, None] * np.eye(3) + 2 * mu[:, None, None] * eps
    avg = sig.mean(axis=1)
    s = avg
    vm = np.sqrt(0.5 * ((s[:, 0, 0] - s[:, 1, 1]) ** 2 + (s[:, 1, 1] - s[:, 2, 2]) ** 2 + (s[:, 2, 2] - s[:, 0, 0]) ** 2)
                 + 3.0 * (s[:, 0, 1] ** 2 + s[:, 1, 2] ** 2 + s[:, 0, 2] ** 2))
    am = int(np.argmax(vm))
    return sig, vm, float(vm[am]), am + 1


# ----------------------------------------------------------------------------------------------
# boundary-node selection and surface traction
# ----------------------------------------------------------------------------------------------
# local face tables, 1-based like the reference: get_face_nodes, FiniteElementAnalysis.jl:42-56
FACE_NODES = {4: [(1, 3, 2), (1, 2, 4), (2, 3, 4), (1, 4, 3)],
              8: [(1, 4, 3, 2), (1, 2, 6, 5), (2, 3, 7, 6), (3, 4, 8, 7), (1, 5, 8, 4), (5, 6, 7, 8)]}


def extract_surface_nodes(cells):
    """`extract_surface_nodes!` (SelectNodesForBC.jl:59-123): a face (sorted node tuple) that belongs to exactly one cell is a
    surface face; the surface nodes are the nodes of those faces.  → sorted 1-based node ids (the reference sorts them at :97)."""
    count = {}
    npc = cells.shape[1]
    for cell in cells:
        for face in FACE_NODES[npc]:
            key = tuple(sorted(int(cell[i - 1]) for i in face))
            count[key] = count.get(key, 0) + 1
    nodes = set()
    for key, c in count.items():
        if c == 1:
            nodes.update(key)
    return np.array(sorted(nodes), dtype=np.int64)


def select_nodes_by_plane(points, cells, point, normal, tolerance=1.0):
    """`select_nodes_by_plane` → `select_surface_nodes_by_plane` (SelectNodesForBC.jl:146-185, :325-335): surface nodes with
    abs(dot(x - point, normal/‖normal‖)) < tolerance (default tolerance 1.0, :327).  → sorted 1-based ids."""
    surf = extract_surface_nodes(cells)
    n = np.asarray(normal, dtype=np.float64) / np.linalg.norm(normal)
    d = np.abs((points[surf - 1] - np.asarray(point, dtype=np.float64)) @ n)
    return surf[d < tolerance]


def select_nodes_by_circle(points, cells, center, normal, radius, tolerance=1.0):
    """`select_nodes_by_circle` → `select_surface_nodes_by_circle` (SelectNodesForBC.jl:207-266, :357-368): nodes of the plane
    selection whose in-plane distance from the centre is <= radius + tolerance."""
    on_plane = select_nodes_by_plane(points, cells, center, normal, tolerance)
    n = np.asarray(normal, dtype=np.float64) / np.linalg.norm(normal)
    v = points[on_plane - 1] - np.asarray(center, dtype=np.float64)
    proj = v - np.outer(v @ n, n)
    return on_plane[np.linalg.norm(proj, axis=1) <= radius + tolerance]


def get_boundary_facets(cells, nodes):
    """`get_boundary_facets` (SurfaceTraction.jl:45-66): every (cell, local face), 1-based, all of whose vertices are in `nodes`
    — interior faces qualify too, exactly as in the reference.  → (n,2) int64 sorted by (cell, face)."""
    s = set(int(g) for g in nodes)
    npc = cells.shape[1]
    out = []
    for e, cell in enumerate(cells):
        for fid, face in enumerate(FACE_NODES[npc]):
            if all(int(cell[i - 1]) in s for i in face):
                out.append((e + 1, fid + 1))
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def facet_quadrature(points, cells, facets):
    """Ferrite `FacetValues(FacetQuadratureRule{Ref*}(2), ip)` as the reference uses it (SurfaceTraction.jl:95-118, 174-204):
    per facet the quadrature points x_q, dΓ_q = ‖∂x/∂s × ∂x/∂t‖ w_q and the values N_a(x_q) of the cell's shape functions.
    Tet4 face: 3-point rule of the triangle (barycentric (2/3,1/6,1/6) permutations, w = 1/6 on the reference triangle of area
    1/2) — point k sits next to face vertex k; Hex8 face: 2x2 Gauss on the bilinear quadrilateral in the face's node order.
    → (xq (nf,nqp,3), dgamma (nf,nqp), N (nf,nqp,npc))."""
    npc = cells.shape[1]
    nf = len(facets)
    nqp = 3 if npc == 4 else 4
    xq = np.zeros((nf, nqp, 3)); dg = np.zeros((nf, nqp)); N = np.zeros((nf, nqp, npc))
    g = 1.0 / np.sqrt(3.0)
    for i, (e, fid) in enumerate(facets):
        loc = [a - 1 for a in FACE_NODES[npc][fid - 1]]
        P = points[cells[e - 1, loc] - 1]
        if npc == 4:
            nvec = np.cross(P[1] - P[0], P[2] - P[0])
            for k in range(3):
                w = np.full(3, 1.0 / 6.0); w[k] = 2.0 / 3.0
                xq[i, k] = w @ P
                dg[i, k] = np.linalg.norm(nvec) / 6.0
                N[i, k, loc] = w
        else:
            for k, (s, t) in enumerate(((-g, -g), (g, -g), (g, g), (-g, g))):
                Nf = 0.25 * np.array([(1 - s) * (1 - t), (1 + s) * (1 - t), (1 + s) * (1 + t), (1 - s) * (1 + t)])
                dNs = 0.25 * np.array([-(1 - t), (1 - t), (1 + t), -(1 + t)])
                dNt = 0.25 * np.array([-(1 - s), -(1 + s), (1 + s), (1 - s)])
                xq[i, k] = Nf @ P
                dg[i, k] = np.linalg.norm(np.cross(dNs @ P, dNt @ P))
                N[i, k, loc] = Nf
    return xq, dg, N


def compute_boundary_area(points, cells, facets):
    """`compute_boundary_area` (SurfaceTraction.jl:88-122): Σ dΓ over the facets' quadrature points."""
    return float(facet_quadrature(points, cells, facets)[1].sum())


def apply_surface_traction(prob, facets, traction_function):
    """`apply_surface_traction!` (SurfaceTraction.jl:160-225): f[celldofs] += Σ_q (N_i · t(x_q)) dΓ_q.  → (area, total force)."""
    xq, dg, N = facet_quadrature(prob.points, prob.cells, facets)
    total = np.zeros(3); area = 0.0
    for i, (e, fid) in enumerate(facets):
        cd = prob.cell_dofs[e - 1].reshape(-1, 3)
        for k in range(xq.shape[1]):
            t = np.asarray(traction_function(*xq[i, k]), dtype=np.float64)
            prob.f[cd - 1] += N[i, k][:, None] * t[None, :] * dg[i, k]
            total += t * dg[i, k]; area += dg[i, k]
    return area, total


def apply_uniform_surface_traction(prob, facets, total_force_vector):
    """`apply_uniform_surface_traction!` (SurfaceTraction.jl:261-287): t = F / area (error if area < 1e-12), then the above."""
    area = compute_boundary_area(prob.points, prob.cells, facets)
    if area < 1e-12:
        raise ValueError("Boundary area is effectively zero. Check facet selection.")
    t = np.asarray(total_force_vector, dtype=np.float64) / area
    return apply_surface_traction(prob, facets, lambda x, y, z: t)
