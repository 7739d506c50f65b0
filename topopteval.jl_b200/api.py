"""Host-side mirror of the reference's Julia API for the strain-energy evaluation path.

Same function names (minus Julia's `!`), argument order, return tuples and error behaviour as
`TopOptEval.FiniteElementAnalysis` (export list FiniteElementAnalysis.jl:11-24, 75-87) and the bits of
MeshImport / ResultsExport / Utils a user script touches (test/runtests.jl:21-89), so the parity tests read
like the reference's own tests.  Every numeric step runs in libtopopt_b200.so on the GPU through the C ABI;
this file only marshals arguments.  The Julia `ccall` shim in `julia/TopOptEvalB200.jl` is the same layer
for Julia users.

`K` and `f` are device-resident: the objects returned by `setup_problem` are handles (`K.to_scipy()`,
`f.to_numpy()` copy them out in the reference's layout: CSC / Ferrite dof order, 1-based where indices).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib, vtu
from ._lib import Context, TopOptError

__all__ = [
    "Grid", "import_mesh", "extract_cell_density", "calculate_volume",
    "create_material_model", "create_simp_material_model", "setup_problem",
    "assemble_stiffness_matrix", "assemble_stiffness_matrix_simp",
    "apply_fixed_boundary", "apply_sliding_boundary", "apply_force",
    "apply_volume_force", "apply_gravity", "apply_acceleration", "apply_variable_density_volume_force",
    "solve_system", "solve_system_simp", "solve_system_robust", "solve_system_robust_simp", "solve_system_adaptive",
    "calculate_stresses", "calculate_stresses_simp", "get_face_nodes",
    "SolverConfig", "export_results", "export_boundary_conditions", "TopOptError",
    "select_nodes_by_plane", "select_nodes_by_circle", "get_node_dofs",
    "get_boundary_facets", "compute_boundary_area", "apply_surface_traction", "apply_uniform_surface_traction",
]


# ------------------------------------------------------------------------------------------------------
# MeshImport (stays host-side; MeshImport.jl:20-164, 177-215)
# ------------------------------------------------------------------------------------------------------
@dataclass
class Grid:
    """What the path needs of a Ferrite.Grid: nodes (nn,3) and homogeneous cells (ne,npc), 1-based."""
    nodes: np.ndarray
    cells: np.ndarray
    cell_type: int = 10

    def getnnodes(self):
        return self.nodes.shape[0]

    def getncells(self):
        return self.cells.shape[0]


def import_mesh(mesh_file: str) -> Grid:
    if not mesh_file.lower().endswith(".vtu"):
        raise TopOptError("Unsupported mesh format: only .vtu is supported by this harness (MeshImport.jl:156)")
    m = vtu.read_vtu(mesh_file)
    return Grid(m.points, m.cells, m.cell_type)


extract_cell_density = vtu.extract_cell_density


# ------------------------------------------------------------------------------------------------------
# material models (FiniteElementAnalysis.jl:103-109, 616-634)
# ------------------------------------------------------------------------------------------------------
def create_material_model(youngs_modulus: float, poissons_ratio: float):
    lam = youngs_modulus * poissons_ratio / ((1 + poissons_ratio) * (1 - 2 * poissons_ratio))
    mu = youngs_modulus / (2 * (1 + poissons_ratio))
    return lam, mu


class SimpMaterialModel:
    """Callable like the closure the reference returns, but carrying (E0, nu, Emin, p) so that the kernel can
    evaluate E(ρ) itself (include/topopt_b200.h, toe_assemble_simp)."""

    def __init__(self, E0, nu, Emin, p):
        self.E0, self.nu, self.Emin, self.p = float(E0), float(nu), float(Emin), float(p)

    def __call__(self, density):
        E = self.Emin + (self.E0 - self.Emin) * density ** self.p
        return E * self.nu / ((1 + self.nu) * (1 - 2 * self.nu)), E / (2 * (1 + self.nu))


def create_simp_material_model(E0: float, nu: float, Emin: float = 1e-6, p: float = 1.0):
    """Defaults are the reference's *code* defaults (:619-620), not its docstring's."""
    return SimpMaterialModel(E0, nu, Emin, p)


# ------------------------------------------------------------------------------------------------------
# setup_problem (FiniteElementAnalysis.jl:151-185)
# ------------------------------------------------------------------------------------------------------
class DofHandler:
    def __init__(self, ctx: Context, grid: Grid):
        self.ctx, self.grid = ctx, grid
        self._node_dofs = None

    def ndofs(self):
        return self.ctx.ndofs

    @property
    def node_first_dof(self):
        if self._node_dofs is None:
            self._node_dofs = self.ctx.node_dofs()
        return self._node_dofs

    def celldofs(self, cell_id: int):
        return self.ctx.cell_dofs(cell_id, 1)[0]

    @property
    def cell_dofs(self):
        return self.ctx.cell_dofs()


@dataclass
class CellValues:
    npc: int
    nqp: int


class StiffnessMatrix:
    """Handle of the device-resident K (SparseMatrixCSC{Float64,Int} on the Julia side)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    @property
    def shape(self):
        return (self.ctx.ndofs, self.ctx.ndofs)

    def nnz(self):
        return self.ctx.nnz

    def pattern(self):
        return self.ctx.pattern()

    def nzval(self):
        return self.ctx.values()

    def to_scipy(self):
        import scipy.sparse as sp
        colptr, rowval = self.ctx.pattern()
        return sp.csc_matrix((self.ctx.values(), rowval - 1, colptr - 1), shape=self.shape)


class LoadVector:
    def __init__(self, ctx: Context):
        self.ctx = ctx

    def to_numpy(self):
        return self.ctx.rhs()

    def set(self, values):
        self.ctx.set_rhs(values)

    def __len__(self):
        return self.ctx.ndofs


def setup_problem(grid: Grid, interpolation_order: int = 1, device: int = 0, ctx: Context | None = None, distributed: bool = False):
    if interpolation_order != 1:
        raise TopOptError("only linear Lagrange interpolation is on the GPU path")
    npc = grid.cells.shape[1]
    print("Setting up problem with %s elements" % ("hexahedral" if npc == 8 else "tetrahedral"))
    ctx = ctx or Context(device)
    ctx.set_mesh(grid.nodes, grid.cells, distributed=distributed)
    n = ctx.build_dofs()
    print("Number of DOFs: %d" % n)
    ctx.build_pattern()
    if not distributed:
        grid._ctx = ctx                       # boundary-node selection on this grid reuses the device copy of the mesh
    return DofHandler(ctx, grid), CellValues(npc, 4 if npc == 4 else 8), StiffnessMatrix(ctx), LoadVector(ctx)


# ------------------------------------------------------------------------------------------------------
# assembly (FiniteElementAnalysis.jl:204-250, 654-707)
# ------------------------------------------------------------------------------------------------------
def assemble_stiffness_matrix(K, f, dh, cellvalues, lam, mu, variant=_lib.ASM_AUTO):
    dh.ctx.assemble_lame(lam, mu, variant)
    print("Stiffness matrix assembled successfully")


def assemble_stiffness_matrix_simp(K, f, dh, cellvalues, material_model, density_data, variant=_lib.ASM_AUTO):
    density_data = np.asarray(density_data, dtype=np.float64)
    if isinstance(material_model, SimpMaterialModel):
        dh.ctx.assemble_simp(material_model.E0, material_model.nu, material_model.Emin, material_model.p, density_data, variant)
    else:  # arbitrary callable ρ ↦ (λ, μ): evaluated on the host, per cell, like :670-674
        lm = [material_model(float(r)) for r in density_data]
        dh.ctx.assemble_lame_per_cell([a for a, _ in lm], [b for _, b in lm], variant)
    print("Stiffness matrix assembled successfully with variable material properties")


# ------------------------------------------------------------------------------------------------------
# constraints (FiniteElementAnalysis.jl:314-333, 356-374) — defined here, applied once in solve_*
# ------------------------------------------------------------------------------------------------------
@dataclass
class ConstraintHandler:
    prescribed_dofs: np.ndarray                       # sorted, 1-based
    inhomogeneities: np.ndarray = field(default=None)

    def __post_init__(self):
        if self.inhomogeneities is None:
            self.inhomogeneities = np.zeros(self.prescribed_dofs.size)


def _constraint(dh, nodes, comps):
    nodes = np.asarray(sorted(set(int(g) for g in nodes)), dtype=np.int64)
    nfd = dh.node_first_dof
    if nodes.size and (nodes.min() < 1 or nodes.max() > nfd.size):
        raise TopOptError("node id out of range in boundary condition")
    base = nfd[nodes - 1]
    base = base[base > 0]
    comps = np.asarray(sorted(set(int(c) for c in comps)), dtype=np.int64)
    return ConstraintHandler(np.unique((base[:, None] + comps[None, :] - 1).reshape(-1)))


def apply_fixed_boundary(K, f, dh, nodes):
    ch = _constraint(dh, nodes, (1, 2, 3))
    print("Defined fixed boundary conditions for %d nodes" % len(nodes))
    return ch


def apply_sliding_boundary(K, f, dh, nodes, fixed_dofs):
    ch = _constraint(dh, nodes, fixed_dofs)
    print("Defined sliding boundary conditions for %d nodes, fixing DOFs: %s" % (len(nodes), list(fixed_dofs)))
    return ch


# ------------------------------------------------------------------------------------------------------
# loads (FiniteElementAnalysis.jl:392-418; VolumeForce.jl)
# ------------------------------------------------------------------------------------------------------
def apply_force(f, dh, nodes, force_vector):
    if len(nodes) == 0:
        raise TopOptError("No nodes provided for force application.")
    dh.ctx.add_nodal_force(nodes, force_vector)
    print("Applied force %s distributed over %d nodes" % (list(force_vector), len(nodes)))


def apply_volume_force(f, dh, cellvalues, body_force_vector, density=1.0):
    tot = dh.ctx.add_volume_force(body_force_vector, rho_uniform=density)
    print("Applied volume force: %s N/m³" % list(body_force_vector))
    print("Total force applied: %s N" % list(tot))
    return tot


def apply_gravity(f, dh, cellvalues, density=1.0, g=9.81, direction=(0.0, 0.0, -1.0)):
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    return apply_volume_force(f, dh, cellvalues, density * g * d, 1.0)          # VolumeForce.jl:121-131


def apply_acceleration(f, dh, cellvalues, acceleration_vector, density=1.0):
    return apply_volume_force(f, dh, cellvalues, density * np.asarray(acceleration_vector, dtype=np.float64), 1.0)   # :151-158


def apply_variable_density_volume_force(f, dh, cellvalues, body_force_vector, density_data):
    tot = dh.ctx.add_volume_force(body_force_vector, density=density_data, skip_below=1e-6)    # VolumeForce.jl:199
    print("Applied variable density volume force")
    print("Total force applied: %s N" % list(tot))
    return tot


# ------------------------------------------------------------------------------------------------------
# boundary-node selection (SelectNodesForBC.jl) and surface traction (SurfaceTraction.jl)
# ------------------------------------------------------------------------------------------------------
def _grid_ctx(grid: Grid) -> Context:
    """The reference caches the surface nodes per grid (GRID_CACHE_STORAGE, SelectNodesForBC.jl:271-301); here the grid keeps
    the ctx that holds its mesh (set by setup_problem, or created on first use)."""
    ctx = getattr(grid, "_ctx", None)
    if ctx is None or getattr(ctx, "h", None) is None:
        ctx = Context(0)
        ctx.set_mesh(grid.nodes, grid.cells)
        ctx.build_dofs()
        ctx.build_pattern()
        grid._ctx = ctx
    return ctx


def select_nodes_by_plane(grid: Grid, point, normal, tolerance: float = 1.0):
    """SelectNodesForBC.jl:325-335 — surface nodes on the plane (point, normal); `tolerance` defaults to 1.0 like the reference."""
    nodes = _grid_ctx(grid).select_nodes_by_plane(point, normal, tolerance)
    print("Selected %d surface nodes on the specified plane" % nodes.size)
    return set(int(g) for g in nodes)


def select_nodes_by_circle(grid: Grid, center, normal, radius: float, tolerance: float = 1.0):
    """SelectNodesForBC.jl:357-368."""
    nodes = _grid_ctx(grid).select_nodes_by_circle(center, normal, radius, tolerance)
    print("Selected %d surface nodes in the circular region" % nodes.size)
    return set(int(g) for g in nodes)


class NodeDofs:
    """`get_node_dofs(dh)` (FiniteElementAnalysis.jl:265-293): node id -> its DOFs, as a read-only mapping over the device-built
    DOF map (the reference builds a Dict by sweeping all cells)."""

    def __init__(self, node_first_dof, dofs_per_node=3):
        self._first, self._n = node_first_dof, dofs_per_node

    def __contains__(self, node):
        return 1 <= node <= self._first.size and self._first[node - 1] > 0

    def __getitem__(self, node):
        if node not in self:
            raise KeyError(node)
        d = int(self._first[node - 1])
        return [d + k for k in range(self._n)]

    def __len__(self):
        return int(np.count_nonzero(self._first))

    def keys(self):
        return (np.nonzero(self._first)[0] + 1).tolist()


def get_node_dofs(dh):
    return NodeDofs(dh.node_first_dof)


def get_boundary_facets(grid: Grid, nodes):
    """SurfaceTraction.jl:45-66 → (n,2) int64 array of (cell_id, local_face_id), 1-based, ascending (the reference returns a Set)."""
    facets = _grid_ctx(grid).boundary_facets(nodes)
    print("Found %d boundary facets" % facets.shape[0])
    return facets


def compute_boundary_area(grid: Grid, dh, boundary_facets) -> float:
    """SurfaceTraction.jl:88-122."""
    return dh.ctx.boundary_area(boundary_facets)


def apply_surface_traction(f, dh, grid: Grid, boundary_facets, traction_function):
    """SurfaceTraction.jl:160-225: the callback (x, y, z) -> [Tx, Ty, Tz] is evaluated on the host at the facets' quadrature
    points; the quadrature itself and the scatter into f run on the GPU."""
    ctx = dh.ctx
    xq, _ = ctx.facet_quadrature(boundary_facets)
    tq = np.array([[np.asarray(traction_function(*x), dtype=np.float64) for x in fx] for fx in xq]).reshape(xq.shape)
    area, total = ctx.add_surface_traction(boundary_facets, traction_qp=tq)
    print("Applied surface traction over %d facets" % xq.shape[0])
    print("  Total boundary area: %s" % round(area, 6))
    print("  Total applied force: %s" % [round(float(v), 6) for v in total])
    return area, total


def apply_uniform_surface_traction(f, dh, grid: Grid, boundary_facets, total_force_vector):
    """SurfaceTraction.jl:261-287: t = F_total / area, error when the area is effectively zero."""
    ctx = dh.ctx
    area = ctx.boundary_area(boundary_facets)
    if area < 1e-12:
        raise TopOptError("Boundary area is effectively zero. Check facet selection.")
    traction = np.asarray(total_force_vector, dtype=np.float64) / area
    print("Uniform surface traction:")
    print("  Boundary area: %s" % round(area, 6))
    print("  Traction magnitude: %s" % round(float(np.linalg.norm(traction)), 6))
    area, total = ctx.add_surface_traction(boundary_facets, traction_uniform=traction)
    print("Applied surface traction over %d facets" % len(boundary_facets))
    return area, total


# ------------------------------------------------------------------------------------------------------
# solves (FiniteElementAnalysis.jl:538-598, 831-862; RobustSolver.jl:24-64, 530-734)
# ------------------------------------------------------------------------------------------------------
@dataclass
class SolverConfig:
    method: str = "auto"                 # :direct, :cg, :minres, :gmres, :bicgstab, :auto
    preconditioner: str = "diagonal"
    tolerance: float = 1e-8
    max_iterations: int = 0              # 0 → 10000 (RobustSolver.jl:49-51)
    memory_limit: float = 0.0
    verbose: bool = True
    restart: int = 30
    drop_tolerance: float = 1e-4
    history: bool = False
    matrix_free: bool = False            # extension: element-by-element operator instead of the assembled K
    l2_norm: bool = False                # extension: stop on ||r||_2 (TOE_PCG_L2_NORM) instead of Krylov.jl's M-norm rule (Jacobi only)

    def __post_init__(self):
        if self.max_iterations == 0:
            self.max_iterations = 10000
        if self.method not in ("auto", "cg", "direct"):
            raise TopOptError("method :%s is not on the GPU path (SPD system: :cg only)" % self.method)
        if self.preconditioner not in ("diagonal", "two_level"):
            raise TopOptError("preconditioner :%s is not on the GPU path (:diagonal = Jacobi as in RobustSolver.jl:231-236, or "
                              ":two_level = Jacobi + rigid-body coarse space)" % self.preconditioner)


class StressField:
    """Lazy stand-in for the reference's Dict{Int,Vector{SymmetricTensor}}: `sf[cell_id]` → (nqp,6) array
    (xx,yy,zz,xy,yz,xz); nothing is copied from the device until it is asked for (`fetch` = the call that produces it)."""

    def __init__(self, ctx, fetch=None):
        self.ctx = ctx
        self._fetch_fn = fetch or (lambda: ctx.stresses(True, True))
        self._sigma = None
        self._vm = None

    def _fetch(self):
        if self._sigma is None:
            self._sigma, self._vm, _, _ = self._fetch_fn()

    @property
    def sigma(self):
        self._fetch(); return self._sigma

    @property
    def von_mises(self):
        self._fetch(); return self._vm

    def __getitem__(self, cell_id):
        return self.sigma[cell_id - 1]

    def __len__(self):
        return self.ctx.ne


def _material_kwargs(material_model, density_data, ne):
    density_data = np.asarray(density_data, dtype=np.float64)
    if density_data.shape != (ne,):
        raise TopOptError("density_data has %d entries, the mesh has %d cells" % (density_data.size, ne))
    if isinstance(material_model, SimpMaterialModel):
        return {"simp": (material_model.E0, material_model.nu, material_model.Emin, material_model.p, density_data)}
    lm = [material_model(float(r)) for r in density_data]        # arbitrary callable ρ ↦ (λ, μ), evaluated per cell like :744-745
    return {"lame_per_cell": (np.array([a for a, _ in lm]), np.array([b for _, b in lm]))}


def calculate_stresses(u, dh, cellvalues, lam, mu):
    """`calculate_stresses(u, dh, cellvalues, λ, μ)` (FiniteElementAnalysis.jl:440-509): stresses of ANY displacement vector
    under ANY (λ, μ) → `(stress_field, max_von_mises, max_stress_cell)`.  One per-cell kernel + arg-max on the GPU; the
    σ arrays come to the host only when `stress_field` is indexed."""
    ctx = dh.ctx
    u = np.array(u, dtype=np.float64)                           # the field belongs to this call, not to the ctx's solution
    _, _, max_vm, max_cell = ctx.calculate_stresses(u, lame=(lam, mu))
    return StressField(ctx, lambda: ctx.calculate_stresses(u, lame=(lam, mu), want_sigma=True, want_vm=True)), max_vm, max_cell


def calculate_stresses_simp(u, dh, cellvalues, material_model, density_data):
    """`calculate_stresses_simp(u, dh, cellvalues, material_model, density_data)` (FiniteElementAnalysis.jl:730-801)."""
    ctx = dh.ctx
    u = np.array(u, dtype=np.float64)
    kw = _material_kwargs(material_model, density_data, ctx.ne)
    _, _, max_vm, max_cell = ctx.calculate_stresses(u, **kw)
    return StressField(ctx, lambda: ctx.calculate_stresses(u, want_sigma=True, want_vm=True, **kw)), max_vm, max_cell


#: Ferrite's local face → local node tables (FiniteElementAnalysis.jl:42-58), 1-based like the reference
_FACE_NODES = {
    4: [(1, 3, 2), (1, 2, 4), (2, 3, 4), (1, 4, 3)],
    8: [(1, 4, 3, 2), (1, 2, 6, 5), (2, 3, 7, 6), (3, 4, 8, 7), (1, 5, 8, 4), (5, 6, 7, 8)],
}


def get_face_nodes(cell):
    """`get_face_nodes(cell)` (FiniteElementAnalysis.jl:42-58): `cell` is a connectivity row (4 or 8 node ids), a Grid, or the
    number of nodes per cell."""
    if isinstance(cell, Grid):
        npc = cell.cells.shape[1]
    elif np.isscalar(cell):
        npc = int(cell)
    else:
        npc = len(cell)
    if npc not in _FACE_NODES:
        raise TopOptError("get_face_nodes: only Tetrahedron (4 nodes) and Hexahedron (8 nodes) cells are supported, got %d nodes" % npc)
    return list(_FACE_NODES[npc])


#: direct-solve entry points run PCG to this tolerance (no factorisation on the GPU path; the reference's
#: `K \ f` reaches a relative residual of ~1e-10 on its fixtures — SURVEY Appendix A6)
DIRECT_EQUIVALENT_TOL = 1e-10


def _solve(dh, constraints, tol, itmax, matrix_free, verbose, history=False, two_level=False, stress_material=None, l2_norm=False):
    """`stress_material`: keyword arguments of Context.calculate_stresses built from the CALLER's material arguments — the reference
    passes λ, μ (or material_model, density_data) of the solve call to calculate_stresses (FiniteElementAnalysis.jl:553 / :854), which
    may legally differ from what K was assembled with."""
    ctx = dh.ctx
    for ch in constraints:                              # SINGLE APPLICATION POINT (:540-542)
        ctx.apply_dirichlet(ch.prescribed_dofs)
    if verbose:
        print("Solving linear system...")
    st = ctx.solve_pcg(tol, tol, itmax, matrix_free=matrix_free, history=history, two_level=two_level, l2_norm=l2_norm)
    if st["breakdown"]:
        raise TopOptError("CG breakdown: p'Ap <= 0 (matrix not positive definite)")
    if not st["converged"] and verbose:
        print("WARNING: PCG did not converge in %d iterations (residual %.3e)" % (st["niter"], st["res_M"]))
    u = ctx.solution()
    energy, _compliance, _ = ctx.energy()
    if stress_material:
        _, _, max_vm, max_cell = ctx.calculate_stresses(None, **stress_material)
        sf = StressField(ctx, lambda: ctx.calculate_stresses(u, want_sigma=True, want_vm=True, **stress_material))
    else:
        _, _, max_vm, max_cell = ctx.stresses(False, False)
        sf = StressField(ctx)
    if verbose:
        print("Analysis complete")
        print("Deformation energy: %r J" % energy)
        print("Maximum von Mises stress: %r at cell %d" % (max_vm, max_cell))
    dh.last_stats = st
    return u, energy, sf, max_vm, max_cell


def _direct_itmax(dh):
    return max(100000, 4 * dh.ctx.ndofs)


def solve_system(K, f, dh, cellvalues, lam, mu, *constraints):
    return _solve(dh, constraints, DIRECT_EQUIVALENT_TOL, _direct_itmax(dh), False, True, stress_material={"lame": (lam, mu)})


def solve_system_simp(K, f, dh, cellvalues, material_model, density_data, *constraints):
    return _solve(dh, constraints, DIRECT_EQUIVALENT_TOL, _direct_itmax(dh), False, True,
                  stress_material=_material_kwargs(material_model, density_data, dh.ctx.ne))


def _robust(dh, constraints, config, stress_material):
    config = config or SolverConfig()
    tl = config.preconditioner == "two_level"
    # :direct, and :auto below 50 000 DOFs (select_solver_method, RobustSolver.jl:206, picks the factorisation there): the GPU path
    # has no factorisation, so PCG runs to the accuracy a direct solve delivers
    if config.method == "direct" or (config.method == "auto" and dh.ctx.ndofs < 50000):
        return _solve(dh, constraints, DIRECT_EQUIVALENT_TOL, _direct_itmax(dh), config.matrix_free, config.verbose, two_level=tl,
                      stress_material=stress_material)
    return _solve(dh, constraints, config.tolerance, config.max_iterations, config.matrix_free, config.verbose, config.history, two_level=tl,
                  stress_material=stress_material, l2_norm=config.l2_norm)


def solve_system_robust(K, f, dh, cellvalues, lam, mu, *constraints, config: SolverConfig | None = None):
    return _robust(dh, constraints, config, {"lame": (lam, mu)})


def solve_system_robust_simp(K, f, dh, cellvalues, material_model, density_data, *constraints, config: SolverConfig | None = None):
    return _robust(dh, constraints, config, _material_kwargs(material_model, density_data, dh.ctx.ne))


def solve_system_adaptive(K, f, dh, cellvalues, lam, mu, *constraints):
    n = dh.ctx.ndofs
    if n < 50000:                                                      # FiniteElementAnalysis.jl:574-575
        return solve_system(K, f, dh, cellvalues, lam, mu, *constraints)
    cfg = SolverConfig(method="auto", preconditioner="diagonal", tolerance=1e-7,
                       max_iterations=min(max(n // 10, 5000), 50000), verbose=True, restart=30, history=True)
    return solve_system_robust(K, f, dh, cellvalues, lam, mu, *constraints, config=cfg)


# ------------------------------------------------------------------------------------------------------
# ResultsExport / Utils (host-side)
# ------------------------------------------------------------------------------------------------------
def export_results(data, dh, output_file: str):
    """`export_results(u, dh, file)` (ResultsExport.jl:25-37): point data `u`; with a StressField: cell data
    `von_mises` (ResultsExport.jl:55-92)."""
    grid = dh.grid
    if isinstance(data, StressField):
        return vtu.write_vtu(output_file, grid.nodes, grid.cells, grid.cell_type, cell_data={"von_mises": data.von_mises})
    u = np.asarray(data, dtype=np.float64)
    nfd = dh.node_first_dof
    un = np.zeros((grid.nodes.shape[0], 3))
    ok = nfd > 0
    idx = nfd[ok] - 1
    un[ok] = np.stack([u[idx], u[idx + 1], u[idx + 2]], axis=1)
    return vtu.write_vtu(output_file, grid.nodes, grid.cells, grid.cell_type, point_data={"u": un})


#: face tables of `get_faces` in ResultsExport.jl:197-217 (NOT Ferrite's order — the export has its own)
_EXPORT_FACES = {
    4: [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)],
    8: [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)],
}


def export_boundary_conditions(grid: Grid, dh, fixed_nodes, force_nodes, output_file: str):
    """`export_boundary_conditions(grid, dh, fixed_nodes, force_nodes, file)` (ResultsExport.jl:108-193): every cell face whose
    nodes all carry the same mark (1 = fixed, 2 = force; force overrides fixed, :124-130) becomes a VTK_TRIANGLE / VTK_QUAD with
    cell data `boundary_type`, in cell order then face order (interior faces shared by two cells appear twice, as in the reference)."""
    print("Exporting mesh with boundary conditions to %s..." % output_file)
    nn = grid.nodes.shape[0]
    bc = np.zeros(nn + 1, dtype=np.int64)
    bc[np.asarray(sorted(fixed_nodes), dtype=np.int64)] = 1
    bc[np.asarray(sorted(force_nodes), dtype=np.int64)] = 2
    faces = np.asarray(_EXPORT_FACES[grid.cells.shape[1]])
    fn = grid.cells[:, faces]                                    # (ne, nfaces, nodes per face), 1-based node ids
    t = bc[fn]
    keep = np.all(t == t[..., :1], axis=2) & (t[..., 0] != 0)
    out = vtu.write_vtu(output_file, grid.nodes, fn[keep], 5 if faces.shape[1] == 3 else 9, cell_data={"boundary_type": t[..., 0][keep]})
    print("Boundary conditions successfully exported to %s" % out)
    return out


def calculate_volume(grid: Grid, density_data=None) -> float:
    """Utils.calculate_volume (Utils.jl:24-92): Σ ρₑ ∫dΩ — host-side diagnostic."""
    X = grid.nodes[grid.cells - 1]
    if grid.cells.shape[1] == 4:
        J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=2)
        vol = np.linalg.det(J) / 6.0
    else:
        from itertools import product
        s = np.array([(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)], float)
        g = 1 / np.sqrt(3.0)
        vol = np.zeros(X.shape[0])
        for qz, qy, qx in product((-g, g), repeat=3):
            dN = 0.125 * np.stack([s[:, 0] * (1 + qy * s[:, 1]) * (1 + qz * s[:, 2]),
                                   s[:, 1] * (1 + qx * s[:, 0]) * (1 + qz * s[:, 2]),
                                   s[:, 2] * (1 + qx * s[:, 0]) * (1 + qy * s[:, 1])], axis=1)
            vol += np.linalg.det(np.einsum("eai,aj->eij", X, dN))
    rho = 1.0 if density_data is None else np.asarray(density_data)
    return float(np.sum(rho * vol))
