/*
 * topopt_b200.h — C ABI of libtopopt_b200.so
 *
 * B200-native (sm_100a) replacement of the strain-energy evaluation path of jezekon/TopOptEval.jl:
 * element stiffness → sparse assembly → loads / Dirichlet → Jacobi-PCG → per-element energy / compliance.
 *
 * The reference has no FFI of its own: its boundary is the Julia function API of
 * TopOptEval.FiniteElementAnalysis (export list src/FiniteElementAnalysis/FiniteElementAnalysis.jl:11-24,
 * 75-87).  Each entry point below names the reference function whose body it replaces; the Julia `ccall`
 * shim (topopteval.jl_b200/julia/TopOptEvalB200.jl) and the Python ctypes mirror
 * (topopteval.jl_b200/api.py) keep the reference's names and argument order on top of these calls.
 *
 * Conventions
 *   - every function returns int status: 0 = OK, <0 = error (toe_last_error() gives the text; the Julia shim
 *     turns it into `error(msg)` like the reference's own `error(...)` calls, e.g. FiniteElementAnalysis.jl:394);
 *   - all indices that cross the ABI are int64 and **1-based** (Julia's Int), all reals are double;
 *   - host pointers are caller-owned, read/written only during the call; device state is owned by the ctx;
 *   - a ctx is driven by one host thread at a time; distinct contexts are independent (one ctx per GPU);
 *   - there is NO CPU fallback: without a usable sm_100 device toe_create fails.
 */
#ifndef TOPOPT_B200_H
#define TOPOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct toe_ctx toe_ctx;

/* assembly variants (toe_assemble_*) */
#define TOE_ASM_AUTO    0   /* = TOE_ASM_GATHER */
#define TOE_ASM_ATOMIC  1   /* one thread per element, red.global.add.f64 scatter */
#define TOE_ASM_GATHER  2   /* one thread per 3x3 block of K, loops its contributing elements in ascending
                               cell order (the reference's accumulation order, Ferrite assemble!): no atomics,
                               deterministic, coalesced writes */
#define TOE_ASM_ROWS    3   /* Tet4: a thread group per node row stages the geometry of the row's cells once in shared
                               memory, then one thread per block of the row accumulates in ascending cell order
                               (≈2.3x fewer FP64 operations than GATHER; same determinism and exact symmetry; values
                               differ from GATHER in the last bits only).  Falls back to GATHER for Hex8 / very wide rows */

/* toe_solve_pcg flags */
#define TOE_PCG_MATRIX_FREE   1   /* element-by-element operator, K is not read (nor needed) */
#define TOE_PCG_NO_GRAPH      2   /* launch kernels directly instead of replaying a CUDA graph */
#define TOE_PCG_TWO_LEVEL     4   /* preconditioner M⁻¹ = D⁻¹ + Z(ZᵀKZ)⁻¹Zᵀ instead of Jacobi: Z = rigid-body modes of the boxes of a
                                     coarse grid over the mesh (≤ 6144 coarse unknowns); same stopping rule on sqrt(r'Mr).
                                     SolverConfig.preconditioner = :two_level in the shims; works on partitioned contexts too */

#define TOE_PCG_L2_NORM       8   /* stop on ||r||_2 <= atol + rtol*||r0||_2 (r = the recurrence residual) instead of Krylov.jl's M-norm rule
                                     (SURVEY §8(b) norm_kind; the reference prints this norm at RobustSolver.jl:468).  Jacobi only.
                                     history / res0_M / res_M then hold l2 norms */

typedef struct toe_pcg_stats {
    int64_t niter;            /* Krylov.jl stats.niter */
    int32_t converged;        /* Krylov.jl stats.solved: sqrt(r'Mr) <= atol + rtol*sqrt(r0'Mr0) */
    int32_t breakdown;        /* 1 if p'Ap <= 0 was met (Krylov.jl stops there) */
    double  res0_M;           /* sqrt(r0' M r0) */
    double  res_M;            /* final sqrt(r' M r) */
    double  rel_res_l2;       /* ||f - K u||_2 / ||f||_2, recomputed after the solve (RobustSolver.jl:468) */
    double  solve_seconds;    /* device time of the iteration loop (CUDA events) */
    double  spmv_seconds;     /* niter x average operator time, measured on a few isolated launches after the solve */
    double  spmv_bytes;       /* algorithmic bytes of one operator application */
    int64_t kernel_launches;  /* kernels launched by this call */
    int64_t restarts;         /* always 0 (round 1 restarted partitioned solves after a breakdown; removed, the field keeps the layout) */
    int64_t coarse_dofs;      /* TOE_PCG_TWO_LEVEL: size of the coarse space (0 otherwise) */
    double  precond_seconds;  /* TOE_PCG_TWO_LEVEL: device time spent building ZᵀKZ and its inverse (part of solve_seconds) */
    double  true_res;         /* norm of f - K u recomputed after the solve, in the norm of the stopping test (M-norm with Jacobi, l2 with
                                 TOE_PCG_L2_NORM or the two-level preconditioner).  A converged Jacobi solve whose true_res exceeds 1e4 x the
                                 tolerance returns an error (corrupted iterates) */
} toe_pcg_stats;

typedef struct toe_timings {  /* device seconds of the last call of each stage (CUDA events) */
    double set_mesh, build_dofs, build_pattern, assemble, loads, dirichlet, solve, energy;
    int64_t kernel_launches;  /* total kernels launched by this ctx so far */
} toe_timings;

/* ---- lifecycle ------------------------------------------------------------------------------------ */
int         toe_version(void);
int         toe_create(int device, toe_ctx** out);
void        toe_destroy(toe_ctx* ctx);
const char* toe_last_error(toe_ctx* ctx);            /* ctx may be NULL: error of the last failed toe_create */
int         toe_get_timings(toe_ctx* ctx, toe_timings* out);
/* CUDA errors that were pending in the runtime's last-error slot before one of this ctx's kernel launches (left behind by an earlier
 * unchecked call); they never fail the launch they precede, but they are counted and the last one is kept (text owned by the ctx).
 * TOE_VERBOSE=1 in the environment also prints them to stderr. */
int         toe_debug_stale_cuda_errors(toe_ctx* ctx, int64_t* count_out, const char** last_out);
/* device stopwatch on the ctx's own stream (CUDA events): what bench.py brackets its timed region with */
int         toe_timer_start(toe_ctx* ctx);
int         toe_timer_stop(toe_ctx* ctx, double* seconds_out);

/* ---- setup_problem (FiniteElementAnalysis.jl:151-185) ------------------------------------------------ */
/* xyz: 3*nn doubles, node-major (= Julia 3×nn column-major, grid.nodes); conn: npc*ne int64, 1-based,
 * cell-major (= Julia npc×ne column-major, grid.cells); npc = 4 (Tetrahedron) or 8 (Hexahedron), chosen by
 * the caller from typeof(getcells(grid,1)) as :157 does.  Rejects npc∉{4,8} and out-of-range node ids. */
int toe_set_mesh(toe_ctx* ctx, int64_t nn, const double* xyz, int64_t ne, int npc, const int64_t* conn);
/* Ferrite close!(dh) (:174-176): first-touch numbering, 3 consecutive DOFs per node. */
int toe_build_dofs(toe_ctx* ctx, int64_t* ndofs_out);
/* node_first_dof[nn]: first DOF of each node (1-based), 0 for nodes in no cell  (what get_node_dofs :265-293 rebuilds) */
int toe_get_node_dofs(toe_ctx* ctx, int64_t* node_first_dof);
/* celldofs of cells first..first+count-1 (1-based), 3*npc each — dh.cell_dofs */
int toe_get_cell_dofs(toe_ctx* ctx, int64_t first, int64_t count, int64_t* out);
/* Ferrite allocate_matrix(dh) (:181): sparsity of K.  Pattern is structurally symmetric, so colptr/rowval
 * (CSC, 1-based, rows ascending per column) double as CSR rowptr/colind. */
int toe_build_pattern(toe_ctx* ctx, int64_t* nnz_out);
int toe_get_pattern(toe_ctx* ctx, int64_t* colptr /* ndofs+1 */, int64_t* rowval /* nnz */);

/* ---- assembly (FiniteElementAnalysis.jl:204-250, 654-707) -------------------------------------------- */
/* All three zero K and f first (start_assemble, :211/:661), reject det J <= 0 (Ferrite reinit!). */
/* assemble_stiffness_matrix!(K,f,dh,cv,λ,μ) */
int toe_assemble_lame(toe_ctx* ctx, double lambda, double mu, int variant);
/* assemble_stiffness_matrix_simp! with the closure of create_simp_material_model(E0,nu,Emin,p) (:616-634):
 * E = Emin + (E0-Emin)*density^p evaluated in the kernel. density: ne doubles. */
int toe_assemble_simp(toe_ctx* ctx, double E0, double nu, double Emin, double p, const double* density, int variant);
/* assemble_stiffness_matrix_simp! with an arbitrary material_model: the shim evaluates it on the host. */
int toe_assemble_lame_per_cell(toe_ctx* ctx, const double* lambda_e, const double* mu_e, int variant);
/* Sets the material exactly like the calls above but does not form K (matrix-free solves). Zeroes f. */
int toe_set_material_lame(toe_ctx* ctx, double lambda, double mu);
int toe_set_material_simp(toe_ctx* ctx, double E0, double nu, double Emin, double p, const double* density);
/* parity hook: Ke of cells first..first+count-1 (1-based), (3npc)^2 doubles each, column-major, with the
 * material of the last assemble/set_material call. */
int toe_ke_batch(toe_ctx* ctx, int64_t first, int64_t count, double* ke_out);
/* K.nzval in the order of toe_get_pattern */
int toe_get_values(toe_ctx* ctx, double* nzval);
int toe_get_diagonal(toe_ctx* ctx, double* diag /* ndofs */);

/* ---- loads ----------------------------------------------------------------------------------------- */
int toe_get_rhs(toe_ctx* ctx, double* f);
int toe_set_rhs(toe_ctx* ctx, const double* f);      /* lets host-side loads (surface-traction callbacks) flow in */
/* apply_force!(f,dh,nodes,F) (:392-418): f[dofs(node)] += F/nnodes; nnodes==0 is an error (:393-395);
 * nodes that belong to no cell are skipped (haskey, :402). */
int toe_add_nodal_force(toe_ctx* ctx, const int64_t* nodes, int64_t nnodes, const double F[3]);
/* apply_volume_force!/apply_gravity!/apply_acceleration! (VolumeForce.jl:26-159): density==NULL, load per
 * cell = rho_uniform*(b/rho_uniform)*N*dΩ.  apply_variable_density_volume_force! (:176-243): density!=NULL,
 * cells with density < skip_below (reference: 1e-6, :199) are skipped.  total_force_out[3] may be NULL. */
int toe_add_volume_force(toe_ctx* ctx, const double b[3], double rho_uniform, const double* density,
                         double skip_below, double* total_force_out);

/* ---- boundary-node selection and surface traction (SelectNodesForBC.jl, SurfaceTraction.jl) ------------------- */
/* Two-call pattern: pass nodes_out = NULL to get the count, then a buffer of that many int64 (ascending 1-based ids).
 * extract_surface_nodes! (SelectNodesForBC.jl:59-123): nodes of the faces that belong to exactly one cell. */
int toe_surface_nodes(toe_ctx* ctx, int64_t* nodes_out, int64_t* count_out);
/* select_nodes_by_plane (:325-335 → :146-185): surface nodes with abs(dot(x - point, normal/|normal|)) < tolerance
 * (the reference's default tolerance is 1.0). */
int toe_select_nodes_by_plane(toe_ctx* ctx, const double point[3], const double normal[3], double tolerance, int64_t* nodes_out, int64_t* count_out);
/* select_nodes_by_circle (:357-368 → :207-266): the plane selection, then in-plane distance from center <= radius + tolerance. */
int toe_select_nodes_by_circle(toe_ctx* ctx, const double center[3], const double normal[3], double radius, double tolerance,
                               int64_t* nodes_out, int64_t* count_out);
/* get_boundary_facets (SurfaceTraction.jl:45-66): (cell, local face) pairs, 1-based, Ferrite's facet numbering
 * (get_face_nodes, FiniteElementAnalysis.jl:42-56), whose vertices ALL lie in `nodes`; ascending.  facets_out = NULL: count only. */
int toe_boundary_facets(toe_ctx* ctx, const int64_t* nodes, int64_t nnodes, int64_t* facets_out /* 2*capacity */, int64_t capacity, int64_t* count_out);
/* compute_boundary_area (:88-122): sum of dΓ over FacetQuadratureRule{Ref*}(2) of every facet. */
int toe_boundary_area(toe_ctx* ctx, const int64_t* facets /* 2*nfacets */, int64_t nfacets, double* area_out);
/* quadrature points (3 per Tet4 face, 4 per Hex8 face) and dΓ of every facet — what a host callback (x,y,z) -> t needs
 * (apply_surface_traction! with an arbitrary Julia function, :160-225); either output may be NULL. */
int toe_facet_quadrature(toe_ctx* ctx, const int64_t* facets, int64_t nfacets, double* xq_out /* 3*nqp*nfacets */, double* dgamma_out /* nqp*nfacets */);
/* apply_surface_traction! (:160-225): f[celldofs] += Σ_q (N_i · t_q) dΓ_q with t_q = traction_qp (3 per quadrature point, in
 * the order of toe_facet_quadrature) or, if that is NULL, the constant traction_uniform (apply_uniform_surface_traction!,
 * :261-287, passes F/area).  area_out / total_force_out[3] (∫dΓ, ∫t dΓ) may be NULL. */
int toe_add_surface_traction(toe_ctx* ctx, const int64_t* facets, int64_t nfacets, const double* traction_qp, const double traction_uniform[3],
                             double* area_out, double* total_force_out);

/* ---- Dirichlet: Ferrite apply!(K,f,ch) (call sites :540-542, :841-843, RobustSolver.jl:542-544) ------- */
/* zero-valued constraints on dofs (1-based): m = mean(abs(diag K)) of the incoming K, stored entries of the
 * prescribed rows and columns become 0.0, K[d,d] = m, f[d] = 0.  Call once per ConstraintHandler, in order. */
int toe_apply_dirichlet(toe_ctx* ctx, const int64_t* dofs, int64_t ndofs, double* mean_diag_out);

/* ---- solve: solve_with_krylov(:cg, :diagonal) (RobustSolver.jl:223-236, 279-338) ---------------------- */
/* Krylov.jl cg semantics: x0 = 0, M = Diagonal(1 ./ D) with D[abs(D)<1e-12] = 1, stop on the M-norm test.
 * history (may be NULL): sqrt(r'Mr) per iteration, history_cap entries at most (stats.residuals). */
int toe_solve_pcg(toe_ctx* ctx, double atol, double rtol, int64_t itmax, int flags,
                  toe_pcg_stats* stats, double* history, int64_t history_cap);
int toe_get_solution(toe_ctx* ctx, double* u /* ndofs, Ferrite dof order */);
int toe_set_solution(toe_ctx* ctx, const double* u);  /* evaluate energies of a displacement field computed elsewhere */

/* ---- energy (FiniteElementAnalysis.jl:550, :851; RobustSolver.jl:604, :717) --------------------------- */
/* half_uKu = 0.5*dot(u,K*u) (sum of the per-element energies ½ uₑᵀKₑuₑ), compliance = f'u,
 * per_elem (ne doubles, may be NULL). */
int toe_energy(toe_ctx* ctx, double* half_uKu, double* compliance, double* per_elem);
/* the reference's exact expression with the assembled, constrained K (one SpMV + dot) */
int toe_energy_assembled(toe_ctx* ctx, double* half_uKu);

/* ---- stress recovery: calculate_stresses(_simp) (:440-509, :730-801) ---------------------------------- */
/* sigma (may be NULL): 6 x nqp x ne doubles (xx,yy,zz,xy,yz,xz per quadrature point, nqp = 4 tet / 8 hex);
 * von_mises (may be NULL): ne doubles, von Mises of the qp-averaged stress; max + 1-based argmax (first max wins). */
int toe_stresses(toe_ctx* ctx, double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell);
/* The reference's free functions: calculate_stresses(u, dh, cellvalues, λ, μ) (FiniteElementAnalysis.jl:440) and
 * calculate_stresses_simp(u, dh, cellvalues, material_model, density_data) (:730) take ANY displacement vector and material.
 * u: n doubles in Ferrite dof order, or NULL = the solution stored in the ctx.  These calls read the mesh only: K, f, the
 * constraints, the material and the stored solution of the ctx are left untouched.  Outputs as for toe_stresses. */
int toe_calculate_stresses(toe_ctx* ctx, const double* u, double lambda, double mu,
                           double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell);
int toe_calculate_stresses_simp(toe_ctx* ctx, const double* u, double E0, double nu, double Emin, double p, const double* density,
                                double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell);
/* material_model given as an arbitrary callable: the shim evaluates it per cell on the host (as toe_assemble_lame_per_cell) */
int toe_calculate_stresses_lame_per_cell(toe_ctx* ctx, const double* u, const double* lambda_e, const double* mu_e,
                                         double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell);

/* ---- operator hooks for tests and bench ----------------------------------------------------------- */
/* y = K x with the current (possibly constrained) operator; matrix_free selects the EbE kernel. */
int toe_spmv(toe_ctx* ctx, const double* x, double* y, int matrix_free);
/* times `reps` back-to-back operator applications on device vectors; returns average seconds and the
 * algorithmic bytes of one application. */
int toe_time_spmv(toe_ctx* ctx, int matrix_free, int reps, double* seconds_out, double* bytes_out);
/* diagnostic: applies the operator to the stored vector u `reps`+1 times and compares every result bit for bit with the first on the
 * device.  what: 1 = local product only, 2 = interface exchange only, 3 = product + exchange (2, 3: partitioned ctx).  The path is
 * deterministic, so both counts must be 0: mismatching_batches = batches of 256 applications in which a difference appeared,
 * mismatching_entries = differing entries in total. */
/* diagnostic (partitioned Jacobi-PCG run with TOE_CG_TRACE=1 in the environment): 4 doubles per iteration j —
 * {γ_j, δ_j as summed over the ranks, this rank's partial of γ_j, this rank's partial of δ_j}; out holds 4*iterations doubles */
int toe_debug_cg_trace(toe_ctx* ctx, double* out, int64_t iterations);
int toe_spmv_soak(toe_ctx* ctx, int matrix_free, int what, int64_t reps, int64_t* mismatching_batches, int64_t* mismatching_entries);

/* ---- multi-GPU: one ctx per GPU / process, element-based domain decomposition --------------------------- */
/* The interface-DOF exchange and the CG scalars travel over NCCL (dlopen'ed libnccl.so.2; one all-gather per operator application,
 * the default below 8 ranks; or send/recv + allreduce) or through a peer-memory kernel over NVLink (default from 8 ranks on) — see
 * toe_comm_info and TOE_DIST_XCHG.
 * Rank 0 creates the id, the host (torch.distributed / MPI / Distributed.jl) broadcasts its 128 bytes. */
int toe_comm_unique_id(char id_out[128]);
int toe_comm_init(toe_ctx* ctx, int nranks, int rank, const char id[128]);
/* Must be called after toe_comm_init and instead of toe_set_mesh: every rank passes the SAME global mesh;
 * the library numbers DOFs globally (identical to the 1-GPU numbering), splits the cells by recursive
 * coordinate bisection of their centroids and keeps only this rank's part (+ interface maps). */
int toe_set_mesh_distributed(toe_ctx* ctx, int64_t nn, const double* xyz, int64_t ne, int npc, const int64_t* conn);
/* part id (0-based) of every global cell (ne int32) — identical on all ranks */
int toe_get_partition(toe_ctx* ctx, int32_t* part_of_cell);
int toe_local_sizes(toe_ctx* ctx, int64_t* ne_local, int64_t* ndofs_local, int64_t* nnz_local, int64_t* n_interface_dofs);
/* transport of the per-iteration interface exchange: 0 = single GPU, 1 = NCCL send/recv + allreduce (TOE_DIST_XCHG=sendrecv),
 * 2 = fused peer-memory kernel (CUDA IPC mailboxes over NVLink/NVSwitch; default from 8 ranks on, TOE_DIST_XCHG=p2p; needs every
 *     rank to map every peer, else 3 is used),
 * 3 = one ncclAllGather per exchange carrying interface values and CG scalars (default below 8 ranks, TOE_DIST_XCHG=allgather) */
int toe_comm_info(toe_ctx* ctx, int* nranks, int* rank, int* transport);

#ifdef __cplusplus
}
#endif
#endif /* TOPOPT_B200_H */
