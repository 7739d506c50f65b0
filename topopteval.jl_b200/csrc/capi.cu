// capi.cu — extern "C" entry points of libtopopt_b200.so (see include/topopt_b200.h for the contract and for the
// reference function each call replaces).
#include "common.cuh"
#include <cstring>
#include <mutex>

// implemented in the other translation units
int assemble_current_material(toe_ctx* ctx, int variant);
int ke_batch(toe_ctx* ctx, i64 first, i64 count, double* out_host);
int add_nodal_force(toe_ctx* ctx, const int64_t* nodes, i64 nnodes, const double F[3]);
int add_volume_force(toe_ctx* ctx, const double b[3], double rho_uniform, const double* density_host, double skip_below, double* total_out);
int apply_dirichlet(toe_ctx* ctx, const int64_t* dofs, i64 nd, double* mean_out);
int get_node_dofs(toe_ctx* ctx, int64_t* out_host);
int get_cell_dofs(toe_ctx* ctx, i64 first, i64 count, int64_t* out_host);
int get_pattern(toe_ctx* ctx, int64_t* colptr_host, int64_t* rowval_host);
int get_values(toe_ctx* ctx, double* nzval_host);
int solve_pcg(toe_ctx* ctx, double atol, double rtol, i64 itmax, int flags, toe_pcg_stats* stats, double* history, i64 history_cap);
int time_spmv(toe_ctx* ctx, int matrix_free, int reps, double* seconds_out, double* bytes_out);
int cg_trace(toe_ctx* ctx, double* out, i64 iterations);
int spmv_soak(toe_ctx* ctx, int matrix_free, int what, i64 reps, i64* mismatching_batches, i64* mismatching_entries);
int energy(toe_ctx* ctx, double* half_uKu, double* compliance, double* per_elem_host);
int energy_assembled(toe_ctx* ctx, double* half_uKu);
int stresses(toe_ctx* ctx, double* sigma_host, double* vm_host, double* max_vm, int64_t* max_cell);
int stresses_with(toe_ctx* ctx, const Material& mat, const double* u_dev, double* sigma_host, double* vm_host, double* max_vm, int64_t* max_cell);
int dist_comm_unique_id(char id_out[128], std::string& err);
int dist_comm_init(toe_ctx* ctx, int nranks, int rank, const char id[128]);
int dist_set_mesh(toe_ctx* ctx, i64 nn, const double* xyz, i64 ne, int npc, const int64_t* conn);
int dist_get_partition(toe_ctx* ctx, int32_t* part);
int dist_local_sizes(toe_ctx* ctx, int64_t* ne_local, int64_t* ndofs_local, int64_t* nnz_local, int64_t* nif);
int dist_gather_solution(toe_ctx* ctx, double* u_host);
int dist_scatter_vector(toe_ctx* ctx, const double* global_host, double* local_dev);
int dist_gather_vector(toe_ctx* ctx, const double* local_dev, double* global_host);
bool dist_active(toe_ctx* ctx);
int dist_info(toe_ctx* ctx, int* nranks, int* rank, int* transport);
int dist_localize_cells(toe_ctx* ctx, const double* global_host, double* local_dev);
int dist_node_dofs(toe_ctx* ctx, const int** node_q_g);
i64 dist_global_ne(toe_ctx* ctx);
i64 dist_global_ndofs(toe_ctx* ctx);

int select_nodes(toe_ctx* ctx, int mode, const double* point, const double* normal, double radius, double tol, int64_t* nodes_out, int64_t* count_out);
int boundary_facets(toe_ctx* ctx, const int64_t* nodes, i64 nnodes, int64_t* facets_out, i64 capacity, int64_t* count_out);
int facet_integrals(toe_ctx* ctx, const int64_t* facets, i64 nf, double* xq_out, double* dg_out, int add_load, const double* t_qp, const double* t_uniform,
                    double* area_out, double* total_force_out);

static std::string g_create_error;
static std::mutex g_mutex;

#define GUARD(ctx) if (!(ctx)) return TOE_ERR_ARG; { cudaError_t _e = cudaSetDevice((ctx)->device); \
    if (_e != cudaSuccess) return toe_fail((ctx), TOE_ERR_CUDA, "cudaSetDevice(%d): %s", (ctx)->device, cudaGetErrorString(_e)); }

extern "C" {

int toe_version(void) { return 100; }

const char* toe_last_error(toe_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    return g_create_error.c_str();
}

int toe_create(int device, toe_ctx** out) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!out) return TOE_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libtopopt_b200 has no CPU fallback";
        return TOE_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return TOE_ERR_ARG; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return TOE_ERR_CUDA; }
    if (prop.major != 10) {
        char buf[512];
        snprintf(buf, sizeof buf, "device %d (%s) is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.name, prop.major, prop.minor);
        g_create_error = buf;
        return TOE_ERR_CUDA;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return TOE_ERR_CUDA; }
    toe_ctx* c = new toe_ctx();
    c->device = device;
    {   // temporaries of the API calls come from the default memory pool (TmpBuf): keep what they free cached instead of returning it
        cudaMemPool_t pool = nullptr;
        unsigned long long keep = ~0ULL;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        cudaGetLastError();
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        g_create_error = "failed to create stream/events";
        delete c;
        return TOE_ERR_CUDA;
    }
    *out = c;
    return TOE_OK;
}

void toe_destroy(toe_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }   // before the communicator: a captured graph may hold NCCL work (TOE_DIST_GRAPH)
    dist_destroy(ctx);
    tl_destroy(ctx);
    if (ctx->cgs_host) cudaFreeHost(ctx->cgs_host);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    cudaStream_t s = ctx->stream;
    delete ctx;                 // DevBuf destructors free device memory
    if (s) cudaStreamDestroy(s);
}

int toe_debug_stale_cuda_errors(toe_ctx* ctx, int64_t* count_out, const char** last_out) {
    if (!ctx) return TOE_ERR_ARG;
    if (count_out) *count_out = ctx->stale_cuda_errors;
    if (last_out) *last_out = ctx->stale_err.c_str();
    return TOE_OK;
}

int toe_get_timings(toe_ctx* ctx, toe_timings* out) {
    if (!ctx || !out) return TOE_ERR_ARG;
    *out = ctx->tm; out->kernel_launches = ctx->launches;
    return TOE_OK;
}

int toe_timer_start(toe_ctx* ctx) {
    GUARD(ctx);
    if (!ctx->ev_t0) { CU(cudaEventCreate(&ctx->ev_t0)); CU(cudaEventCreate(&ctx->ev_t1)); }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventRecord(ctx->ev_t0, ctx->stream));
    return TOE_OK;
}
int toe_timer_stop(toe_ctx* ctx, double* seconds_out) {
    GUARD(ctx);
    if (!ctx->ev_t0) return toe_fail(ctx, TOE_ERR_STATE, "toe_timer_stop without toe_timer_start");
    CU(cudaEventRecord(ctx->ev_t1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev_t1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    if (seconds_out) *seconds_out = ms * 1e-3;
    return TOE_OK;
}

int toe_set_mesh(toe_ctx* ctx, int64_t nn, const double* xyz, int64_t ne, int npc, const int64_t* conn) {
    GUARD(ctx);
    if (dist_active(ctx)) return toe_fail(ctx, TOE_ERR_STATE, "toe_set_mesh: this ctx has a communicator; use toe_set_mesh_distributed");
    StageTimer T(ctx, &ctx->tm.set_mesh);
    TRY(mesh_upload(ctx, nn, xyz, ne, npc, conn));
    return T.finish();
}

int toe_build_dofs(toe_ctx* ctx, int64_t* ndofs_out) {
    GUARD(ctx);
    StageTimer T(ctx, &ctx->tm.build_dofs);
    if (!ctx->have_dofs) TRY(mesh_build_dofs(ctx));
    TRY(T.finish());
    if (ndofs_out) *ndofs_out = dist_global_ndofs(ctx);      // partitioned: the global count (u crosses the ABI in global Ferrite order)
    return TOE_OK;
}

int toe_get_node_dofs(toe_ctx* ctx, int64_t* node_first_dof) { GUARD(ctx); if (!node_first_dof) return TOE_ERR_ARG; return get_node_dofs(ctx, node_first_dof); }
int toe_get_cell_dofs(toe_ctx* ctx, int64_t first, int64_t count, int64_t* out) {
    GUARD(ctx); if (!out) return TOE_ERR_ARG;
    if (dist_active(ctx)) return toe_fail(ctx, TOE_ERR_STATE, "toe_get_cell_dofs: not available on a partitioned ctx (cells are distributed)");
    return get_cell_dofs(ctx, first, count, out);
}

int toe_build_pattern(toe_ctx* ctx, int64_t* nnz_out) {
    GUARD(ctx);
    StageTimer T(ctx, &ctx->tm.build_pattern);
    if (!ctx->have_dofs) TRY(mesh_build_dofs(ctx));
    if (!ctx->have_pattern) TRY(mesh_build_pattern(ctx));
    TRY(T.finish());
    if (nnz_out) *nnz_out = 9 * (int64_t)ctx->nnzb;
    return TOE_OK;
}

int toe_get_pattern(toe_ctx* ctx, int64_t* colptr, int64_t* rowval) {
    GUARD(ctx); if (!colptr || !rowval) return TOE_ERR_ARG;
    if (dist_active(ctx)) return toe_fail(ctx, TOE_ERR_STATE, "toe_get_pattern: K is sub-assembled per partition; not available on a partitioned ctx");
    return get_pattern(ctx, colptr, rowval);
}

static int set_material_common(toe_ctx* ctx) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "material: call setup_problem (toe_build_dofs) first");
    ctx->have_diag = false; ctx->have_solution = false;
    ctx->op_generation++;
    return TOE_OK;
}

static int set_lame(toe_ctx* ctx, double lambda, double mu) {
    TRY(set_material_common(ctx));
    ctx->mat = Material();
    ctx->mat.mode = MAT_UNIFORM; ctx->mat.lambda = lambda; ctx->mat.mu = mu;
    return TOE_OK;
}

static int set_simp(toe_ctx* ctx, double E0, double nu, double Emin, double p, const double* density) {
    TRY(set_material_common(ctx));
    if (!density) return toe_fail(ctx, TOE_ERR_ARG, "SIMP material needs a density vector");
    CU(ctx->density.alloc(ctx->ne));
    if (dist_active(ctx)) TRY(dist_localize_cells(ctx, density, ctx->density.p));     // density_data is indexed by global cell id
    else CU(cudaMemcpyAsync(ctx->density.p, density, ctx->ne * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->mat = Material();
    ctx->mat.mode = MAT_SIMP; ctx->mat.E0 = E0; ctx->mat.nu = nu; ctx->mat.Emin = Emin; ctx->mat.p = p;
    ctx->mat.density = ctx->density.p;
    return TOE_OK;
}

static int set_percell(toe_ctx* ctx, const double* lam, const double* mu) {
    TRY(set_material_common(ctx));
    if (!lam || !mu) return toe_fail(ctx, TOE_ERR_ARG, "per-cell material needs lambda and mu vectors");
    CU(ctx->lam_e.alloc(ctx->ne)); CU(ctx->mu_e.alloc(ctx->ne));
    if (dist_active(ctx)) { TRY(dist_localize_cells(ctx, lam, ctx->lam_e.p)); TRY(dist_localize_cells(ctx, mu, ctx->mu_e.p)); }
    else {
        CU(cudaMemcpyAsync(ctx->lam_e.p, lam, ctx->ne * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->mu_e.p, mu, ctx->ne * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->mat = Material();
    ctx->mat.mode = MAT_PERCELL; ctx->mat.lam_e = ctx->lam_e.p; ctx->mat.mu_e = ctx->mu_e.p;
    return TOE_OK;
}

// matrix-free path: same state change as an assembly (K and f zeroed, constraints dropped) without forming K
static int reset_for_matrix_free(toe_ctx* ctx) {
    TRY(ensure_vectors(ctx));
    size_t n = 3 * (size_t)ctx->nq;
    CU(cudaMemsetAsync(ctx->f.p, 0, n * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(ctx->dflag.p, 0, n, ctx->stream));
    CU(cudaMemsetAsync(ctx->dval.p, 0, n * sizeof(double), ctx->stream));
    ctx->any_dirichlet = false; ctx->have_K = false;
    return TOE_OK;
}

int toe_assemble_lame(toe_ctx* ctx, double lambda, double mu, int variant) {
    GUARD(ctx); TRY(set_lame(ctx, lambda, mu)); return assemble_current_material(ctx, variant);
}
int toe_assemble_simp(toe_ctx* ctx, double E0, double nu, double Emin, double p, const double* density, int variant) {
    GUARD(ctx); TRY(set_simp(ctx, E0, nu, Emin, p, density)); return assemble_current_material(ctx, variant);
}
int toe_assemble_lame_per_cell(toe_ctx* ctx, const double* lambda_e, const double* mu_e, int variant) {
    GUARD(ctx); TRY(set_percell(ctx, lambda_e, mu_e)); return assemble_current_material(ctx, variant);
}
int toe_set_material_lame(toe_ctx* ctx, double lambda, double mu) {
    GUARD(ctx); TRY(set_lame(ctx, lambda, mu)); return reset_for_matrix_free(ctx);
}
int toe_set_material_simp(toe_ctx* ctx, double E0, double nu, double Emin, double p, const double* density) {
    GUARD(ctx); TRY(set_simp(ctx, E0, nu, Emin, p, density)); return reset_for_matrix_free(ctx);
}

int toe_ke_batch(toe_ctx* ctx, int64_t first, int64_t count, double* ke_out) {
    GUARD(ctx); if (!ke_out) return TOE_ERR_ARG;
    if (dist_active(ctx)) return toe_fail(ctx, TOE_ERR_STATE, "toe_ke_batch: not available on a partitioned ctx");
    return ke_batch(ctx, first, count, ke_out);
}
int toe_get_values(toe_ctx* ctx, double* nzval) {
    GUARD(ctx); if (!nzval) return TOE_ERR_ARG;
    if (dist_active(ctx)) return toe_fail(ctx, TOE_ERR_STATE, "toe_get_values: K is sub-assembled per partition; not available on a partitioned ctx");
    return get_values(ctx, nzval);
}

static int get_vec(toe_ctx* ctx, const double* dev, double* host) {
    if (!host) return TOE_ERR_ARG;
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "DOFs not built");
    TRY(ensure_vectors(ctx));
    if (dist_active(ctx)) return dist_gather_vector(ctx, dev, host);
    CU(cudaMemcpyAsync(host, dev, 3 * (size_t)ctx->nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}
static int set_vec(toe_ctx* ctx, const double* host, double* dev) {
    if (!host) return TOE_ERR_ARG;
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "DOFs not built");
    TRY(ensure_vectors(ctx));
    if (dist_active(ctx)) return dist_scatter_vector(ctx, host, dev);
    CU(cudaMemcpyAsync(dev, host, 3 * (size_t)ctx->nq * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

int toe_get_diagonal(toe_ctx* ctx, double* diag) { GUARD(ctx); TRY(compute_diag(ctx)); return get_vec(ctx, ctx->diag.p, diag); }
int toe_get_rhs(toe_ctx* ctx, double* f) { GUARD(ctx); TRY(ensure_vectors(ctx)); return get_vec(ctx, ctx->f.p, f); }
int toe_set_rhs(toe_ctx* ctx, const double* f) { GUARD(ctx); TRY(ensure_vectors(ctx)); ctx->have_solution = false; return set_vec(ctx, f, ctx->f.p); }
int toe_get_solution(toe_ctx* ctx, double* u) { GUARD(ctx); TRY(ensure_vectors(ctx)); return get_vec(ctx, ctx->u.p, u); }
int toe_set_solution(toe_ctx* ctx, const double* u) { GUARD(ctx); TRY(ensure_vectors(ctx)); TRY(set_vec(ctx, u, ctx->u.p)); ctx->have_solution = true; return TOE_OK; }

int toe_add_nodal_force(toe_ctx* ctx, const int64_t* nodes, int64_t nnodes, const double F[3]) { GUARD(ctx); if (!F) return TOE_ERR_ARG; return add_nodal_force(ctx, nodes, nnodes, F); }
int toe_add_volume_force(toe_ctx* ctx, const double b[3], double rho_uniform, const double* density, double skip_below, double* total_force_out) {
    GUARD(ctx); if (!b) return TOE_ERR_ARG; return add_volume_force(ctx, b, rho_uniform, density, skip_below, total_force_out);
}
int toe_apply_dirichlet(toe_ctx* ctx, const int64_t* dofs, int64_t ndofs, double* mean_diag_out) { GUARD(ctx); return apply_dirichlet(ctx, dofs, ndofs, mean_diag_out); }

int toe_solve_pcg(toe_ctx* ctx, double atol, double rtol, int64_t itmax, int flags, toe_pcg_stats* stats, double* history, int64_t history_cap) {
    GUARD(ctx);
    if (stats) memset(stats, 0, sizeof *stats);
    return solve_pcg(ctx, atol, rtol, itmax, flags, stats, history, history_cap);
}

int toe_energy(toe_ctx* ctx, double* half_uKu, double* compliance, double* per_elem) { GUARD(ctx); return energy(ctx, half_uKu, compliance, per_elem); }
int toe_energy_assembled(toe_ctx* ctx, double* half_uKu) { GUARD(ctx); return energy_assembled(ctx, half_uKu); }
int toe_stresses(toe_ctx* ctx, double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell) {
    GUARD(ctx); return stresses(ctx, sigma, von_mises, max_von_mises, max_stress_cell);
}

// calculate_stresses(u, dh, cv, λ, μ) / calculate_stresses_simp(u, dh, cv, material_model, density_data): any u, any material,
// nothing of the ctx's own state is changed (u goes to the scratch vector, per-cell material data to temporaries)
static int stresses_for(toe_ctx* ctx, const double* u, Material mat, const double* a_host, const double* b_host,
                        double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "calculate_stresses: DOFs not built (setup_problem)");
    TRY(ensure_vectors(ctx));
    const double* u_dev = ctx->u.p;
    if (u) { TRY(set_vec(ctx, u, ctx->tmp.p)); u_dev = ctx->tmp.p; }
    else if (!ctx->have_solution) return toe_fail(ctx, TOE_ERR_STATE, "calculate_stresses: no displacement vector given and none stored");
    DevBuf<double> a, b;
    const double* hosts[2] = {a_host, b_host};
    DevBuf<double>* devs[2] = {&a, &b};
    for (int k = 0; k < 2; k++) {
        if (!hosts[k]) continue;
        CU(devs[k]->alloc(ctx->ne));
        if (dist_active(ctx)) TRY(dist_localize_cells(ctx, hosts[k], devs[k]->p));          // per-cell data is indexed by global cell id
        else CU(cudaMemcpyAsync(devs[k]->p, hosts[k], ctx->ne * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (mat.mode == MAT_SIMP) mat.density = a.p;
    if (mat.mode == MAT_PERCELL) { mat.lam_e = a.p; mat.mu_e = b.p; }
    return stresses_with(ctx, mat, u_dev, sigma, von_mises, max_von_mises, max_stress_cell);     // synchronises before the temporaries go
}
int toe_calculate_stresses(toe_ctx* ctx, const double* u, double lambda, double mu,
                           double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell) {
    GUARD(ctx);
    Material m = Material(); m.mode = MAT_UNIFORM; m.lambda = lambda; m.mu = mu;
    return stresses_for(ctx, u, m, nullptr, nullptr, sigma, von_mises, max_von_mises, max_stress_cell);
}
int toe_calculate_stresses_simp(toe_ctx* ctx, const double* u, double E0, double nu, double Emin, double p, const double* density,
                                double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell) {
    GUARD(ctx);
    if (!density) return toe_fail(ctx, TOE_ERR_ARG, "calculate_stresses_simp: density_data is required");
    Material m = Material(); m.mode = MAT_SIMP; m.E0 = E0; m.nu = nu; m.Emin = Emin; m.p = p;
    return stresses_for(ctx, u, m, density, nullptr, sigma, von_mises, max_von_mises, max_stress_cell);
}
int toe_calculate_stresses_lame_per_cell(toe_ctx* ctx, const double* u, const double* lambda_e, const double* mu_e,
                                         double* sigma, double* von_mises, double* max_von_mises, int64_t* max_stress_cell) {
    GUARD(ctx);
    if (!lambda_e || !mu_e) return toe_fail(ctx, TOE_ERR_ARG, "calculate_stresses: per-cell material needs lambda and mu vectors");
    Material m = Material(); m.mode = MAT_PERCELL;
    return stresses_for(ctx, u, m, lambda_e, mu_e, sigma, von_mises, max_von_mises, max_stress_cell);
}

int toe_surface_nodes(toe_ctx* ctx, int64_t* nodes_out, int64_t* count_out) { GUARD(ctx); return select_nodes(ctx, 0, nullptr, nullptr, 0.0, 0.0, nodes_out, count_out); }
int toe_select_nodes_by_plane(toe_ctx* ctx, const double point[3], const double normal[3], double tolerance, int64_t* nodes_out, int64_t* count_out) {
    GUARD(ctx); return select_nodes(ctx, 1, point, normal, 0.0, tolerance, nodes_out, count_out);
}
int toe_select_nodes_by_circle(toe_ctx* ctx, const double center[3], const double normal[3], double radius, double tolerance, int64_t* nodes_out, int64_t* count_out) {
    GUARD(ctx); return select_nodes(ctx, 2, center, normal, radius, tolerance, nodes_out, count_out);
}
int toe_boundary_facets(toe_ctx* ctx, const int64_t* nodes, int64_t nnodes, int64_t* facets_out, int64_t capacity, int64_t* count_out) {
    GUARD(ctx); return boundary_facets(ctx, nodes, nnodes, facets_out, capacity, count_out);
}
int toe_boundary_area(toe_ctx* ctx, const int64_t* facets, int64_t nfacets, double* area_out) {
    GUARD(ctx); return facet_integrals(ctx, facets, nfacets, nullptr, nullptr, 0, nullptr, nullptr, area_out, nullptr);
}
int toe_facet_quadrature(toe_ctx* ctx, const int64_t* facets, int64_t nfacets, double* xq_out, double* dgamma_out) {
    GUARD(ctx); return facet_integrals(ctx, facets, nfacets, xq_out, dgamma_out, 0, nullptr, nullptr, nullptr, nullptr);
}
int toe_add_surface_traction(toe_ctx* ctx, const int64_t* facets, int64_t nfacets, const double* traction_qp, const double traction_uniform[3],
                             double* area_out, double* total_force_out) {
    GUARD(ctx); return facet_integrals(ctx, facets, nfacets, nullptr, nullptr, 1, traction_qp, traction_uniform, area_out, total_force_out);
}

int toe_spmv(toe_ctx* ctx, const double* x, double* y, int matrix_free) {
    GUARD(ctx);
    if (!x || !y) return TOE_ERR_ARG;
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "DOFs not built");
    TRY(ensure_vectors(ctx));
    TRY(set_vec(ctx, x, ctx->r.p));
    TRY(op_apply(ctx, ctx->r.p, ctx->tmp.p, matrix_free, nullptr, false));
    return get_vec(ctx, ctx->tmp.p, y);
}

int toe_time_spmv(toe_ctx* ctx, int matrix_free, int reps, double* seconds_out, double* bytes_out) { GUARD(ctx); return time_spmv(ctx, matrix_free, reps, seconds_out, bytes_out); }
int toe_debug_cg_trace(toe_ctx* ctx, double* out, int64_t iterations) { GUARD(ctx); if (!out) return TOE_ERR_ARG; return cg_trace(ctx, out, iterations); }
int toe_spmv_soak(toe_ctx* ctx, int matrix_free, int what, int64_t reps, int64_t* mismatching_batches, int64_t* mismatching_entries) {
    GUARD(ctx);
    i64 a = 0, b = 0;
    int st = spmv_soak(ctx, matrix_free, what, reps, &a, &b);
    if (mismatching_batches) *mismatching_batches = a;
    if (mismatching_entries) *mismatching_entries = b;
    return st;
}

int toe_comm_unique_id(char id_out[128]) {
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!id_out) return TOE_ERR_ARG;
    return dist_comm_unique_id(id_out, g_create_error);
}
int toe_comm_init(toe_ctx* ctx, int nranks, int rank, const char id[128]) { GUARD(ctx); if (!id) return TOE_ERR_ARG; return dist_comm_init(ctx, nranks, rank, id); }
int toe_set_mesh_distributed(toe_ctx* ctx, int64_t nn, const double* xyz, int64_t ne, int npc, const int64_t* conn) {
    GUARD(ctx);
    StageTimer T(ctx, &ctx->tm.set_mesh);
    TRY(dist_set_mesh(ctx, nn, xyz, ne, npc, conn));
    return T.finish();
}
int toe_get_partition(toe_ctx* ctx, int32_t* part_of_cell) { GUARD(ctx); if (!part_of_cell) return TOE_ERR_ARG; return dist_get_partition(ctx, part_of_cell); }
int toe_comm_info(toe_ctx* ctx, int* nranks, int* rank, int* transport) { if (!ctx) return TOE_ERR_ARG; return dist_info(ctx, nranks, rank, transport); }
int toe_local_sizes(toe_ctx* ctx, int64_t* ne_local, int64_t* ndofs_local, int64_t* nnz_local, int64_t* n_interface_dofs) {
    GUARD(ctx); return dist_local_sizes(ctx, ne_local, ndofs_local, nnz_local, n_interface_dofs);
}

}  // extern "C"
