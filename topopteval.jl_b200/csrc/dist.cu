// dist.cu — multi-GPU layer (one ctx per GPU / process).  Placeholder: the single-GPU hooks are no-ops.
#include "common.cuh"

struct DistState { int nranks = 1, rank = 0; };

bool dist_active(toe_ctx* ctx) { return ctx->dist != nullptr; }
void dist_destroy(toe_ctx* ctx) { delete ctx->dist; ctx->dist = nullptr; }
int dist_post_spmv(toe_ctx* ctx, double*) { (void)ctx; return TOE_OK; }
int dist_allreduce(toe_ctx* ctx, double*, int) { (void)ctx; return TOE_OK; }
int dist_comm_unique_id(char id_out[128], std::string& err) { (void)id_out; err = "multi-GPU layer not built"; return TOE_ERR_COMM; }
int dist_comm_init(toe_ctx* ctx, int, int, const char*) { return toe_fail(ctx, TOE_ERR_COMM, "multi-GPU layer not built"); }
int dist_set_mesh(toe_ctx* ctx, i64, const double*, i64, int, const int64_t*) { return toe_fail(ctx, TOE_ERR_COMM, "multi-GPU layer not built"); }
int dist_get_partition(toe_ctx* ctx, int32_t*) { return toe_fail(ctx, TOE_ERR_COMM, "multi-GPU layer not built"); }
int dist_local_sizes(toe_ctx* ctx, int64_t* a, int64_t* b, int64_t* c, int64_t* d) {
    if (a) *a = ctx->ne; if (b) *b = 3 * (int64_t)ctx->nq; if (c) *c = 9 * ctx->nnzb; if (d) *d = 0; return TOE_OK;
}
int dist_gather_vector(toe_ctx* ctx, const double*, double*) { return toe_fail(ctx, TOE_ERR_COMM, "multi-GPU layer not built"); }
int dist_scatter_vector(toe_ctx* ctx, const double*, double*) { return toe_fail(ctx, TOE_ERR_COMM, "multi-GPU layer not built"); }
int solve_pcg_dist(toe_ctx* ctx, double, double, i64, int, toe_pcg_stats*, double*, i64) { return toe_fail(ctx, TOE_ERR_COMM, "multi-GPU layer not built"); }
