// mesh.cu — setup_problem on the GPU: first-touch DOF numbering (Ferrite close!(dh), call site
// FiniteElementAnalysis.jl:174-176), node→element incidence, sorted block sparsity pattern
// (Ferrite allocate_matrix(dh), :181) and the block→element contribution lists used by the
// gather assembly.  All of it is integer work, done once per mesh.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// exclusive scan (int32), reduce-then-scan over tiles of SCAN_TILE items
// ---------------------------------------------------------------------------------------------------------
static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 8;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void k_scan_tile_sums(const int* __restrict__ in, int* __restrict__ sums, i64 n) {
    __shared__ int sh[32];
    i64 base = (i64)blockIdx.x * SCAN_TILE;
    int s = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        i64 i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < SCAN_THREADS / 32 ? sh[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = s;
    }
}

// scans one tile; thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
// `in` and `out` may be the same array (every thread reads its own items before it writes them): no __restrict__ on the two
__global__ void k_scan_tiles(const int* in, int* out, const int* __restrict__ tile_offsets, i64 n) {
    __shared__ int sh[SCAN_THREADS / 32];
    i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { i64 i = base + k; v[k] = i < n ? in[i] : 0; s += v[k]; }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = s;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sh[w] = incl;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; k++) woff += sh[k];
    int run = (tile_offsets ? tile_offsets[blockIdx.x] : 0) + woff + incl - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { i64 i = base + k; if (i < n) out[i] = run; run += v[k]; }
}

// 64-bit total of a non-negative int array: the 32-bit scan below is only valid if this fits an int
__global__ void k_total_i64(const int* __restrict__ in, i64 n, unsigned long long* out) {
    unsigned long long s = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) s += (unsigned long long)(unsigned int)in[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

__global__ void k_write_total(const int* __restrict__ in_last, const int* __restrict__ out_last_excl, int* out_total, int have) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out_total = have ? (*out_last_excl + *in_last) : 0;
}

// out[0..n) = exclusive scan of in[0..n), out[n] = total.  `out` may alias `in` only if out has n+1 entries and
// the caller no longer needs in (the scan kernel reads a tile fully before writing it).
int scan_exclusive_i32(toe_ctx* ctx, const int* in, int* out, i64 n, i64* total_out) {
    if (n == 0) { CU(cudaMemsetAsync(out, 0, sizeof(int), ctx->stream)); if (total_out) *total_out = 0; return TOE_OK; }
    i64 ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    {   // overflow guard: the prefix sums are 32-bit, so the total must fit (checked in 64 bits BEFORE scanning; in may alias out)
        TmpBuf<unsigned long long> tot64(ctx->stream); CU(tot64.alloc(1));
        CU(cudaMemsetAsync(tot64.p, 0, sizeof(unsigned long long), ctx->stream));
        LAUNCH(ctx, k_total_i64, min_u(div_up(n, 256), 2048u), 256, 0, in, n, tot64.p);
        unsigned long long h = 0;
        CU(cudaMemcpyAsync(&h, tot64.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (h > 2147483647ULL) return toe_fail(ctx, TOE_ERR_MESH, "index overflow: a prefix sum would reach %llu > 2^31-1", h);
    }
    // keep the last input element: aliasing would overwrite it before k_write_total reads it
    TmpBuf<int> last_in(ctx->stream); CU(last_in.alloc(1));
    CU(cudaMemcpyAsync(last_in.p, in + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ntiles == 1) {
        LAUNCH(ctx, k_scan_tiles, 1, SCAN_THREADS, 0, in, out, (const int*)nullptr, n);
    } else {
        TmpBuf<int> sums(ctx->stream); CU(sums.alloc(ntiles + 1));
        LAUNCH(ctx, k_scan_tile_sums, (unsigned)ntiles, SCAN_THREADS, 0, in, sums.p, n);
        i64 dummy;
        TRY(scan_exclusive_i32(ctx, sums.p, sums.p, ntiles, &dummy));
        LAUNCH(ctx, k_scan_tiles, (unsigned)ntiles, SCAN_THREADS, 0, in, out, (const int*)sums.p, n);
        CU(cudaStreamSynchronize(ctx->stream));   // sums goes out of scope
    }
    LAUNCH(ctx, k_write_total, 1, 32, 0, (const int*)last_in.p, (const int*)(out + (n - 1)), out + n, 1);
    int tot = 0;
    CU(cudaMemcpyAsync(&tot, out + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (tot < 0) return toe_fail(ctx, TOE_ERR_MESH, "index overflow in prefix sum");
    if (total_out) *total_out = tot;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// mesh upload
// ---------------------------------------------------------------------------------------------------------
__global__ void k_conn_to_i32(const int64_t* __restrict__ conn1, int* __restrict__ conn0, i64 total, i64 nn, int* err) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t g = conn1[i];
    if (g < 1 || g > nn) { atomicExch(err, 1); conn0[i] = 0; return; }
    conn0[i] = (int)(g - 1);
}

int mesh_upload(toe_ctx* ctx, i64 nn, const double* xyz, i64 ne, int npc, const int64_t* conn) {
    if (npc != 4 && npc != 8) return toe_fail(ctx, TOE_ERR_ARG, "unsupported cell type: %d nodes per cell (Tetrahedron=4 and Hexahedron=8 are supported)", npc);
    if (nn <= 0 || ne <= 0 || !xyz || !conn) return toe_fail(ctx, TOE_ERR_ARG, "empty mesh (nn=%lld, ne=%lld)", (long long)nn, (long long)ne);
    if (nn > 2147483647LL / 3 || ne * npc > 2147483647LL)
        return toe_fail(ctx, TOE_ERR_ARG, "mesh too large for 32-bit device indices (nn=%lld, ne=%lld)", (long long)nn, (long long)ne);
    ctx->have_mesh = ctx->have_dofs = ctx->have_pattern = ctx->have_contrib = ctx->have_K = ctx->have_solution = false;
    ctx->have_tiles = false; ctx->have_surface = false;
    tl_invalidate(ctx);
    ctx->have_diag = false; ctx->any_dirichlet = false;
    ctx->mat.mode = MAT_NONE;
    ctx->op_generation++;
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; ctx->graph_key = -1; }
    ctx->nn = nn; ctx->ne = ne; ctx->npc = npc;
    i64 total = ne * npc;
    CU(ctx->xyz.alloc(3 * nn));
    CU(ctx->conn0.alloc(total));
    CU(ctx->errflag.alloc(4));
    CU(cudaMemsetAsync(ctx->errflag.p, 0, 4 * sizeof(int), ctx->stream));
    CU(cudaMemcpyAsync(ctx->xyz.p, xyz, 3 * nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    {
        TmpBuf<int64_t> tmp(ctx->stream); CU(tmp.alloc(total));
        CU(cudaMemcpyAsync(tmp.p, conn, total * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, k_conn_to_i32, div_up(total, 256), 256, 0, (const int64_t*)tmp.p, ctx->conn0.p, total, nn, ctx->errflag.p);
        int e = 0;
        CU(cudaMemcpyAsync(&e, ctx->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (e) return toe_fail(ctx, TOE_ERR_MESH, "cell connectivity refers to a node id outside 1..%lld", (long long)nn);
    }
    ctx->have_mesh = true;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// first-touch DOF numbering
// ---------------------------------------------------------------------------------------------------------
__global__ void k_fill_u64(u64* p, u64 v, i64 n) { i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

__global__ void k_first_touch(const int* __restrict__ conn0, u64* key, i64 total) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) atomicMin(&key[conn0[i]], (u64)i);
}
__global__ void k_flag_first(const int* __restrict__ conn0, const u64* __restrict__ key, int* __restrict__ flag, i64 total) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) flag[i] = (key[conn0[i]] == (u64)i) ? 1 : 0;
}
__global__ void k_node_q(const u64* __restrict__ key, const int* __restrict__ rank, int* __restrict__ node_q,
                         const double* __restrict__ xyz, double* __restrict__ xq, i64 nn) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nn) return;
    u64 k = key[g];
    if (k == ~0ULL) { node_q[g] = -1; return; }
    int q = rank[k];
    node_q[g] = q;
    xq[3 * (size_t)q] = xyz[3 * g]; xq[3 * (size_t)q + 1] = xyz[3 * g + 1]; xq[3 * (size_t)q + 2] = xyz[3 * g + 2];
}
__global__ void k_cq(const int* __restrict__ conn0, const int* __restrict__ node_q, int* __restrict__ cq, i64 total) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) cq[i] = node_q[conn0[i]];
}

int mesh_build_dofs(toe_ctx* ctx) {
    if (!ctx->have_mesh) return toe_fail(ctx, TOE_ERR_STATE, "toe_build_dofs: no mesh set");
    i64 total = ctx->ne * ctx->npc;
    TmpBuf<u64> key(ctx->stream); CU(key.alloc(ctx->nn));
    TmpBuf<int> flag(ctx->stream); CU(flag.alloc(total + 1));
    LAUNCH(ctx, k_fill_u64, div_up(ctx->nn, 256), 256, 0, key.p, ~0ULL, ctx->nn);
    LAUNCH(ctx, k_first_touch, div_up(total, 256), 256, 0, (const int*)ctx->conn0.p, key.p, total);
    LAUNCH(ctx, k_flag_first, div_up(total, 256), 256, 0, (const int*)ctx->conn0.p, (const u64*)key.p, flag.p, total);
    i64 nq = 0;
    TRY(scan_exclusive_i32(ctx, flag.p, flag.p, total, &nq));
    ctx->nq = (int)nq;
    CU(ctx->node_q.alloc(ctx->nn));
    CU(ctx->xq.alloc(3 * nq));
    CU(ctx->cq.alloc(total));
    LAUNCH(ctx, k_node_q, div_up(ctx->nn, 256), 256, 0, (const u64*)key.p, (const int*)flag.p, ctx->node_q.p,
           (const double*)ctx->xyz.p, ctx->xq.p, ctx->nn);
    LAUNCH(ctx, k_cq, div_up(total, 256), 256, 0, (const int*)ctx->conn0.p, (const int*)ctx->node_q.p, ctx->cq.p, total);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_dofs = true;
    ctx->have_pattern = ctx->have_contrib = ctx->have_K = false; ctx->have_tiles = false;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// node -> element incidence, entries e*npc+a ascending
// ---------------------------------------------------------------------------------------------------------
__global__ void k_count_inc(const int* __restrict__ cq, int* cnt, i64 total) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) atomicAdd(&cnt[cq[i]], 1);
}
__global__ void k_fill_inc(const int* __restrict__ cq, const int* __restrict__ inc_ptr, int* cursor, int* __restrict__ inc, i64 total) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int q = cq[i];
    int pos = inc_ptr[q] + atomicAdd(&cursor[q], 1);
    inc[pos] = (int)i;
}
__global__ void k_sort_inc(const int* __restrict__ inc_ptr, int* __restrict__ inc, int nq) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int lo = inc_ptr[q], hi = inc_ptr[q + 1];
    for (int i = lo + 1; i < hi; i++) {          // insertion sort, lists are short
        int v = inc[i], j = i - 1;
        while (j >= lo && inc[j] > v) { inc[j + 1] = inc[j]; j--; }
        inc[j + 1] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// block pattern: row q = sorted set of dof-nodes sharing a cell with q
// one thread per row, "next larger value" sweeps over the candidates (no scratch memory)
// ---------------------------------------------------------------------------------------------------------
template <int NPC, bool FILL>
__global__ void k_adjacency(const int* __restrict__ inc_ptr, const int* __restrict__ inc, const int* __restrict__ cq,
                            int* __restrict__ deg, const int* __restrict__ blk_ptr, int* __restrict__ blk_col,
                            int* __restrict__ diag_slot, int nq, int* max_deg) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int lo = inc_ptr[q], hi = inc_ptr[q + 1];
    int prev = -1, count = 0;
    int base = FILL ? blk_ptr[q] : 0;
    while (true) {
        int next = 0x7fffffff;
        for (int k = lo; k < hi; k++) {
            int e = inc[k] / NPC;
            const int* c = cq + (size_t)e * NPC;
#pragma unroll
            for (int a = 0; a < NPC; a++) { int v = c[a]; if (v > prev && v < next) next = v; }
        }
        if (next == 0x7fffffff) break;
        if (FILL) { blk_col[base + count] = next; if (next == q) diag_slot[q] = base + count; }
        prev = next; count++;
    }
    if (!FILL) { deg[q] = count; atomicMax(max_deg, count); }
}

int mesh_build_pattern(toe_ctx* ctx) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "toe_build_pattern: DOFs not built");
    i64 total = ctx->ne * ctx->npc;
    int nq = ctx->nq;
    CU(ctx->inc_ptr.alloc(nq + 1));
    CU(ctx->inc.alloc(total));
    {
        TmpBuf<int> cursor(ctx->stream); CU(cursor.alloc(nq));
        CU(cudaMemsetAsync(ctx->inc_ptr.p, 0, (nq + 1) * sizeof(int), ctx->stream));
        CU(cudaMemsetAsync(cursor.p, 0, nq * sizeof(int), ctx->stream));
        LAUNCH(ctx, k_count_inc, div_up(total, 256), 256, 0, (const int*)ctx->cq.p, ctx->inc_ptr.p, total);
        i64 t = 0;
        TRY(scan_exclusive_i32(ctx, ctx->inc_ptr.p, ctx->inc_ptr.p, nq, &t));
        LAUNCH(ctx, k_fill_inc, div_up(total, 256), 256, 0, (const int*)ctx->cq.p, (const int*)ctx->inc_ptr.p, cursor.p, ctx->inc.p, total);
        LAUNCH(ctx, k_sort_inc, div_up(nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, ctx->inc.p, nq);
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CU(ctx->blk_ptr.alloc(nq + 1));
    CU(ctx->diag_slot.alloc(nq));
    CU(cudaMemsetAsync(ctx->errflag.p + 3, 0, sizeof(int), ctx->stream));
    if (ctx->npc == 4)
        LAUNCH(ctx, (k_adjacency<4, false>), div_up(nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p,
               ctx->blk_ptr.p, (const int*)nullptr, (int*)nullptr, (int*)nullptr, nq, ctx->errflag.p + 3);
    else
        LAUNCH(ctx, (k_adjacency<8, false>), div_up(nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p,
               ctx->blk_ptr.p, (const int*)nullptr, (int*)nullptr, (int*)nullptr, nq, ctx->errflag.p + 3);
    CU(cudaMemcpyAsync(&ctx->max_deg, ctx->errflag.p + 3, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    i64 nnzb = 0;
    TRY(scan_exclusive_i32(ctx, ctx->blk_ptr.p, ctx->blk_ptr.p, nq, &nnzb));
    if (nnzb * 9 > 2147483647LL * 4) return toe_fail(ctx, TOE_ERR_MESH, "pattern too large: %lld blocks", (long long)nnzb);
    ctx->nnzb = nnzb;
    ctx->ldv = (nnzb + 15) / 16 * 16;       // plane stride: keeps every plane 128-byte aligned for the bulk copies of the SpMV
    CU(ctx->blk_col.alloc(nnzb + 16));
    CU(cudaMemsetAsync(ctx->blk_col.p + nnzb, 0, 16 * sizeof(int), ctx->stream));
    if (ctx->npc == 4)
        LAUNCH(ctx, (k_adjacency<4, true>), div_up(nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p,
               (int*)nullptr, (const int*)ctx->blk_ptr.p, ctx->blk_col.p, ctx->diag_slot.p, nq, (int*)nullptr);
    else
        LAUNCH(ctx, (k_adjacency<8, true>), div_up(nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p,
               (int*)nullptr, (const int*)ctx->blk_ptr.p, ctx->blk_col.p, ctx->diag_slot.p, nq, (int*)nullptr);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_pattern = true;
    ctx->have_contrib = false; ctx->have_K = false;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// block -> contributing (e,a,b), off-diagonal blocks only, ascending e.  One thread per block slot walks the
// (sorted) incidence list of its row node and keeps the cells that also hold the column node.
// ---------------------------------------------------------------------------------------------------------
template <int NPC, bool FILL>
__global__ void k_contrib(const int* __restrict__ blk_ptr, const int* __restrict__ blk_col, const int* __restrict__ inc_ptr,
                          const int* __restrict__ inc, const int* __restrict__ cq, int* __restrict__ cnt,
                          const int* __restrict__ ctr_ptr, int* __restrict__ ctr, unsigned short* __restrict__ rctr, int nq) {
    // thread per (row q, k-th block of the row): rows are short, so a flat loop over slots with a row lookup
    // would need a search; instead a warp-strided loop inside the row keeps it simple.
    int q = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);     // 8 lanes per row
    if (q >= nq) return;
    int sub = threadIdx.x & 7;
    int lo = inc_ptr[q], hi = inc_ptr[q + 1];
    for (int s = blk_ptr[q] + sub; s < blk_ptr[q + 1]; s += 8) {
        int col = blk_col[s];
        if (col == q) { if (!FILL) cnt[s] = 0; continue; }
        int n = 0;
        int base = FILL ? ctr_ptr[s] : 0;
        for (int k = lo; k < hi; k++) {
            int ea = inc[k];
            int e = ea / NPC, a = ea - e * NPC;
            const int* c = cq + (size_t)e * NPC;
#pragma unroll
            for (int b = 0; b < NPC; b++)
                if (c[b] == col) {
                    if (FILL) {
                        ctr[base + n] = ctr_pack<NPC>(e, a, b);
                        if (rctr) rctr[base + n] = (unsigned short)(((k - lo) << 4) | (a << 2) | b);     // row-relative form (Tet4, < 4096 cells per node)
                    }
                    n++;
                }
        }
        if (!FILL) cnt[s] = n;
    }
}

__global__ void k_max_inc(const int* __restrict__ inc_ptr, int nq, int* out) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int v = q < nq ? inc_ptr[q + 1] - inc_ptr[q] : 0;
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}

int mesh_build_contrib(toe_ctx* ctx) {
    if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "contribution lists need the pattern");
    if (ctx->have_contrib) return TOE_OK;
    if (ctx->ne > (ctx->npc == 4 ? (1LL << 27) : (1LL << 25)))
        return toe_fail(ctx, TOE_ERR_MESH, "gather assembly packs (cell,a,b) into 32 bits: ne=%lld is too large", (long long)ctx->ne);
    int nq = ctx->nq;
    CU(ctx->ctr_ptr.alloc(ctx->nnzb + 1));
    unsigned grid = div_up(nq, 16);
#define ARGS(cntp, ptrp, ctrp) (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, \
        (const int*)ctx->cq.p, cntp, ptrp, ctrp, rctrp, nq
    unsigned short* rctrp = nullptr;
    if (ctx->npc == 4) LAUNCH(ctx, (k_contrib<4, false>), grid, 128, 0, ARGS(ctx->ctr_ptr.p, (const int*)nullptr, (int*)nullptr));
    else               LAUNCH(ctx, (k_contrib<8, false>), grid, 128, 0, ARGS(ctx->ctr_ptr.p, (const int*)nullptr, (int*)nullptr));
    i64 total = 0;
    TRY(scan_exclusive_i32(ctx, ctx->ctr_ptr.p, ctx->ctr_ptr.p, ctx->nnzb, &total));
    CU(ctx->ctr.alloc(total));
    // largest incidence list: the row-relative lists hold the list position in 12 bits
    CU(cudaMemsetAsync(ctx->errflag.p + 3, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_max_inc, div_up(nq, 256), 256, 0, (const int*)ctx->inc_ptr.p, nq, ctx->errflag.p + 3);
    CU(cudaMemcpyAsync(&ctx->max_inc, ctx->errflag.p + 3, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_rctr = (ctx->npc == 4 && ctx->max_inc < 4096);
    if (ctx->have_rctr) { CU(ctx->rctr.alloc(total)); rctrp = ctx->rctr.p; }
    if (ctx->npc == 4) LAUNCH(ctx, (k_contrib<4, true>), grid, 128, 0, ARGS((int*)nullptr, (const int*)ctx->ctr_ptr.p, ctx->ctr.p));
    else               LAUNCH(ctx, (k_contrib<8, true>), grid, 128, 0, ARGS((int*)nullptr, (const int*)ctx->ctr_ptr.p, ctx->ctr.p));
#undef ARGS
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_contrib = true;
    return TOE_OK;
}
