// ebe_tile.cu — matrix-free element-by-element operator, tile form (north-star kernel 3b).
//
//   y = Σ_e Pₑᵀ Kₑ Pₑ x   with every cell computed ONCE, Kₑ never formed:
//   H = Σ_b x_b ⊗ g_b,  σ = λ tr(ε) I + 2μ ε,  (Kₑxₑ)_a = w σ g_a       (constitutive_relation, FiniteElementAnalysis.jl:126-129)
//
// Cells are cut into tiles of 1024/npc consecutive cells (256 tets, 128 hexes).  Per tile, built once per mesh on the GPU
// (`k_tile_build`, one bitonic sort of ≤1024 (node, ref) keys in shared memory): the sorted list of its local nodes, the
// local connectivity (uint16) and, for every local node, the refs (cell, corner) that touch it.
//   k_ebe_tile   CTA per tile: gathers coordinates + x of the tile's nodes into shared memory once, one thread per cell
//                computes Kₑxₑ from the gradients in registers into a shared scratch, one thread per local node adds up
//                its refs in a fixed order and writes 3 doubles to the tile's rows of a staging array (coalesced).
//   k_ebe_nodes  thread per node: sums its ≤ (#tiles touching it) staging rows in ascending order, applies the
//                constrained rows (m·x on prescribed dofs), stores y and folds the CG dot product p'Ap.
// No atomics anywhere: bit-reproducible.  Algorithmic traffic ≈ 36 B/cell (conn + coordinates + x + y); the staging
// round trip adds ≈ 33 B/cell.  Compared with the node-gather form (k_ebe_gather, solver.cu) the geometry and the Hooke
// product are evaluated once per cell instead of once per (cell, corner).
#include "element.cuh"
#include <climits>

static const int TILE_REFS = 1024;          // refs (cell corners) per tile

// ---------------------------------------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------------------------------------
// sorts the (node<<10 | ref) keys of one tile, emits: local node list (into a ≤1024-per-tile scratch), count, local
// connectivity and the node-sorted ref ids + first ref per local node
template <int NPC>
__global__ void __launch_bounds__(256) k_tile_build(const int* __restrict__ cq, i64 ne, int* __restrict__ tile_m, int* __restrict__ nodes_tmp,
                                                    unsigned short* __restrict__ lconn, unsigned short* __restrict__ inc_sorted,
                                                    unsigned short* __restrict__ nstart_tmp) {
    const int TE = TILE_REFS / NPC;
    __shared__ u64 key[TILE_REFS];
    __shared__ int uniq[TILE_REFS];          // first: flags / scan, then: unique node ids
    __shared__ int wsum[8];
    __shared__ int total;
    const int t = blockIdx.x, tid = threadIdx.x;
    const i64 e0 = (i64)t * TE;
    const int nel = (int)min((i64)TE, ne - e0);
    const int nref = nel * NPC;
    for (int i = tid; i < TILE_REFS; i += 256)
        key[i] = i < nref ? (((u64)(unsigned)cq[e0 * NPC + i]) << 10) | (u64)i : ~0ULL;
    __syncthreads();
    for (int k = 2; k <= TILE_REFS; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < TILE_REFS; i += 256) {
                int p = i ^ j;
                if (p > i) {
                    u64 a = key[i], b = key[p];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { key[i] = b; key[p] = a; }
                }
            }
            __syncthreads();
        }
    // flags of first occurrences; each thread owns 4 consecutive sorted positions
    int f[4], s = 0;
    for (int k = 0; k < 4; k++) {
        int i = 4 * tid + k;
        bool first = i < nref && (i == 0 || (key[i] >> 10) != (key[i - 1] >> 10));
        f[k] = first ? 1 : 0; s += f[k];
    }
    int lane = tid & 31, w = tid >> 5, incl = s;
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; k++) woff += wsum[k];
    if (tid == 255) total = woff + incl;
    int run = woff + incl - s;                 // exclusive rank of this thread's first flag
    int lidx[4];
    for (int k = 0; k < 4; k++) { run += f[k]; lidx[k] = run - 1; }     // local node index of sorted position 4*tid+k
    __syncthreads();
    const int m = total;
    for (int k = 0; k < 4; k++) {
        int i = 4 * tid + k;
        if (i >= nref) continue;
        int node = (int)(key[i] >> 10), ref = (int)(key[i] & 1023);
        if (f[k]) { uniq[lidx[k]] = node; nstart_tmp[(size_t)t * TILE_REFS + lidx[k]] = (unsigned short)i; }
        lconn[e0 * NPC + ref] = (unsigned short)lidx[k];
        inc_sorted[e0 * NPC + i] = (unsigned short)ref;
    }
    __syncthreads();
    for (int j = tid; j < m; j += 256) nodes_tmp[(size_t)t * TILE_REFS + j] = uniq[j];
    if (tid == 0) tile_m[t] = m;
}

__global__ void k_tile_compact(const int* __restrict__ tile_off, const int* __restrict__ nodes_tmp, const unsigned short* __restrict__ nstart_tmp,
                               int* __restrict__ tile_nodes, unsigned short* __restrict__ tile_nstart, int* __restrict__ cnt, int* max_m) {
    const int t = blockIdx.x;
    const int off = tile_off[t], m = tile_off[t + 1] - off;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        int q = nodes_tmp[(size_t)t * TILE_REFS + j];
        tile_nodes[off + j] = q;
        tile_nstart[off + j] = nstart_tmp[(size_t)t * TILE_REFS + j];
        atomicAdd(&cnt[q], 1);
    }
    if (threadIdx.x == 0) atomicMax(max_m, m);
}

__global__ void k_nst_fill(const int* __restrict__ tile_nodes, const int* __restrict__ nst_ptr, int* cursor, int* __restrict__ nst, i64 nslots) {
    i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    int q = tile_nodes[s];
    nst[nst_ptr[q] + atomicAdd(&cursor[q], 1)] = (int)s;
}
__global__ void k_nst_sort(const int* __restrict__ nst_ptr, int* __restrict__ nst, int nq) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int lo = nst_ptr[q], hi = nst_ptr[q + 1];
    for (int i = lo + 1; i < hi; i++) { int v = nst[i], j = i - 1; while (j >= lo && nst[j] > v) { nst[j + 1] = nst[j]; j--; } nst[j + 1] = v; }
}

int mesh_build_tiles(toe_ctx* ctx) {
    if (ctx->have_tiles) return TOE_OK;
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "matrix-free operator: DOFs not built");
    const int npc = ctx->npc, TE = TILE_REFS / npc;
    const i64 ne = ctx->ne;
    const int ntiles = (int)((ne + TE - 1) / TE), nq = ctx->nq;
    CU(ctx->tile_off.alloc(ntiles + 1));
    CU(ctx->tile_lconn.alloc(ne * npc)); CU(ctx->tile_inc.alloc(ne * npc));
    DevBuf<int> nodes_tmp; DevBuf<unsigned short> nstart_tmp;
    CU(nodes_tmp.alloc((size_t)ntiles * TILE_REFS)); CU(nstart_tmp.alloc((size_t)ntiles * TILE_REFS));
    if (npc == 4) LAUNCH(ctx, k_tile_build<4>, ntiles, 256, 0, (const int*)ctx->cq.p, ne, ctx->tile_off.p, nodes_tmp.p, ctx->tile_lconn.p, ctx->tile_inc.p, nstart_tmp.p);
    else          LAUNCH(ctx, k_tile_build<8>, ntiles, 256, 0, (const int*)ctx->cq.p, ne, ctx->tile_off.p, nodes_tmp.p, ctx->tile_lconn.p, ctx->tile_inc.p, nstart_tmp.p);
    i64 nslots = 0;
    TRY(scan_exclusive_i32(ctx, ctx->tile_off.p, ctx->tile_off.p, ntiles, &nslots));
    CU(ctx->tile_nodes.alloc(nslots)); CU(ctx->tile_nstart.alloc(nslots));
    CU(ctx->nst_ptr.alloc(nq + 1)); CU(ctx->nst.alloc(nslots));
    CU(ctx->tile_stage.alloc(3 * (size_t)nslots));
    DevBuf<int> cursor; CU(cursor.alloc(nq + 1));
    CU(cudaMemsetAsync(ctx->nst_ptr.p, 0, (nq + 1) * sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(cursor.p, 0, (nq + 1) * sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(ctx->errflag.p + 3, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_tile_compact, ntiles, 128, 0, (const int*)ctx->tile_off.p, (const int*)nodes_tmp.p, (const unsigned short*)nstart_tmp.p,
           ctx->tile_nodes.p, ctx->tile_nstart.p, ctx->nst_ptr.p, ctx->errflag.p + 3);
    int max_m = 0;
    CU(cudaMemcpyAsync(&max_m, ctx->errflag.p + 3, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    i64 tot = 0;
    TRY(scan_exclusive_i32(ctx, ctx->nst_ptr.p, ctx->nst_ptr.p, nq, &tot));
    LAUNCH(ctx, k_nst_fill, div_up(nslots, 256), 256, 0, (const int*)ctx->tile_nodes.p, (const int*)ctx->nst_ptr.p, cursor.p, ctx->nst.p, nslots);
    LAUNCH(ctx, k_nst_sort, div_up(nq, 128), 128, 0, (const int*)ctx->nst_ptr.p, ctx->nst.p, nq);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ntiles = ntiles; ctx->tile_elems = TE; ctx->tile_max_nodes = max_m; ctx->tile_slots = nslots;
    ctx->have_tiles = true;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// operator
// ---------------------------------------------------------------------------------------------------------
template <int NPC, bool MASK>
__global__ void __launch_bounds__(TILE_REFS / NPC) k_ebe_tile(const int* __restrict__ tile_off, const int* __restrict__ tile_nodes,
                                                              const unsigned short* __restrict__ lconn, const unsigned short* __restrict__ inc_sorted,
                                                              const unsigned short* __restrict__ nstart, const double* __restrict__ xq, Material mat,
                                                              const unsigned char* __restrict__ dflag, const double* __restrict__ x,
                                                              double* __restrict__ stage, i64 ne, const int* done_flag) {
    const int TE = TILE_REFS / NPC;
    TOE_DYN_SMEM(double, sm, 16);
    if (done_flag && *done_flag) return;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int off = __ldg(&tile_off[t]), m = __ldg(&tile_off[t + 1]) - off;
    const i64 e0 = (i64)t * TE;
    const int nel = (int)min((i64)TE, ne - e0);
    const int nref = nel * NPC;
    double* scratch = sm;                        // [TILE_REFS][3]: (Kₑxₑ)_a per ref
    double* Xs = sm + 3 * TILE_REFS;             // [m][3] coordinates
    double* xs = Xs + 3 * m;                     // [m][3] x
    unsigned short* s_inc = reinterpret_cast<unsigned short*>(xs + 3 * m);      // [TILE_REFS] node-sorted refs
    unsigned short* s_nst = s_inc + TILE_REFS;                                  // [m] first ref of each local node
    // every global load of the tile is issued before the first barrier: metadata, connectivity, material, node data
    uint2 lc2 = make_uint2(0, 0); uint4 lc4 = make_uint4(0, 0, 0, 0);
    double lam = 0.0, mu = 0.0;
    if (tid < nel) {
        if (NPC == 4) lc2 = __ldg(reinterpret_cast<const uint2*>(lconn + (size_t)(e0 + tid) * 4));
        else          lc4 = __ldg(reinterpret_cast<const uint4*>(lconn + (size_t)(e0 + tid) * 8));
        material_at(mat, (int)(e0 + tid), lam, mu);
    }
    {
        const unsigned int* src = reinterpret_cast<const unsigned int*>(inc_sorted + e0 * NPC);    // e0*NPC is a multiple of 1024 → 4-byte aligned
        unsigned int* dst = reinterpret_cast<unsigned int*>(s_inc);
        for (int i = tid; i < (nref + 1) / 2; i += TE) dst[i] = __ldg(src + i);
    }
    for (int j = tid; j < m; j += TE) {
        s_nst[j] = __ldg(&nstart[off + j]);
        int q = __ldg(&tile_nodes[off + j]);
        double c[3], v[3];
        load3(xq, q, c); load3(x, q, v);
        if (MASK) {
#pragma unroll
            for (int k = 0; k < 3; k++) if (dflag[3 * (size_t)q + k]) v[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) { Xs[3 * j + k] = c[k]; xs[3 * j + k] = v[k]; }
    }
    __syncthreads();
    if (tid < nel) {
        int l[NPC];
        if (NPC == 4) {
            uint2 p = lc2;
            l[0] = p.x & 0xffff; l[1] = p.x >> 16; l[2] = p.y & 0xffff; l[3] = p.y >> 16;
            double X[4][3], g[4][3];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int k = 0; k < 3; k++) X[a][k] = Xs[3 * l[a] + k];
            double det = tet_grads(X, g);
            double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) {
                    double xv = xs[3 * l[b] + c2];
#pragma unroll
                    for (int i2 = 0; i2 < 3; i2++) H[c2][i2] += xv * g[b][i2];
                }
            double S[3][3]; hooke_from_grad(H, lam, mu, S);
            double w = det * (1.0 / 6.0);
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++)
                    scratch[3 * (4 * tid + a) + c2] = w * (S[c2][0] * g[a][0] + S[c2][1] * g[a][1] + S[c2][2] * g[a][2]);
        } else {
            uint4 p = lc4;
            l[0] = p.x & 0xffff; l[1] = p.x >> 16; l[2] = p.y & 0xffff; l[3] = p.y >> 16;
            l[4] = p.z & 0xffff; l[5] = p.z >> 16; l[6] = p.w & 0xffff; l[7] = p.w >> 16;
            double X[8][3], xe[8][3], out[8][3];
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int k = 0; k < 3; k++) { X[a][k] = Xs[3 * l[a] + k]; xe[a][k] = xs[3 * l[a] + k]; out[a][k] = 0.0; }
            for (int gp = 0; gp < 8; gp++) {
                double g[8][3], N[8];
                double det = hex_grads_at(X, gp, g, N);
                double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
                for (int b = 0; b < 8; b++)
#pragma unroll
                    for (int c2 = 0; c2 < 3; c2++)
#pragma unroll
                        for (int i2 = 0; i2 < 3; i2++) H[c2][i2] += xe[b][c2] * g[b][i2];
                double S[3][3]; hooke_from_grad(H, lam, mu, S);
#pragma unroll
                for (int a = 0; a < 8; a++)
#pragma unroll
                    for (int c2 = 0; c2 < 3; c2++) out[a][c2] += det * (S[c2][0] * g[a][0] + S[c2][1] * g[a][1] + S[c2][2] * g[a][2]);
            }
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) scratch[3 * (8 * tid + a) + c2] = out[a][c2];
        }
    }
    __syncthreads();
    for (int j = tid; j < m; j += TE) {
        int lo = s_nst[j];
        int hi = j + 1 < m ? (int)s_nst[j + 1] : nref;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int p = lo; p < hi; p++) {                          // refs of this node in ascending (cell, corner) order
            int ref = s_inc[p];
            s0 += scratch[3 * ref]; s1 += scratch[3 * ref + 1]; s2 += scratch[3 * ref + 2];
        }
        double* o = stage + 3 * (size_t)(off + j);
        o[0] = s0; o[1] = s1; o[2] = s2;
    }
}

template <bool CG>
__global__ void __launch_bounds__(256) k_ebe_nodes(const int* __restrict__ nst_ptr, const int* __restrict__ nst, const double* __restrict__ stage,
                                                   const unsigned char* __restrict__ dflag, const double* __restrict__ dval, int any_dirichlet,
                                                   const unsigned char* __restrict__ owned, const double* __restrict__ x, double* __restrict__ y, int nq,
                                                   const int* done_flag, CGScalars* cg, double* partials, unsigned int* counter, double* dot_out) {
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    double dotv = 0.0;
    if (q < nq) {
        double ya[3] = {0, 0, 0};
        for (int i = __ldg(&nst_ptr[q]); i < __ldg(&nst_ptr[q + 1]); i++) {
            const double* s = stage + 3 * (size_t)__ldg(&nst[i]);
            ya[0] += s[0]; ya[1] += s[1]; ya[2] += s[2];
        }
        double xs[3]; load3(x, q, xs);
        if (any_dirichlet) {
#pragma unroll
            for (int c = 0; c < 3; c++) { size_t d = 3 * (size_t)q + c; if (dflag[d]) ya[c] = (!owned || owned[q]) ? dval[d] * xs[c] : 0.0; }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) y[3 * (size_t)q + c] = ya[c];
        dotv = ya[0] * xs[0] + ya[1] * xs[1] + ya[2] * xs[2];
    }
    if (CG) {
        double d = block_sum(dotv, red);
        double tot;
        if (grid_sum_last_block(d, partials, counter, red, &tot)) { if (dot_out) *dot_out = tot; else cg_after_pAp(cg, tot); }
    }
}

int ebe_tile_launch(toe_ctx* ctx, const double* x, double* y, CGScalars* cg, bool mask, const int* done_flag, double* dot_out) {
    TRY(mesh_build_tiles(ctx));
    const int npc = ctx->npc;
    size_t smem = (3 * (size_t)TILE_REFS + 6 * (size_t)ctx->tile_max_nodes) * sizeof(double) + (TILE_REFS + (size_t)ctx->tile_max_nodes + 8) * sizeof(unsigned short);
    static size_t attr_smem[2] = {0, 0};
    if (smem > 48 * 1024 && smem > attr_smem[npc == 8]) {
        CU(cudaFuncSetAttribute(k_ebe_tile<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_ebe_tile<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_ebe_tile<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_ebe_tile<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem[npc == 8] = smem;
    }
#define TILE_ARGS (const int*)ctx->tile_off.p, (const int*)ctx->tile_nodes.p, (const unsigned short*)ctx->tile_lconn.p, (const unsigned short*)ctx->tile_inc.p, \
        (const unsigned short*)ctx->tile_nstart.p, (const double*)ctx->xq.p, ctx->mat, (const unsigned char*)ctx->dflag.p, x, ctx->tile_stage.p, ctx->ne, done_flag
    if (npc == 4) { if (mask) LAUNCH(ctx, (k_ebe_tile<4, true>), ctx->ntiles, 256, smem, TILE_ARGS); else LAUNCH(ctx, (k_ebe_tile<4, false>), ctx->ntiles, 256, smem, TILE_ARGS); }
    else          { if (mask) LAUNCH(ctx, (k_ebe_tile<8, true>), ctx->ntiles, 128, smem, TILE_ARGS); else LAUNCH(ctx, (k_ebe_tile<8, false>), ctx->ntiles, 128, smem, TILE_ARGS); }
#undef TILE_ARGS
    unsigned grid = div_up(ctx->nq, 256);
#define NODE_ARGS (const int*)ctx->nst_ptr.p, (const int*)ctx->nst.p, (const double*)ctx->tile_stage.p, (const unsigned char*)ctx->dflag.p, (const double*)ctx->dval.p, \
        (int)ctx->any_dirichlet, ctx->owned, x, y, ctx->nq, done_flag, cg, ctx->partials.p, ctx->counters.p + 1, dot_out
    if (cg) LAUNCH(ctx, k_ebe_nodes<true>, grid, 256, 0, NODE_ARGS);
    else    LAUNCH(ctx, k_ebe_nodes<false>, grid, 256, 0, NODE_ARGS);
#undef NODE_ARGS
    return TOE_OK;
}
