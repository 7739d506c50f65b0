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
#include <cstdlib>
#include <cstdio>

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

__global__ void k_tile_pad(const int* __restrict__ tile_m, int* __restrict__ padded, int ntiles) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles) padded[t] = (tile_m[t] + 7) & ~7;       // 8-row granularity: every tile's slice of the metadata arrays is 16-byte aligned
}

__global__ void k_tile_compact(const int* __restrict__ tile_off, const int* __restrict__ tile_m, const int* __restrict__ nodes_tmp,
                               const unsigned short* __restrict__ nstart_tmp,
                               int* __restrict__ tile_nodes, unsigned short* __restrict__ tile_nstart, int* __restrict__ cnt, int* max_m) {
    const int t = blockIdx.x;
    const int off = tile_off[t], m = tile_m[t], mp = tile_off[t + 1] - off;
    for (int j = threadIdx.x; j < mp; j += blockDim.x) {
        if (j < m) {
            int q = nodes_tmp[(size_t)t * TILE_REFS + j];
            tile_nodes[off + j] = q;
            tile_nstart[off + j] = nstart_tmp[(size_t)t * TILE_REFS + j];
            atomicAdd(&cnt[q], 1);
        } else {
            tile_nodes[off + j] = -1;                       // padding row
            tile_nstart[off + j] = 0;
        }
    }
    if (threadIdx.x == 0) atomicMax(max_m, m);
}

__global__ void k_nst_fill(const int* __restrict__ tile_nodes, const int* __restrict__ nst_ptr, int* cursor, int* __restrict__ nst, i64 nslots) {
    i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    int q = tile_nodes[s];
    if (q < 0) return;
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
    CU(ctx->tile_off.alloc(ntiles + 1)); CU(ctx->tile_m.alloc(ntiles + 1));
    CU(ctx->tile_lconn.alloc(ne * npc + 64)); CU(ctx->tile_inc.alloc(ne * npc + 64));      // slack: the pipelined kernel copies in 16-byte units
    DevBuf<int> nodes_tmp; DevBuf<unsigned short> nstart_tmp;
    CU(nodes_tmp.alloc((size_t)ntiles * TILE_REFS)); CU(nstart_tmp.alloc((size_t)ntiles * TILE_REFS));
    if (npc == 4) LAUNCH(ctx, k_tile_build<4>, ntiles, 256, 0, (const int*)ctx->cq.p, ne, ctx->tile_m.p, nodes_tmp.p, ctx->tile_lconn.p, ctx->tile_inc.p, nstart_tmp.p);
    else          LAUNCH(ctx, k_tile_build<8>, ntiles, 256, 0, (const int*)ctx->cq.p, ne, ctx->tile_m.p, nodes_tmp.p, ctx->tile_lconn.p, ctx->tile_inc.p, nstart_tmp.p);
    LAUNCH(ctx, k_tile_pad, div_up(ntiles, 256), 256, 0, (const int*)ctx->tile_m.p, ctx->tile_off.p, ntiles);
    i64 nslots = 0;
    TRY(scan_exclusive_i32(ctx, ctx->tile_off.p, ctx->tile_off.p, ntiles, &nslots));
    CU(ctx->tile_nodes.alloc(nslots + 8)); CU(ctx->tile_nstart.alloc(nslots + 8));
    CU(ctx->nst_ptr.alloc(nq + 1)); CU(ctx->nst.alloc(nslots));
    CU(ctx->tile_stage.alloc(3 * (size_t)nslots));
    DevBuf<int> cursor; CU(cursor.alloc(nq + 1));
    CU(cudaMemsetAsync(ctx->nst_ptr.p, 0, (nq + 1) * sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(cursor.p, 0, (nq + 1) * sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(ctx->errflag.p + 3, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_tile_compact, ntiles, 128, 0, (const int*)ctx->tile_off.p, (const int*)ctx->tile_m.p, (const int*)nodes_tmp.p, (const unsigned short*)nstart_tmp.p,
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
__global__ void __launch_bounds__(TILE_REFS / NPC, NPC == 4 ? 4 : 1) k_ebe_tile(const int* __restrict__ tile_off, const int* __restrict__ tile_m, const int* __restrict__ tile_nodes,
                                                              const unsigned short* __restrict__ lconn, const unsigned short* __restrict__ inc_sorted,
                                                              const unsigned short* __restrict__ nstart, const double* __restrict__ xq, Material mat,
                                                              const unsigned char* __restrict__ dflag, const double* __restrict__ x,
                                                              double* __restrict__ stage, i64 ne, const int* done_flag) {
    const int TE = TILE_REFS / NPC;
    TOE_DYN_SMEM(double, sm, 16);
    if (done_flag && *done_flag) return;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int off = __ldg(&tile_off[t]), m = __ldg(&tile_m[t]);
    const i64 e0 = (i64)t * TE;
    const int nel = (int)min((i64)TE, ne - e0);
    const int nref = nel * NPC;
    // Shared-memory layout chosen for the bank structure (round-1 ncu: 62 % of this kernel's shared-memory wavefronts were bank
    // conflicts and l1tex ran at 87 % of peak — the kernel was shared-memory-bandwidth bound):
    //   sA[a][cell] double2 = first two components of (Kₑxₑ)_a, sB[a][cell] = the third: a thread's stores go to consecutive
    //                addresses of consecutive threads (the old [ref][3] layout had a 96-byte stride = 4-way conflicts);
    //   nd[node]     6 doubles (X0 X1 X2 x0 x1 x2) = three 16-byte loads per corner instead of six 8-byte ones; the 48-byte stride
    //                maps 8 consecutive nodes onto the 8 distinct 16-byte bank groups.
    double2* sA = reinterpret_cast<double2*>(sm);                               // [NPC][TE]
    double* sB = sm + 2 * TILE_REFS;                                            // [NPC][TE]
    double* nd = sm + 3 * TILE_REFS;                                            // [m][6]
    unsigned short* s_inc = reinterpret_cast<unsigned short*>(nd + 6 * m);      // [TILE_REFS] node-sorted refs
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
        double2* o = reinterpret_cast<double2*>(nd + 6 * j);
        o[0] = make_double2(c[0], c[1]); o[1] = make_double2(c[2], v[0]); o[2] = make_double2(v[1], v[2]);
    }
    __syncthreads();
    const double2* nd2 = reinterpret_cast<const double2*>(nd);
    if (tid < nel) {
        int l[NPC];
        if (NPC == 4) {
            uint2 p = lc2;
            l[0] = p.x & 0xffff; l[1] = p.x >> 16; l[2] = p.y & 0xffff; l[3] = p.y >> 16;
            double X[4][3], xv[4][3], g[4][3];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const double2 p0 = nd2[3 * l[a]], p1 = nd2[3 * l[a] + 1], p2 = nd2[3 * l[a] + 2];
                X[a][0] = p0.x; X[a][1] = p0.y; X[a][2] = p1.x; xv[a][0] = p1.y; xv[a][1] = p2.x; xv[a][2] = p2.y;
            }
            double det = tet_grads(X, g);
            double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) {
#pragma unroll
                    for (int i2 = 0; i2 < 3; i2++) H[c2][i2] += xv[b][c2] * g[b][i2];
                }
            double S[3][3]; hooke_from_grad(H, lam, mu, S);
            double w = det * (1.0 / 6.0);
#pragma unroll
            for (int a = 0; a < 4; a++) {
                double r[3];
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) r[c2] = w * (S[c2][0] * g[a][0] + S[c2][1] * g[a][1] + S[c2][2] * g[a][2]);
                sA[a * TE + tid] = make_double2(r[0], r[1]); sB[a * TE + tid] = r[2];
            }
        } else {
            uint4 p = lc4;
            l[0] = p.x & 0xffff; l[1] = p.x >> 16; l[2] = p.y & 0xffff; l[3] = p.y >> 16;
            l[4] = p.z & 0xffff; l[5] = p.z >> 16; l[6] = p.w & 0xffff; l[7] = p.w >> 16;
            double X[8][3], xe[8][3], out[8][3];
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const double2 p0 = nd2[3 * l[a]], p1 = nd2[3 * l[a] + 1], p2 = nd2[3 * l[a] + 2];
                X[a][0] = p0.x; X[a][1] = p0.y; X[a][2] = p1.x; xe[a][0] = p1.y; xe[a][1] = p2.x; xe[a][2] = p2.y;
                out[a][0] = out[a][1] = out[a][2] = 0.0;
            }
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
            for (int a = 0; a < 8; a++) { sA[a * TE + tid] = make_double2(out[a][0], out[a][1]); sB[a * TE + tid] = out[a][2]; }
        }
    }
    __syncthreads();
    for (int j = tid; j < m; j += TE) {
        int lo = s_nst[j];
        int hi = j + 1 < m ? (int)s_nst[j + 1] : nref;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int p = lo; p < hi; p++) {                          // refs of this node in ascending (cell, corner) order
            const int ref = s_inc[p];
            const int slot = (ref & (NPC - 1)) * TE + (ref / NPC);
            const double2 v = sA[slot];
            s0 += v.x; s1 += v.y; s2 += sB[slot];
        }
        double* o = stage + 3 * (size_t)(off + j);
        o[0] = s0; o[1] = s1; o[2] = s2;
    }
}

// ---------------------------------------------------------------------------------------------------------
// operator, pipelined form (opt-in: TOE_EBE_PIPE=1; Tet4 and Hex8, no column masking)
//
// k_ebe_tile spends most of a tile's life waiting: new CTA → tile_off → tile_nodes → x/xq is a chain of three dependent
// DRAM/L2 latencies before the first FMA (ncu, r1: FP64 pipe 22 %, no dominant stall).  Here a PERSISTENT CTA walks tiles
// t = blockIdx.x, += gridDim.x with a three-deep software pipeline of asynchronous copies (cp.async → SASS LDGSTS):
//     iteration i:   wait for {node data of tile i, metadata of tile i+1}            cp.async.wait_all + barrier
//                    issue  {metadata of tile i+2 (contiguous, 16-byte copies),
//                            node gather of tile i+1 (8-byte copies, addresses from its metadata in shared memory)}
//                    compute tile i from shared memory (same arithmetic and summation order as k_ebe_tile → same bits)
// so every global latency is hidden behind a whole tile of compute.  Metadata slices are 16-byte aligned because every
// tile's rows are padded to a multiple of 8 (mesh_build_tiles).
// ---------------------------------------------------------------------------------------------------------
#ifndef TOE_EMU
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else
static inline void cp_async_16(void* dst, const void* src) { emu::cp_async(dst, src, 16); }
static inline void cp_async_8(void* dst, const void* src) { emu::cp_async(dst, src, 8); }
static inline void cp_async_commit() {}
static inline void cp_async_wait_all() { emu::cp_async_wait_all(); }
#endif

// shared-memory plan of k_ebe_pipe (bytes; every piece 16-byte aligned).  mp = padded maximum of local nodes per tile.
struct PipeLayout {
    size_t scratch, data_stride, meta_stride, nodes_off, nst_off, lc_off, inc_off, total;
    __host__ __device__ explicit PipeLayout(int mp) {
        scratch = (size_t)3 * TILE_REFS * sizeof(double);
        data_stride = (size_t)6 * mp * sizeof(double);                         // Xs[3mp] + xs[3mp]
        nodes_off = 0;
        nst_off = (size_t)mp * sizeof(int);
        lc_off = nst_off + (((size_t)mp * sizeof(unsigned short) + 15) & ~(size_t)15);
        inc_off = lc_off + (size_t)TILE_REFS * sizeof(unsigned short);
        meta_stride = inc_off + (size_t)TILE_REFS * sizeof(unsigned short);
        total = scratch + 2 * data_stride + 3 * meta_stride;
    }
};

template <int NPC>
__global__ void __launch_bounds__(TILE_REFS / NPC) k_ebe_pipe(const int* __restrict__ tile_off, const int* __restrict__ tile_m, const int* __restrict__ tile_nodes,
                                                              const unsigned short* __restrict__ lconn, const unsigned short* __restrict__ inc_sorted,
                                                              const unsigned short* __restrict__ nstart, const double* __restrict__ xq, Material mat,
                                                              const double* __restrict__ x, double* __restrict__ stage, i64 ne, int ntiles, int mp,
                                                              const int* done_flag) {
    const int TE = TILE_REFS / NPC;
    TOE_DYN_SMEM(unsigned char, smraw, 16);
    if (done_flag && *done_flag) return;
    const PipeLayout L(mp);
    const int tid = threadIdx.x;
    double* scratch = reinterpret_cast<double*>(smraw);
    unsigned char* data0 = smraw + L.scratch;
    unsigned char* meta0 = data0 + 2 * L.data_stride;

    // scalars of the tiles in flight: this one, next, next-but-one (plain loads, always one iteration ahead of their use)
    int t0 = blockIdx.x, t1 = t0 + gridDim.x, t2 = t1 + gridDim.x;
    int off0 = 0, m0 = 0, off1 = 0, m1 = 0, off2 = 0, m2 = 0;
    if (t0 < ntiles) { off0 = __ldg(&tile_off[t0]); m0 = __ldg(&tile_m[t0]); }
    if (t1 < ntiles) { off1 = __ldg(&tile_off[t1]); m1 = __ldg(&tile_m[t1]); }
    if (t2 < ntiles) { off2 = __ldg(&tile_off[t2]); m2 = __ldg(&tile_m[t2]); }

    auto issue_meta = [&](int t, int off, int m, int slot) {
        if (t >= ntiles) return;
        unsigned char* M = meta0 + (size_t)slot * L.meta_stride;
        const i64 e0 = (i64)t * TE;
        const int nel = (int)min((i64)TE, ne - e0);
        const int nref = nel * NPC;
        const int n_nodes16 = (m + 3) >> 2, n_nst16 = (m + 7) >> 3, n_ref16 = (nref + 7) >> 3;
        const unsigned char* gn = reinterpret_cast<const unsigned char*>(tile_nodes + off);
        const unsigned char* gs = reinterpret_cast<const unsigned char*>(nstart + off);
        const unsigned char* gl = reinterpret_cast<const unsigned char*>(lconn + e0 * NPC);
        const unsigned char* gi = reinterpret_cast<const unsigned char*>(inc_sorted + e0 * NPC);
        for (int k = tid; k < n_nodes16; k += TE) cp_async_16(M + L.nodes_off + 16 * k, gn + 16 * k);
        for (int k = tid; k < n_nst16; k += TE) cp_async_16(M + L.nst_off + 16 * k, gs + 16 * k);
        for (int k = tid; k < n_ref16; k += TE) { cp_async_16(M + L.lc_off + 16 * k, gl + 16 * k); cp_async_16(M + L.inc_off + 16 * k, gi + 16 * k); }
    };
    auto issue_gather = [&](int t, int m, int mslot, int dslot) {
        if (t >= ntiles) return;
        const int* nodes = reinterpret_cast<const int*>(meta0 + (size_t)mslot * L.meta_stride + L.nodes_off);
        double* Xs = reinterpret_cast<double*>(data0 + (size_t)dslot * L.data_stride);
        double* xs = Xs + 3 * mp;
        for (int j = tid; j < m; j += TE) {
            const int q = nodes[j];
            const double* gc = xq + 3 * (size_t)q;
            const double* gx = x + 3 * (size_t)q;
#pragma unroll
            for (int k = 0; k < 3; k++) { cp_async_8(Xs + 3 * j + k, gc + k); cp_async_8(xs + 3 * j + k, gx + k); }
        }
    };

    // prologue: metadata of the first tile, then its node data together with the metadata of the second
    issue_meta(t0, off0, m0, 0);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    issue_gather(t0, m0, 0, 0);
    issue_meta(t1, off1, m1, 1);
    cp_async_commit();
    double lam = 0.0, mu = 0.0;                       // material of this thread's cell in the current tile (prefetched one tile ahead)
    if (t0 < ntiles) { const i64 e = (i64)t0 * TE + tid; if (e < ne) material_at(mat, (int)e, lam, mu); }

    for (int i = 0; t0 < ntiles; i++) {
        const int ms = i % 3, ds = i & 1;
        cp_async_wait_all();
        __syncthreads();                               // node data of tile i and metadata of tile i+1 are in shared memory, tile i-1 is fully consumed
        issue_meta(t2, off2, m2, (i + 2) % 3);
        issue_gather(t1, m1, (i + 1) % 3, ds ^ 1);
        cp_async_commit();
        const int t3 = t2 + gridDim.x;
        int off3 = 0, m3 = 0;
        if (t3 < ntiles) { off3 = __ldg(&tile_off[t3]); m3 = __ldg(&tile_m[t3]); }
        double lam_n = 0.0, mu_n = 0.0;
        if (t1 < ntiles) { const i64 e = (i64)t1 * TE + tid; if (e < ne) material_at(mat, (int)e, lam_n, mu_n); }

        // ---- compute tile i (t0) ----
        const unsigned char* M = meta0 + (size_t)ms * L.meta_stride;
        const unsigned short* s_nst = reinterpret_cast<const unsigned short*>(M + L.nst_off);
        const unsigned short* s_lc = reinterpret_cast<const unsigned short*>(M + L.lc_off);
        const unsigned short* s_inc = reinterpret_cast<const unsigned short*>(M + L.inc_off);
        const double* Xs = reinterpret_cast<const double*>(data0 + (size_t)ds * L.data_stride);
        const double* xs = Xs + 3 * mp;
        const i64 e0 = (i64)t0 * TE;
        const int nel = (int)min((i64)TE, ne - e0);
        const int nref = nel * NPC;
        const int m = m0, off = off0;
        if (tid < nel) {
            int l[NPC];
#pragma unroll
            for (int a = 0; a < NPC; a++) l[a] = s_lc[NPC * tid + a];
            if (NPC == 4) {
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
            for (int p = lo; p < hi; p++) {
                int ref = s_inc[p];
                s0 += scratch[3 * ref]; s1 += scratch[3 * ref + 1]; s2 += scratch[3 * ref + 2];
            }
            double* o = stage + 3 * (size_t)(off + j);
            o[0] = s0; o[1] = s1; o[2] = s2;
        }
        // rotate the pipeline registers
        t0 = t1; off0 = off1; m0 = m1;
        t1 = t2; off1 = off2; m1 = m2;
        t2 = t3; off2 = off3; m2 = m3;
        lam = lam_n; mu = mu_n;
    }
    cp_async_wait_all();
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
    const bool use_pipe = getenv("TOE_EBE_PIPE") != nullptr;                   // pipelined persistent form: measured, no gain (0.3864 vs 0.3865 ms at 10M tets — the tile kernel is bound by shared memory and the FP64 pipe, not by global latency)
    bool piped = false;
    if (use_pipe && !mask) {
        const int mp = (ctx->tile_max_nodes + 7) & ~7;
        const PipeLayout L(mp);
        if (L.total <= 200 * 1024) {
            size_t* pipe_attr = ctx->pipe_attr_smem;
            if (L.total > 48 * 1024 && L.total > pipe_attr[npc == 8]) {
                CU(cudaFuncSetAttribute(k_ebe_pipe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
                CU(cudaFuncSetAttribute(k_ebe_pipe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
                pipe_attr[npc == 8] = L.total;
            }
            int per_sm = (int)((220 * 1024) / (L.total + 1024));
            const int cap = npc == 4 ? 4 : 8;                                  // 1024 threads per SM at most
            per_sm = per_sm < 1 ? 1 : (per_sm > cap ? cap : per_sm);
            unsigned grid = min_u((unsigned)ctx->ntiles, (unsigned)(N_SM * per_sm));
            if (const char* eg = getenv("TOE_EBE_PIPE_GRID")) { int v = atoi(eg); if (v >= 1 && (unsigned)v < grid) grid = (unsigned)v; }   // tests: several tiles per CTA on small meshes
#define PIPE_ARGS (const int*)ctx->tile_off.p, (const int*)ctx->tile_m.p, (const int*)ctx->tile_nodes.p, (const unsigned short*)ctx->tile_lconn.p, \
            (const unsigned short*)ctx->tile_inc.p, (const unsigned short*)ctx->tile_nstart.p, (const double*)ctx->xq.p, ctx->mat, x, ctx->tile_stage.p, ctx->ne, \
            ctx->ntiles, mp, done_flag
            if (npc == 4) LAUNCH(ctx, k_ebe_pipe<4>, grid, 256, L.total, PIPE_ARGS);
            else          LAUNCH(ctx, k_ebe_pipe<8>, grid, 128, L.total, PIPE_ARGS);
#undef PIPE_ARGS
            piped = true;
        }
    }
    size_t smem = (3 * (size_t)TILE_REFS + 6 * (size_t)ctx->tile_max_nodes) * sizeof(double) + (TILE_REFS + (size_t)ctx->tile_max_nodes + 8) * sizeof(unsigned short);
    size_t* attr_smem = ctx->ebe_attr_smem;
    if (smem > 48 * 1024 && smem > attr_smem[npc == 8]) {
        CU(cudaFuncSetAttribute(k_ebe_tile<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_ebe_tile<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_ebe_tile<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_ebe_tile<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem[npc == 8] = smem;
    }
#define TILE_ARGS (const int*)ctx->tile_off.p, (const int*)ctx->tile_m.p, (const int*)ctx->tile_nodes.p, (const unsigned short*)ctx->tile_lconn.p, (const unsigned short*)ctx->tile_inc.p, \
        (const unsigned short*)ctx->tile_nstart.p, (const double*)ctx->xq.p, ctx->mat, (const unsigned char*)ctx->dflag.p, x, ctx->tile_stage.p, ctx->ne, done_flag
    if (piped) {}
    else if (npc == 4) { if (mask) LAUNCH(ctx, (k_ebe_tile<4, true>), ctx->ntiles, 256, smem, TILE_ARGS); else LAUNCH(ctx, (k_ebe_tile<4, false>), ctx->ntiles, 256, smem, TILE_ARGS); }
    else               { if (mask) LAUNCH(ctx, (k_ebe_tile<8, true>), ctx->ntiles, 128, smem, TILE_ARGS); else LAUNCH(ctx, (k_ebe_tile<8, false>), ctx->ntiles, 128, smem, TILE_ARGS); }
#undef TILE_ARGS
    unsigned grid = div_up(ctx->nq, 256);
#define NODE_ARGS (const int*)ctx->nst_ptr.p, (const int*)ctx->nst.p, (const double*)ctx->tile_stage.p, (const unsigned char*)ctx->dflag.p, (const double*)ctx->dval.p, \
        (int)ctx->any_dirichlet, ctx->owned, x, y, ctx->nq, done_flag, cg, ctx->partials.p, ctx->counters.p + 1, dot_out
    if (cg) LAUNCH(ctx, k_ebe_nodes<true>, grid, 256, 0, NODE_ARGS);
    else    LAUNCH(ctx, k_ebe_nodes<false>, grid, 256, 0, NODE_ARGS);
#undef NODE_ARGS
    return TOE_OK;
}
