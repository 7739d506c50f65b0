// assemble.cu — kernels (1)+(2) of the north star: batched element stiffness (Tet4 / Hex8, SIMP modulus),
// assembly into the precomputed block pattern, loads and Dirichlet handling.
//
// Reference loops replaced: FiniteElementAnalysis.jl:204-250 / :654-707 (assembly), :392-418 (nodal force),
// VolumeForce.jl:26-94 / :176-243 (body loads), Ferrite apply!(K,f,ch) called at :540-542 / :841-843.
//
// Layout of K in HBM: the pattern is stored once per 3x3 node block (blk_ptr/blk_col, sorted); the values are
// 9 planes of nnzb doubles: val[(3c+d)*nnzb + s] = K[3q+c, 3q'+d] for block slot s = (row node q, col node q').
// Consecutive threads own consecutive slots, so every plane access of the assembly and of the SpMV is a
// fully coalesced 8-byte stream.
#include "element.cuh"
#include <climits>
#include <cstdlib>

// ---------------------------------------------------------------------------------------------------------
// vectors / scratch
// ---------------------------------------------------------------------------------------------------------
static const int PARTIALS_CAP = 1 << 16;

int dist_node_dofs(toe_ctx* ctx, const int** node_q_g);
int dist_localize_cells(toe_ctx* ctx, const double* global_host, double* local_dev);

int ensure_vectors(toe_ctx* ctx) {
    size_t n = 3 * (size_t)ctx->nq;
    bool fresh = ctx->f.n < n || !ctx->f.p;
    CU(ctx->f.alloc(n)); CU(ctx->u.alloc(n)); CU(ctx->r.alloc(n)); CU(ctx->p.alloc(n)); CU(ctx->Ap.alloc(n));
    CU(ctx->Minv.alloc(n)); CU(ctx->diag.alloc(n)); CU(ctx->tmp.alloc(n));
    CU(ctx->dflag.alloc(n)); CU(ctx->dval.alloc(n));
    CU(ctx->partials.alloc((size_t)PARTIALS_CAP + (size_t)ctx->nq / 32));
    CU(ctx->counters.alloc(16));
    CU(ctx->cgs.alloc(1));
    CU(ctx->errflag.alloc(4));
    if (fresh) {
        CU(cudaMemsetAsync(ctx->f.p, 0, n * sizeof(double), ctx->stream));
        CU(cudaMemsetAsync(ctx->u.p, 0, n * sizeof(double), ctx->stream));
        CU(cudaMemsetAsync(ctx->dflag.p, 0, n, ctx->stream));
        CU(cudaMemsetAsync(ctx->dval.p, 0, n * sizeof(double), ctx->stream));
        CU(cudaMemsetAsync(ctx->counters.p, 0, 16 * sizeof(unsigned int), ctx->stream));
    }
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Ke batch (parity hook): one thread per (cell, a, b) node pair writes its 3x3 block, column-major Ke
// ---------------------------------------------------------------------------------------------------------
template <int NPC>
__device__ __forceinline__ double pair_block(const int* __restrict__ cq, const double* __restrict__ xq, const Material& mat,
                                             int e, int a, int b, double B[9]) {
    double lam, mu; material_at(mat, e, lam, mu);
    int q[NPC]; double X[NPC][3];
    if (NPC == 4) {
        tet_load(cq, xq, e, q, (double(*)[3])X);
        double g[4][3];
        double det = tet_grads((const double(*)[3])X, g);
        block_ab(g[a], g[b], lam, mu, det * (1.0 / 6.0), B, false);
        return det;
    } else {
        hex_load(cq, xq, e, q, (double(*)[3])X);
        double mindet = 1e300;
#pragma unroll
        for (int k = 0; k < 9; k++) B[k] = 0.0;
        for (int gp = 0; gp < 8; gp++) {
            double g[8][3], N[8];
            double det = hex_grads_at((const double(*)[3])X, gp, g, N);
            mindet = fmin(mindet, det);
            double ga[3] = {g[0][0], g[0][1], g[0][2]}, gb[3] = {g[0][0], g[0][1], g[0][2]};
#pragma unroll
            for (int k = 1; k < 8; k++) {     // select without dynamic register indexing
                if (k == a) { ga[0] = g[k][0]; ga[1] = g[k][1]; ga[2] = g[k][2]; }
                if (k == b) { gb[0] = g[k][0]; gb[1] = g[k][1]; gb[2] = g[k][2]; }
            }
            block_ab(ga, gb, lam, mu, det, B, true);
        }
        return mindet;
    }
}

// tet specialisation with register-resident gradient selection
__device__ __forceinline__ void sel4(const double g[4][3], int a, double o[3]) {
    o[0] = g[0][0]; o[1] = g[0][1]; o[2] = g[0][2];
#pragma unroll
    for (int k = 1; k < 4; k++) if (k == a) { o[0] = g[k][0]; o[1] = g[k][1]; o[2] = g[k][2]; }
}

template <int NPC>
__global__ void k_ke_batch(const int* __restrict__ cq, const double* __restrict__ xq, Material mat, i64 first, i64 count,
                           double* __restrict__ out, int* err) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * NPC * NPC) return;
    int b = (int)(t % NPC); int a = (int)((t / NPC) % NPC); i64 k = t / (NPC * NPC);
    int e = (int)(first + k);
    double B[9];
    double det;
    if (NPC == 4) {
        double lam, mu; material_at(mat, e, lam, mu);
        int q[4]; double X[4][3], g[4][3], ga[3], gb[3];
        tet_load(cq, xq, e, q, X);
        det = tet_grads(X, g);
        sel4(g, a, ga); sel4(g, b, gb);
        block_ab(ga, gb, lam, mu, det * (1.0 / 6.0), B, false);
    } else {
        det = pair_block<NPC>(cq, xq, mat, e, a, b, B);
    }
    if (!(det > 0.0)) atomicMin(err + 1, e);
    const int nb = 3 * NPC;
    double* ke = out + (size_t)k * nb * nb;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int d = 0; d < 3; d++) ke[(3 * a + c) + (size_t)nb * (3 * b + d)] = B[3 * c + d];
}

// ---------------------------------------------------------------------------------------------------------
// assembly, variant ATOMIC: one thread per cell, red.global.add.f64 into the planes (baseline / cross-check)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_slot(const int* __restrict__ blk_ptr, const int* __restrict__ blk_col, int row, int col) {
    int lo = __ldg(&blk_ptr[row]), hi = __ldg(&blk_ptr[row + 1]) - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(&blk_col[mid]) < col) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_asm_atomic_tet(const int* __restrict__ cq, const double* __restrict__ xq, Material mat,
                                 const int* __restrict__ blk_ptr, const int* __restrict__ blk_col,
                                 double* __restrict__ val, i64 nnzb, int ne, int* err) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    double lam, mu; material_at(mat, e, lam, mu);
    int q[4]; double X[4][3], g[4][3];
    tet_load(cq, xq, e, q, X);
    double det = tet_grads(X, g);
    if (!(det > 0.0)) { atomicMin(err + 1, e); return; }
    double w = det * (1.0 / 6.0);
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            double B[9];
            block_ab(g[a], g[b], lam, mu, w, B, false);
            int s = find_slot(blk_ptr, blk_col, q[a], q[b]);
#pragma unroll
            for (int k = 0; k < 9; k++) atomicAdd(&val[(size_t)k * nnzb + s], B[k]);
        }
}

// hex: one thread per (cell, a): block row a of Ke, accumulated over the 8 Gauss points in registers
__global__ void k_asm_atomic_hex(const int* __restrict__ cq, const double* __restrict__ xq, Material mat,
                                 const int* __restrict__ blk_ptr, const int* __restrict__ blk_col,
                                 double* __restrict__ val, i64 nnzb, int ne, int* err) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)ne * 8) return;
    int e = (int)(t >> 3), a = (int)(t & 7);
    double lam, mu; material_at(mat, e, lam, mu);
    int q[8]; double X[8][3];
    hex_load(cq, xq, e, q, X);
    double M[8][9];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int k = 0; k < 9; k++) M[b][k] = 0.0;
    bool bad = false;
    for (int gp = 0; gp < 8; gp++) {
        double g[8][3], N[8];
        double det = hex_grads_at(X, gp, g, N);
        bad |= !(det > 0.0);
        double ga[3] = {g[0][0], g[0][1], g[0][2]};
#pragma unroll
        for (int k = 1; k < 8; k++) if (k == a) { ga[0] = g[k][0]; ga[1] = g[k][1]; ga[2] = g[k][2]; }
#pragma unroll
        for (int b = 0; b < 8; b++) block_ab(ga, g[b], lam, mu, det, M[b], true);
    }
    if (bad) { atomicMin(err + 1, e); return; }
    int qa = q[0];
#pragma unroll
    for (int k = 1; k < 8; k++) if (k == a) qa = q[k];
#pragma unroll
    for (int b = 0; b < 8; b++) {
        int s = find_slot(blk_ptr, blk_col, qa, q[b]);
#pragma unroll
        for (int k = 0; k < 9; k++) atomicAdd(&val[(size_t)k * nnzb + s], M[b][k]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// assembly, variant GATHER: one thread per block slot sums its contributions in ascending cell order.
// Off-diagonal blocks walk the precomputed (e,a,b) list; diagonal blocks walk the node's incidence list.
// ---------------------------------------------------------------------------------------------------------
template <int NPC>
__global__ void __launch_bounds__(128) k_asm_offdiag(const int* __restrict__ ctr_ptr, const int* __restrict__ ctr,
                                                     const int* __restrict__ cq, const double* __restrict__ xq, Material mat,
                                                     double* __restrict__ val, i64 nnzb, i64 ldv) {
    i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nnzb) return;
    int lo = __ldg(&ctr_ptr[s]), hi = __ldg(&ctr_ptr[s + 1]);
    if (lo == hi) return;                       // diagonal block: written by k_asm_diag
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) acc[k] = 0.0;
    for (int i = lo; i < hi; i++) {
        int e, a, b; ctr_unpack<NPC>(__ldg(&ctr[i]), e, a, b);
        if (NPC == 4) {
            double lam, mu; material_at(mat, e, lam, mu);
            int q[4]; double X[4][3], g[4][3], ga[3], gb[3];
            tet_load(cq, xq, e, q, X);
            double det = tet_grads(X, g);
            sel4(g, a, ga); sel4(g, b, gb);
            block_ab(ga, gb, lam, mu, det * (1.0 / 6.0), acc, true);
        } else {
            double B[9];
            pair_block<NPC>(cq, xq, mat, e, a, b, B);
#pragma unroll
            for (int k = 0; k < 9; k++) acc[k] += B[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) val[(size_t)k * ldv + s] = acc[k];
}

template <int NPC>
__global__ void __launch_bounds__(128) k_asm_diag(const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                  const int* __restrict__ diag_slot,
                                                  const int* __restrict__ cq, const double* __restrict__ xq, Material mat,
                                                  double* __restrict__ val, i64 nnzb, int nq, int* err) {
    int qn = blockIdx.x * blockDim.x + threadIdx.x;
    if (qn >= nq) return;
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) acc[k] = 0.0;
    int lo = __ldg(&inc_ptr[qn]), hi = __ldg(&inc_ptr[qn + 1]);
    for (int i = lo; i < hi; i++) {
        int ea = __ldg(&inc[i]);
        int e = ea / NPC, a = ea - e * NPC;
        if (NPC == 4) {
            double lam, mu; material_at(mat, e, lam, mu);
            int q[4]; double X[4][3], g[4][3], ga[3];
            tet_load(cq, xq, e, q, X);
            double det = tet_grads(X, g);
            if (!(det > 0.0)) atomicMin(err + 1, e);
            sel4(g, a, ga);
            block_ab(ga, ga, lam, mu, det * (1.0 / 6.0), acc, true);
        } else {
            double B[9];
            double det = pair_block<NPC>(cq, xq, mat, e, a, a, B);
            if (!(det > 0.0)) atomicMin(err + 1, e);
#pragma unroll
            for (int k = 0; k < 9; k++) acc[k] += B[k];
        }
    }
    int s = __ldg(&diag_slot[qn]);
#pragma unroll
    for (int k = 0; k < 9; k++) val[(size_t)k * nnzb + s] = acc[k];
}

// ---------------------------------------------------------------------------------------------------------
// assembly, variant ROWS (Tet4): a group of G threads owns one node row of K.
//   pass 1  the group evaluates the geometry of the row's cells ONCE (gradients, w·λ, w·μ; ASM_CH cells per pass) into shared
//           memory — GATHER re-derives it for every block, 4x per (row, cell).  Layout: 14 planes of ASM_CH doubles per row, so
//           that the threads of a group, which read different cells of the row, hit different banks (cell index = bank);
//   pass 2  thread `lane` owns block slot blk_ptr[row]+lane and walks its contribution list in ascending cell order.  The list is
//           the row-relative form built at set-up (mesh.cu: position of the cell in the row's incidence list, a, b in 16 bits),
//           so a contribution is 2 bytes of index and 8 shared-memory doubles:
//              B += wλ g_a⊗g_b + wμ g_b⊗g_a + wμ (g_a·g_b) I        (9 DMUL + 21 DFMA + 2 DADD per contribution)
//   the diagonal block (one contribution per cell of the row, 4x an off-diagonal block) is spread over the group in pass 1.
// Products g_a[c]·g_b[d] are shared by block (a,b) and its transpose and the accumulation order is the cell order, so K
// stays bitwise symmetric and bit-reproducible.  Groups of ≤ 32 threads live inside one warp: no CTA-wide barrier.
// ---------------------------------------------------------------------------------------------------------
static const int ASM_ROWS_THREADS = 128;
static const int ASM_CH = 32;                 // cells of a row staged per pass
static const int ASM_GEO = 14;                // planes per row: g[4][3], w·λ, w·μ

__device__ __forceinline__ void block_acc(const double ga[3], const double gb[3], double wl, double wm, double acc[9]) {
    double p[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int d = 0; d < 3; d++) p[c][d] = ga[c] * gb[d];
    const double dot = (p[0][0] + p[1][1]) + p[2][2];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double v = fma(wl, p[c][d], acc[3 * c + d]);
            acc[3 * c + d] = fma(wm, p[d][c], v);
        }
#pragma unroll
    for (int c = 0; c < 3; c++) acc[4 * c] = fma(wm, dot, acc[4 * c]);
}

template <int G>
__global__ void __launch_bounds__(ASM_ROWS_THREADS, G <= 32 ? 5 : 4) k_asm_rows_tet(const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                                   const int* __restrict__ blk_ptr, const int* __restrict__ blk_col,
                                                                   const int* __restrict__ ctr_ptr, const unsigned short* __restrict__ rctr,
                                                                   const int* __restrict__ cq, const double* __restrict__ xq, Material mat,
                                                                   double* __restrict__ val, i64 ldv, int nq, int* err) {
    constexpr int ROWS = ASM_ROWS_THREADS / G;
    constexpr bool IN_WARP = (G <= 32);            // the group is (part of) one warp: warp-level synchronisation suffices
    constexpr int CPL = ASM_CH / G > 0 ? ASM_CH / G : 1;      // cells per lane and pass (2 for G = 16, 1 otherwise)
    __shared__ double geo[ROWS][ASM_GEO][ASM_CH];
    __shared__ double dpart[IN_WARP ? 1 : ASM_ROWS_THREADS][6];      // G = 64 only: per-thread partials of the row's diagonal block
    __shared__ int s_maxcells;
    const int rl = threadIdx.x / G, lane = threadIdx.x - rl * G;
    const int row = blockIdx.x * ROWS + rl;
    const bool live = row < nq;
    int i_lo = 0, i_hi = 0, s = -1, col = -1, ci = 0, chi = 0;
    if (live) {
        i_lo = __ldg(&inc_ptr[row]); i_hi = __ldg(&inc_ptr[row + 1]);
        const int b0 = __ldg(&blk_ptr[row]), b1 = __ldg(&blk_ptr[row + 1]);
        if (lane < b1 - b0) { s = b0 + lane; col = __ldg(&blk_col[s]); ci = __ldg(&ctr_ptr[s]); chi = __ldg(&ctr_ptr[s + 1]); }
    }
    const int ncell = i_hi - i_lo;
    int npass;
    if (IN_WARP) {                                  // uniform over the warp (it may hold two or more rows)
        int m = ncell;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        npass = (m + ASM_CH - 1) / ASM_CH;
    } else {
        if (threadIdx.x == 0) s_maxcells = 0;
        __syncthreads();
        if (live && lane == 0) atomicMax(&s_maxcells, ncell);
        __syncthreads();
        npass = (s_maxcells + ASM_CH - 1) / ASM_CH;
    }
    const bool is_diag = (s >= 0 && col == row);
    double acc[9], dacc[6];                         // dacc: upper triangle of the diagonal block (00 01 02 11 12 22)
#pragma unroll
    for (int k = 0; k < 9; k++) acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) dacc[k] = 0.0;
    unsigned ent = (ci < chi) ? (unsigned)__ldg(&rctr[ci]) : 0u;          // next contribution of this thread's block
    for (int pass = 0; pass < npass; pass++) {
        const int k0 = pass * ASM_CH;
        int cnt = ncell - k0;
        cnt = cnt < 0 ? 0 : (cnt > ASM_CH ? ASM_CH : cnt);
        // the incidence entries and connectivities of all this lane's cells are requested before any geometry is evaluated: the
        // chain incidence → connectivity → coordinates is paid once per pass, not once per cell
        int ea[CPL]; int4 cn[CPL];
#pragma unroll
        for (int j = 0; j < CPL; j++) { const int k = lane + j * G; ea[j] = k < cnt ? __ldg(&inc[i_lo + k0 + k]) : -1; }
#pragma unroll
        for (int j = 0; j < CPL; j++) cn[j] = ea[j] >= 0 ? __ldg(reinterpret_cast<const int4*>(cq) + (ea[j] >> 2)) : make_int4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < CPL; j++) {
            if (ea[j] < 0) continue;
            const int k = lane + j * G, e = ea[j] >> 2;
            double lam, mu; material_at(mat, e, lam, mu);
            double X[4][3], g[4][3];
            load3(xq, cn[j].x, X[0]); load3(xq, cn[j].y, X[1]); load3(xq, cn[j].z, X[2]); load3(xq, cn[j].w, X[3]);
            const double det = tet_grads(X, g);
            if (!(det > 0.0)) atomicMin(err + 1, e);
            const double w = det * (1.0 / 6.0);
            const double wl = w * lam, wm = w * mu;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int i = 0; i < 3; i++) geo[rl][3 * a + i][k] = g[a][i];
            geo[rl][12][k] = wl; geo[rl][13][k] = wm;
            // the diagonal block has a contribution from every cell of the row (4x the work of an off-diagonal block): it is
            // spread over the group here — each thread adds its cells (ascending), the partials are summed in a fixed order below
            double ga[3];
            sel4(g, ea[j] & 3, ga);
            const double dot = (ga[0] * ga[0] + ga[1] * ga[1]) + ga[2] * ga[2];
            const double wlm = wl + wm;
            dacc[0] = fma(wlm, ga[0] * ga[0], fma(wm, dot, dacc[0])); dacc[1] = fma(wlm, ga[0] * ga[1], dacc[1]); dacc[2] = fma(wlm, ga[0] * ga[2], dacc[2]);
            dacc[3] = fma(wlm, ga[1] * ga[1], fma(wm, dot, dacc[3])); dacc[4] = fma(wlm, ga[1] * ga[2], dacc[4]);
            dacc[5] = fma(wlm, ga[2] * ga[2], fma(wm, dot, dacc[5]));
        }
        if (IN_WARP) __syncwarp(); else __syncthreads();
        if (!is_diag) {
            while (ci < chi) {
                const int k = (int)(ent >> 4) - k0;
                if (k >= cnt) break;                                     // belongs to a later pass
                const int a = (ent >> 2) & 3, b = ent & 3;
                ci++;
                const unsigned nxt = (ci < chi) ? (unsigned)__ldg(&rctr[ci]) : 0u;
                const double (*gr)[ASM_CH] = geo[rl];
                const double ga[3] = {gr[3 * a][k], gr[3 * a + 1][k], gr[3 * a + 2][k]};
                const double gb[3] = {gr[3 * b][k], gr[3 * b + 1][k], gr[3 * b + 2][k]};
                block_acc(ga, gb, gr[12][k], gr[13][k], acc);
                ent = nxt;
            }
        }
        if (IN_WARP) __syncwarp(); else __syncthreads();
    }
    // diagonal block = sum of the group's partials, in a fixed order: a butterfly over the lanes of the group (all lanes end up with
    // the same bits), or — groups of two warps — through shared memory in lane order
    if (IN_WARP) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 6; k++) dacc[k] += __shfl_xor_sync(0xffffffffu, dacc[k], o);
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) dpart[threadIdx.x][k] = dacc[k];
        __syncthreads();
        if (is_diag) {
#pragma unroll
            for (int k = 0; k < 6; k++) dacc[k] = 0.0;
            for (int l = 0; l < G; l++)
#pragma unroll
                for (int k = 0; k < 6; k++) dacc[k] += dpart[rl * G + l][k];
        }
    }
    if (is_diag) { acc[0] = dacc[0]; acc[1] = dacc[1]; acc[2] = dacc[2]; acc[3] = dacc[1]; acc[4] = dacc[3]; acc[5] = dacc[4]; acc[6] = dacc[2]; acc[7] = dacc[4]; acc[8] = dacc[5]; }
    if (s >= 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) val[(size_t)k * ldv + s] = acc[k];
    }
}

__global__ void k_simp_lame(Material mat, double* __restrict__ lam_e, double* __restrict__ mu_e, int ne) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    double lam, mu; material_at(mat, e, lam, mu);
    lam_e[e] = lam; mu_e[e] = mu;
}

static int check_detj(toe_ctx* ctx, const char* who, int* aux_flag = nullptr) {
    int h[4];
    CU(cudaMemcpyAsync(h, ctx->errflag.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (aux_flag) *aux_flag = h[2];
    if (h[1] != INT_MAX) {
        int bad = h[1];
        int reset = INT_MAX;
        cudaMemcpyAsync(ctx->errflag.p + 1, &reset, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        return toe_fail(ctx, TOE_ERR_MESH, "%s: det(J) is not positive in cell %d (Ferrite reinit! throws here too)", who, bad + 1);
    }
    return TOE_OK;
}

static int reset_detj_flag(toe_ctx* ctx) {
    int reset = INT_MAX;
    CU(cudaMemcpyAsync(ctx->errflag.p + 1, &reset, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    return TOE_OK;
}

int assemble_current_material(toe_ctx* ctx, int variant) {
    if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "assemble: call toe_build_pattern first (setup_problem)");
    if (ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "assemble: no material set");
    TRY(ensure_vectors(ctx));
    if (variant == TOE_ASM_AUTO) variant = TOE_ASM_GATHER;
    if (variant != TOE_ASM_ATOMIC && variant != TOE_ASM_GATHER && variant != TOE_ASM_ROWS) return toe_fail(ctx, TOE_ERR_ARG, "unknown assembly variant %d", variant);
    if (variant != TOE_ASM_ATOMIC) TRY(mesh_build_contrib(ctx));      // one-off per mesh, outside the timed stage
    // ROWS is written for Tet4, rows of at most 64 blocks and < 4096 cells around a node; everything else takes the GATHER kernels
    if (variant == TOE_ASM_ROWS && (ctx->npc != 4 || ctx->max_deg > 64 || !ctx->have_rctr)) variant = TOE_ASM_GATHER;
    size_t n = 3 * (size_t)ctx->nq;
    i64 nnzb = ctx->nnzb, ldv = ctx->ldv;
    if (ctx->val.n < 9 * (size_t)ldv + 16) {
        CU(ctx->val.alloc(9 * (size_t)ldv + 16));
        CU(cudaMemsetAsync(ctx->val.p, 0, ctx->val.bytes(), ctx->stream));   // plane padding is read (never used) by the SpMV's bulk copies
    }
    TRY(reset_detj_flag(ctx));
    Material amat = ctx->mat;
    if (variant != TOE_ASM_ATOMIC && ctx->mat.mode == MAT_SIMP) { CU(ctx->lamw.alloc(ctx->ne)); CU(ctx->muw.alloc(ctx->ne)); }
    StageTimer T(ctx, &ctx->tm.assemble);
    if (variant != TOE_ASM_ATOMIC && ctx->mat.mode == MAT_SIMP) {
        // the gather kernels visit a cell once per block it contributes to: evaluate E(ρ)=Emin+(E0-Emin)ρ^p once per cell instead
        LAUNCH(ctx, k_simp_lame, div_up(ctx->ne, 256), 256, 0, ctx->mat, ctx->lamw.p, ctx->muw.p, (int)ctx->ne);
        amat.mode = MAT_PERCELL; amat.lam_e = ctx->lamw.p; amat.mu_e = ctx->muw.p;
    }
    // start_assemble(K, f): zero K and f (FiniteElementAnalysis.jl:211 / :661)
    CU(cudaMemsetAsync(ctx->f.p, 0, n * sizeof(double), ctx->stream));
    // assembling a fresh K discards constraints applied to the previous one
    CU(cudaMemsetAsync(ctx->dflag.p, 0, n, ctx->stream));
    CU(cudaMemsetAsync(ctx->dval.p, 0, n * sizeof(double), ctx->stream));
    ctx->any_dirichlet = false;
    int ne = (int)ctx->ne;
    if (variant == TOE_ASM_ATOMIC) {
        CU(cudaMemsetAsync(ctx->val.p, 0, (9 * (size_t)ldv + 16) * sizeof(double), ctx->stream));
        if (ctx->npc == 4)
            LAUNCH(ctx, k_asm_atomic_tet, div_up(ne, 128), 128, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, ctx->mat,
                   (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, ctx->val.p, ldv, ne, ctx->errflag.p);
        else
            LAUNCH(ctx, k_asm_atomic_hex, div_up((i64)ne * 8, 64), 64, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, ctx->mat,
                   (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, ctx->val.p, ldv, ne, ctx->errflag.p);
    } else if (variant == TOE_ASM_ROWS) {
        CU(cudaMemsetAsync(ctx->errflag.p + 2, 0, sizeof(int), ctx->stream));
#define ROWS_ARGS (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, (const int*)ctx->ctr_ptr.p, \
        (const unsigned short*)ctx->rctr.p, (const int*)ctx->cq.p, (const double*)ctx->xq.p, amat, ctx->val.p, ldv, ctx->nq, ctx->errflag.p
        if (!ctx->rows_attr_set) {                            // 5 CTAs x 29 KB per SM: ask for the large shared-memory carve-out
            CU(cudaFuncSetAttribute(k_asm_rows_tet<16>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            CU(cudaFuncSetAttribute(k_asm_rows_tet<32>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            ctx->rows_attr_set = true;
        }
        int gmin = 0;                                         // TOE_ASM_ROWS_G=32|64 forces a wider group (tests)
        if (const char* eg = getenv("TOE_ASM_ROWS_G")) gmin = atoi(eg);
        if (ctx->max_deg <= 16 && gmin <= 16)      LAUNCH(ctx, k_asm_rows_tet<16>, div_up(ctx->nq, ASM_ROWS_THREADS / 16), ASM_ROWS_THREADS, 0, ROWS_ARGS);
        else if (ctx->max_deg <= 32 && gmin <= 32) LAUNCH(ctx, k_asm_rows_tet<32>, div_up(ctx->nq, ASM_ROWS_THREADS / 32), ASM_ROWS_THREADS, 0, ROWS_ARGS);
        else                         LAUNCH(ctx, k_asm_rows_tet<64>, div_up(ctx->nq, ASM_ROWS_THREADS / 64), ASM_ROWS_THREADS, 0, ROWS_ARGS);
#undef ROWS_ARGS
    } else {
        if (ctx->npc == 4) {
            LAUNCH(ctx, k_asm_offdiag<4>, div_up(nnzb, 128), 128, 0, (const int*)ctx->ctr_ptr.p, (const int*)ctx->ctr.p,
                   (const int*)ctx->cq.p, (const double*)ctx->xq.p, amat, ctx->val.p, nnzb, ldv);
            LAUNCH(ctx, k_asm_diag<4>, div_up(ctx->nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->diag_slot.p,
                   (const int*)ctx->cq.p, (const double*)ctx->xq.p, amat, ctx->val.p, ldv, ctx->nq, ctx->errflag.p);
        } else {
            LAUNCH(ctx, k_asm_offdiag<8>, div_up(nnzb, 128), 128, 0, (const int*)ctx->ctr_ptr.p, (const int*)ctx->ctr.p,
                   (const int*)ctx->cq.p, (const double*)ctx->xq.p, amat, ctx->val.p, nnzb, ldv);
            LAUNCH(ctx, k_asm_diag<8>, div_up(ctx->nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->diag_slot.p,
                   (const int*)ctx->cq.p, (const double*)ctx->xq.p, amat, ctx->val.p, ldv, ctx->nq, ctx->errflag.p);
        }
    }
    TRY(T.finish());
    int lists_bad = 0;
    TRY(check_detj(ctx, "assemble_stiffness_matrix", variant == TOE_ASM_ROWS ? &lists_bad : nullptr));
    if (lists_bad) return toe_fail(ctx, TOE_ERR_STATE, "assemble (ROWS): incidence and contribution lists are out of step (internal error)");
    ctx->have_K = true; ctx->have_diag = false; ctx->have_solution = false;
    ctx->op_generation++;
    return TOE_OK;
}

int ke_batch(toe_ctx* ctx, i64 first, i64 count, double* out_host) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "toe_ke_batch: DOFs not built");
    if (ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "toe_ke_batch: no material set");
    if (first < 1 || count < 0 || first - 1 + count > ctx->ne) return toe_fail(ctx, TOE_ERR_ARG, "toe_ke_batch: cell range out of bounds");
    if (count == 0) return TOE_OK;
    TRY(ensure_vectors(ctx));
    TRY(reset_detj_flag(ctx));
    int nb = 3 * ctx->npc;
    DevBuf<double> out; CU(out.alloc((size_t)count * nb * nb));
    i64 threads = count * ctx->npc * ctx->npc;
    if (ctx->npc == 4) LAUNCH(ctx, k_ke_batch<4>, div_up(threads, 128), 128, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, ctx->mat, first - 1, count, out.p, ctx->errflag.p);
    else               LAUNCH(ctx, k_ke_batch<8>, div_up(threads, 64), 64, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, ctx->mat, first - 1, count, out.p, ctx->errflag.p);
    CU(cudaMemcpyAsync(out_host, out.p, (size_t)count * nb * nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(check_detj(ctx, "toe_ke_batch"));
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// loads
// ---------------------------------------------------------------------------------------------------------
__global__ void k_nodal_force(const int64_t* __restrict__ nodes, i64 nnodes, const int* __restrict__ node_q, i64 nn,
                              double fx, double fy, double fz, double* f, int* err) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    int64_t g = nodes[i];
    if (g < 1 || g > nn) { atomicExch(err + 2, 1); return; }
    int q = node_q[g - 1];
    if (q < 0) return;                                  // node in no cell: skipped like `haskey` (:402)
    atomicAdd(&f[3 * (size_t)q], fx); atomicAdd(&f[3 * (size_t)q + 1], fy); atomicAdd(&f[3 * (size_t)q + 2], fz);
}

int add_nodal_force(toe_ctx* ctx, const int64_t* nodes, i64 nnodes, const double F[3]) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "apply_force!: DOFs not built");
    if (nnodes <= 0 || !nodes) return toe_fail(ctx, TOE_ERR_ARG, "No nodes provided for force application.");   // :393-395
    TRY(ensure_vectors(ctx));
    StageTimer T(ctx, &ctx->tm.loads);
    TmpBuf<int64_t> d(ctx->stream); CU(d.alloc(nnodes));
    CU(cudaMemcpyAsync(d.p, nodes, nnodes * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->errflag.p + 2, 0, sizeof(int), ctx->stream));
    double inv = 1.0 / (double)nnodes;    // force_vector ./ length(nodes) (:401)
    LAUNCH(ctx, k_nodal_force, div_up(nnodes, 128), 128, 0, (const int64_t*)d.p, nnodes, (const int*)ctx->node_q.p, ctx->nn,
           F[0] / (double)nnodes, F[1] / (double)nnodes, F[2] / (double)nnodes, ctx->f.p, ctx->errflag.p);
    (void)inv;
    int e = 0;
    CU(cudaMemcpyAsync(&e, ctx->errflag.p + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(T.finish());
    if (e) return toe_fail(ctx, TOE_ERR_ARG, "apply_force!: node id outside 1..%lld", (long long)ctx->nn);
    ctx->have_solution = false;
    return TOE_OK;
}

// one thread per dof-node, gathers ρ b ∫N_a dΩ over its cells in ascending cell order
template <int NPC>
__global__ void k_volume_force(const int* __restrict__ inc_ptr, const int* __restrict__ inc, const int* __restrict__ cq,
                               const double* __restrict__ xq, const double* __restrict__ density, double rho_uniform, double skip_below,
                               double bx, double by, double bz, double* __restrict__ f, int nq, double* __restrict__ partials) {
    __shared__ double sh[32];
    int qn = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (qn < nq) {
        int lo = inc_ptr[qn], hi = inc_ptr[qn + 1];
        for (int i = lo; i < hi; i++) {
            int ea = inc[i];
            int e = ea / NPC, a = ea - e * NPC;
            double rho = density ? density[e] : rho_uniform;
            if (density && rho < skip_below) continue;          // VolumeForce.jl:199
            double w;                                            // ∫ N_a dΩ over the cell
            if (NPC == 4) {
                int q[4]; double X[4][3], g[4][3];
                tet_load(cq, xq, e, q, X);
                double det = tet_grads(X, g);
                w = det * (1.0 / 24.0);                         // Σ_q N_a(q) detJ/24 = detJ/24
            } else {
                int q[8]; double X[8][3];
                hex_load(cq, xq, e, q, X);
                w = 0.0;
                for (int gp = 0; gp < 8; gp++) {
                    double g[8][3], N[8];
                    double det = hex_grads_at(X, gp, g, N);
                    double Na = N[0];
#pragma unroll
                    for (int k = 1; k < 8; k++) if (k == a) Na = N[k];
                    w += Na * det;
                }
            }
            s += rho * w;
        }
        f[3 * (size_t)qn] += s * bx; f[3 * (size_t)qn + 1] += s * by; f[3 * (size_t)qn + 2] += s * bz;
    }
    double bs = block_sum(s, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
}

__global__ void k_axpy1(const double* __restrict__ x, double* __restrict__ y, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

__global__ void k_sum_partials(const double* __restrict__ partials, int n, double* out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) *out = s;
}

int add_volume_force(toe_ctx* ctx, const double b[3], double rho_uniform, const double* density_host, double skip_below, double* total_out) {
    if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "apply_volume_force!: call setup_problem first");
    TRY(ensure_vectors(ctx));
    DevBuf<double> dens;
    const double* dptr = nullptr;
    double bb[3] = {b[0], b[1], b[2]};
    if (density_host) {
        CU(dens.alloc(ctx->ne));
        if (ctx->dist) TRY(dist_localize_cells(ctx, density_host, dens.p));
        else CU(cudaMemcpyAsync(dens.p, density_host, ctx->ne * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dptr = dens.p;
    } else {
        if (rho_uniform == 0.0) return toe_fail(ctx, TOE_ERR_ARG, "apply_volume_force!: density must be non-zero");
        // body_force_per_mass = b ./ density (VolumeForce.jl:29), multiplied back by density at :76
        for (int k = 0; k < 3; k++) bb[k] = b[k] / rho_uniform;
    }
    StageTimer T(ctx, &ctx->tm.loads);
    unsigned grid = div_up(ctx->nq, 128);
    TmpBuf<double> part(ctx->stream); CU(part.alloc(grid + 1));
    // partitioned: the per-rank partial load goes to a scratch vector, is summed over the interface, then added to f
    double* target = ctx->f.p;
    size_t n = 3 * (size_t)ctx->nq;
    if (ctx->dist) { target = ctx->tmp.p; CU(cudaMemsetAsync(target, 0, n * sizeof(double), ctx->stream)); }
    if (ctx->npc == 4)
        LAUNCH(ctx, k_volume_force<4>, grid, 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p, (const double*)ctx->xq.p,
               dptr, rho_uniform, skip_below, bb[0], bb[1], bb[2], target, ctx->nq, part.p);
    else
        LAUNCH(ctx, k_volume_force<8>, grid, 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p, (const double*)ctx->xq.p,
               dptr, rho_uniform, skip_below, bb[0], bb[1], bb[2], target, ctx->nq, part.p);
    LAUNCH(ctx, k_sum_partials, 1, 256, 0, (const double*)part.p, (int)grid, part.p + grid);
    if (ctx->dist) {
        TRY(dist_post_spmv(ctx, target));
        LAUNCH(ctx, k_axpy1, div_up((i64)n, 256), 256, 0, (const double*)target, ctx->f.p, n);
        TRY(dist_allreduce(ctx, part.p + grid, 1));
    }
    double mass = 0.0;
    CU(cudaMemcpyAsync(&mass, part.p + grid, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(T.finish());
    if (total_out) for (int k = 0; k < 3; k++) total_out[k] = mass * bb[k];
    ctx->have_solution = false;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// diagonal of the current operator and Dirichlet conditions
// ---------------------------------------------------------------------------------------------------------
__global__ void k_diag_from_K(const double* __restrict__ val, i64 nnzb, const int* __restrict__ diag_slot, double* __restrict__ diag, int nq) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int s = diag_slot[q];
    diag[3 * (size_t)q] = val[s]; diag[3 * (size_t)q + 1] = val[4 * (size_t)nnzb + s]; diag[3 * (size_t)q + 2] = val[8 * (size_t)nnzb + s];
}

// matrix-free diagonal: diagonal of Σ_e Ke by node gather
template <int NPC>
__global__ void k_diag_ebe(const int* __restrict__ inc_ptr, const int* __restrict__ inc, const int* __restrict__ cq,
                           const double* __restrict__ xq, Material mat, double* __restrict__ diag, int nq) {
    int qn = blockIdx.x * blockDim.x + threadIdx.x;
    if (qn >= nq) return;
    double acc[3] = {0, 0, 0};
    int lo = inc_ptr[qn], hi = inc_ptr[qn + 1];
    for (int i = lo; i < hi; i++) {
        int ea = inc[i];
        int e = ea / NPC, a = ea - e * NPC;
        double B[9];
        if (NPC == 4) {
            double lam, mu; material_at(mat, e, lam, mu);
            int q[4]; double X[4][3], g[4][3], ga[3];
            tet_load(cq, xq, e, q, X);
            double det = tet_grads(X, g);
            sel4(g, a, ga);
            block_ab(ga, ga, lam, mu, det * (1.0 / 6.0), B, false);
        } else {
            pair_block<NPC>(cq, xq, mat, e, a, a, B);
        }
        acc[0] += B[0]; acc[1] += B[4]; acc[2] += B[8];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) diag[3 * (size_t)qn + c] = acc[c];
}

// prescribed dofs carry the m of their handler on the diagonal of the constrained operator
__global__ void k_diag_override(const unsigned char* __restrict__ dflag, const double* __restrict__ dval, double* __restrict__ diag, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && dflag[i]) diag[i] = dval[i];
}

int compute_diag(toe_ctx* ctx) {
    if (ctx->have_diag) return TOE_OK;
    TRY(ensure_vectors(ctx));
    size_t n = 3 * (size_t)ctx->nq;
    if (ctx->have_K) {
        LAUNCH(ctx, k_diag_from_K, div_up(ctx->nq, 256), 256, 0, (const double*)ctx->val.p, ctx->ldv, (const int*)ctx->diag_slot.p, ctx->diag.p, ctx->nq);
    } else {
        if (ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "no operator: assemble K or set a material first");
        if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "no incidence lists: call toe_build_pattern first");
        if (ctx->npc == 4)
            LAUNCH(ctx, k_diag_ebe<4>, div_up(ctx->nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p,
                   (const double*)ctx->xq.p, ctx->mat, ctx->diag.p, ctx->nq);
        else
            LAUNCH(ctx, k_diag_ebe<8>, div_up(ctx->nq, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p,
                   (const double*)ctx->xq.p, ctx->mat, ctx->diag.p, ctx->nq);
    }
    TRY(dist_align(ctx));                        // first exchange of a step: the ranks enter it together
    TRY(dist_post_spmv(ctx, ctx->diag.p));      // sub-assembled partitions: sum the interface contributions
    if (ctx->any_dirichlet)
        LAUNCH(ctx, k_diag_override, div_up((i64)n, 256), 256, 0, (const unsigned char*)ctx->dflag.p, (const double*)ctx->dval.p, ctx->diag.p, n);
    ctx->have_diag = true;
    return TOE_OK;
}

__global__ void k_abs_sum(const double* __restrict__ x, size_t n, const unsigned char* __restrict__ owned, double* __restrict__ partials,
                          unsigned int* counter, double* out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (!owned || owned[i / 3]) s += fabs(x[i]);
    s = block_sum(s, sh);
    double tot;
    if (grid_sum_last_block(s, partials, counter, sh, &tot)) *out = tot;
}

__global__ void k_mark_dirichlet(const int64_t* __restrict__ dofs, i64 nd, size_t n, const int* __restrict__ glob2loc,
                                 const double* __restrict__ m_dev, double inv_n,
                                 unsigned char* dflag, double* dval, double* diag, double* f, int* err) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    int64_t d = dofs[i];
    if (d < 1 || (size_t)d > n) { atomicExch(err + 2, 1); return; }        // n = global number of DOFs
    if (glob2loc) {                                                         // partitioned: keep only dofs of local nodes
        int l = glob2loc[(d - 1) / 3];
        if (l < 0) return;
        d = 3 * (int64_t)l + (d - 1) % 3 + 1;
    }
    double m = *m_dev * inv_n;
    dflag[d - 1] = 1; dval[d - 1] = m; diag[d - 1] = m; f[d - 1] = 0.0;
}

// 8 lanes per block row: zero the stored entries of prescribed rows / columns, put m on the diagonal
__global__ void k_dirichlet_K(const int* __restrict__ blk_ptr, const int* __restrict__ blk_col, const unsigned char* __restrict__ dflag,
                              const double* __restrict__ dval, const unsigned char* __restrict__ owned, double* __restrict__ val, i64 nnzb, int nq) {
    int q = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
    if (q >= nq) return;
    int sub = threadIdx.x & 7;
    unsigned char fr[3] = {dflag[3 * (size_t)q], dflag[3 * (size_t)q + 1], dflag[3 * (size_t)q + 2]};
    for (int s = blk_ptr[q] + sub; s < blk_ptr[q + 1]; s += 8) {
        int col = blk_col[s];
        unsigned char fc[3] = {dflag[3 * (size_t)col], dflag[3 * (size_t)col + 1], dflag[3 * (size_t)col + 2]};
        if (!(fr[0] | fr[1] | fr[2] | fc[0] | fc[1] | fc[2])) continue;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int d = 0; d < 3; d++)
                if (fr[c] | fc[d]) {
                    double v = 0.0;
                    if (col == q && c == d && (!owned || owned[q])) v = dval[3 * (size_t)q + c];   // sub-assembled K: m sits on the owner's copy only
                    val[(size_t)(3 * c + d) * nnzb + s] = v;
                }
    }
}

int apply_dirichlet(toe_ctx* ctx, const int64_t* dofs, i64 nd, double* mean_out) {
    if (!ctx->have_K && ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "apply!: assemble K (or set a material) first");
    TRY(ensure_vectors(ctx));
    size_t n = 3 * (size_t)ctx->nq;
    if (nd < 0 || (nd > 0 && !dofs)) return toe_fail(ctx, TOE_ERR_ARG, "apply!: bad dof list");
    TRY(compute_diag(ctx));
    StageTimer T(ctx, &ctx->tm.dirichlet);
    // m = mean(abs(diag K)) of the incoming K  (Ferrite apply!, meandiag)
    double* m_dev = &ctx->cgs.p->aux;
    LAUNCH(ctx, k_abs_sum, min_u(div_up((i64)n, 256), 1024u), 256, 0, (const double*)ctx->diag.p, n, ctx->owned, ctx->partials.p, ctx->counters.p, m_dev);
    TRY(dist_allreduce(ctx, m_dev, 1));
    double msum = 0.0;
    CU(cudaMemcpyAsync(&msum, m_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    size_t nglob = ctx->n_global ? (size_t)ctx->n_global : n;
    double inv_n = 1.0 / (double)nglob;
    if (nd > 0) {
        TmpBuf<int64_t> d(ctx->stream); CU(d.alloc(nd));
        CU(cudaMemcpyAsync(d.p, dofs, nd * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->errflag.p + 2, 0, sizeof(int), ctx->stream));
        LAUNCH(ctx, k_mark_dirichlet, div_up(nd, 128), 128, 0, (const int64_t*)d.p, nd, nglob, ctx->glob2loc, (const double*)m_dev, inv_n,
               ctx->dflag.p, ctx->dval.p, ctx->diag.p, ctx->f.p, ctx->errflag.p);
        int e = 0;
        CU(cudaMemcpyAsync(&e, ctx->errflag.p + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (e) return toe_fail(ctx, TOE_ERR_ARG, "apply!: prescribed dof outside 1..%lld", (long long)nglob);
        if (ctx->have_K)
            LAUNCH(ctx, k_dirichlet_K, div_up(ctx->nq, 16), 128, 0, (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p,
                   (const unsigned char*)ctx->dflag.p, (const double*)ctx->dval.p, ctx->owned, ctx->val.p, ctx->ldv, ctx->nq);
        ctx->any_dirichlet = true;
        ctx->op_generation++;
    }
    TRY(T.finish());
    if (mean_out) *mean_out = msum * inv_n;
    ctx->have_solution = false;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// ABI views of dh / K in the reference's (Julia, 1-based, CSC) layout
// ---------------------------------------------------------------------------------------------------------
__global__ void k_node_first_dof(const int* __restrict__ node_q, int64_t* __restrict__ out, i64 nn) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nn) out[g] = node_q[g] >= 0 ? 3 * (int64_t)node_q[g] + 1 : 0;
}
__global__ void k_cell_dofs(const int* __restrict__ cq, int npc, i64 first0, i64 count, int64_t* __restrict__ out) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * npc) return;
    int64_t q = cq[first0 * npc + t];
    out[3 * t] = 3 * q + 1; out[3 * t + 1] = 3 * q + 2; out[3 * t + 2] = 3 * q + 3;
}
// scalar CSC pattern from the block pattern: column J = 3Q+D holds rows 3Q'+C, Q' ascending, C = 0..2
__global__ void k_scalar_pattern(const int* __restrict__ blk_ptr, const int* __restrict__ blk_col, int64_t* __restrict__ colptr,
                                 int64_t* __restrict__ rowval, int nq) {
    int q = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
    if (q > nq) return;
    int sub = threadIdx.x & 7;
    if (q == nq) { if (sub == 0) colptr[3 * (size_t)nq] = 9 * (int64_t)blk_ptr[nq] + 1; return; }
    int lo = blk_ptr[q], deg = blk_ptr[q + 1] - lo;
    if (sub < 3) colptr[3 * (size_t)q + sub] = 9 * (int64_t)lo + (int64_t)sub * 3 * deg + 1;
    for (int k = sub; k < deg; k += 8) {
        int64_t col = blk_col[lo + k];
#pragma unroll
        for (int D = 0; D < 3; D++)
#pragma unroll
            for (int C = 0; C < 3; C++) rowval[9 * (size_t)lo + (size_t)D * 3 * deg + 3 * k + C] = 3 * col + C + 1;
    }
}
// nzval in the order of k_scalar_pattern: entry (row 3Q'+C, col 3Q+D) = K[3Q+D, 3Q'+C] by symmetry
__global__ void k_scalar_values(const int* __restrict__ blk_ptr, const double* __restrict__ val, i64 nnzb, double* __restrict__ nzval, int nq) {
    int q = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
    if (q >= nq) return;
    int sub = threadIdx.x & 7;
    int lo = blk_ptr[q], deg = blk_ptr[q + 1] - lo;
    for (int k = sub; k < deg; k += 8) {
#pragma unroll
        for (int D = 0; D < 3; D++)
#pragma unroll
            for (int C = 0; C < 3; C++)
                nzval[9 * (size_t)lo + (size_t)D * 3 * deg + 3 * k + C] = val[(size_t)(3 * D + C) * nnzb + lo + k];
    }
}

int get_node_dofs(toe_ctx* ctx, int64_t* out_host) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "DOFs not built");
    TmpBuf<int64_t> d(ctx->stream); CU(d.alloc(ctx->nn));
    const int* nq_map = ctx->node_q.p;
    if (ctx->dist) TRY(dist_node_dofs(ctx, &nq_map));          // partitioned: ctx->node_q is the local map, the ABI wants the global one
    LAUNCH(ctx, k_node_first_dof, div_up(ctx->nn, 256), 256, 0, nq_map, d.p, ctx->nn);
    CU(cudaMemcpyAsync(out_host, d.p, ctx->nn * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

int get_cell_dofs(toe_ctx* ctx, i64 first, i64 count, int64_t* out_host) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "DOFs not built");
    if (first < 1 || count < 0 || first - 1 + count > ctx->ne) return toe_fail(ctx, TOE_ERR_ARG, "cell range out of bounds");
    if (count == 0) return TOE_OK;
    size_t total = (size_t)count * ctx->npc * 3;
    DevBuf<int64_t> d; CU(d.alloc(total));
    LAUNCH(ctx, k_cell_dofs, div_up(count * ctx->npc, 256), 256, 0, (const int*)ctx->cq.p, ctx->npc, first - 1, count, d.p);
    CU(cudaMemcpyAsync(out_host, d.p, total * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

int get_pattern(toe_ctx* ctx, int64_t* colptr_host, int64_t* rowval_host) {
    if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "pattern not built");
    size_t n = 3 * (size_t)ctx->nq, nnz = 9 * (size_t)ctx->nnzb;
    DevBuf<int64_t> cp, rv; CU(cp.alloc(n + 1)); CU(rv.alloc(nnz));
    LAUNCH(ctx, k_scalar_pattern, div_up((i64)ctx->nq + 1, 16), 128, 0, (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, cp.p, rv.p, ctx->nq);
    CU(cudaMemcpyAsync(colptr_host, cp.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(rowval_host, rv.p, nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

int get_values(toe_ctx* ctx, double* nzval_host) {
    if (!ctx->have_K) return toe_fail(ctx, TOE_ERR_STATE, "K not assembled");
    size_t nnz = 9 * (size_t)ctx->nnzb;
    DevBuf<double> d; CU(d.alloc(nnz));
    LAUNCH(ctx, k_scalar_values, div_up(ctx->nq, 16), 128, 0, (const int*)ctx->blk_ptr.p, (const double*)ctx->val.p, ctx->ldv, d.p, ctx->nq);
    CU(cudaMemcpyAsync(nzval_host, d.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}
