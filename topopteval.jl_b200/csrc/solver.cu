// solver.cu — kernels (3)+(4) of the north star: Jacobi-preconditioned CG with an assembled block-CSR SpMV that
// stages row products in shared memory, a matrix-free element-by-element operator, warp-shuffle dot products,
// per-element strain energy, compliance and stress recovery.
//
// Reference loops replaced: Krylov.jl cg called at RobustSolver.jl:337 with the Jacobi preconditioner of :231-236;
// energy lines FiniteElementAnalysis.jl:550 / :851, RobustSolver.jl:604 / :717; calculate_stresses :440-509 / :730-801.
#include "element.cuh"
#include <cmath>
#include <cstdlib>

// ---------------------------------------------------------------------------------------------------------
// assembled operator: block-CSR SpMV.  A CTA owns SPMV_ROWS consecutive node rows (a contiguous range of block
// slots).  Phase 1 streams the slots — thread t takes slot s0+t, s0+t+blockDim, … so the 9 value planes and the
// column indices are read as fully coalesced streams — and leaves the three row products of each slot in shared
// memory; phase 2 lets one thread per scalar row add up its staged segment.  The optional epilogue fuses the CG
// dot product p'Ap.
// ---------------------------------------------------------------------------------------------------------
static const int SPMV_ROWS = 64;           // node rows per CTA → 192 scalar rows, one per thread in phase 2
static const int SPMV_CAP = 960;           // block slots staged per pass (multiple of 4): 64 rows x 15 blocks of a structured tet mesh
static const int SPMV_LDS = SPMV_CAP + 6;  // shared-memory plane stride (even → 16-byte aligned planes; staggers the banks of the 3 product planes)

// ---- sm_100a async-copy plumbing (inline PTX): mbarrier + cp.async.bulk (SASS: UBLKCP / SYNCS) ------------------
#ifndef TOE_EMU
__device__ __forceinline__ smem_ptr_t smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(smem_ptr_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(smem_ptr_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(smem_ptr_t bar, unsigned phase) {
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    } while (!ok);
}
__device__ __forceinline__ u64 l2_evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// bytes must be a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(smem_ptr_t dst, const void* src, unsigned bytes, smem_ptr_t bar, u64 policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(smem_ptr_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_sync(int group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(256) : "memory"); }
#else
// tests/cuda_emu: the same protocol on the host (phase / arrival / tx-byte accounting in the barrier word, immediate copies)
static inline smem_ptr_t smem_u32(const void* p) { return (size_t)p; }
static inline void mbar_init(smem_ptr_t bar, unsigned count) { emu::mbar_init((void*)bar, count); }
static inline void mbar_expect_tx(smem_ptr_t bar, unsigned bytes) { emu::mbar_expect_tx((void*)bar, bytes); }
static inline void mbar_wait(smem_ptr_t bar, unsigned phase) { emu::mbar_wait((void*)bar, phase); }
static inline u64 l2_evict_first_policy() { return 0; }
static inline void bulk_g2s(smem_ptr_t dst, const void* src, unsigned bytes, smem_ptr_t bar, u64) { emu::bulk_g2s((void*)dst, src, bytes, (void*)bar); }
static inline void fence_proxy_async() {}
static inline void mbar_arrive(smem_ptr_t bar) { emu::mbar_arrive((void*)bar); }
static inline void consumer_sync(int group) { emu::named_barrier(group + 1, 256); }
#endif

// Persistent, warp-specialised, 3-stage pipelined block-CSR SpMV.
//   work item  = a chunk of `R` consecutive node rows = one contiguous slot range [s0,s1) of every value plane / blk_col
//   producer   = warp 0, one lane: waits for a free stage, arms its `full` mbarrier with the byte count and issues
//                10 bulk async copies (9 value planes + column indices, L2 evict-first) global → shared.  Two or three
//                chunks (≈73 KB each) are always in flight per SM without occupying registers.
//   consumers  = 8 warps: phase 1 — slot k of the stage: 9 values + column from shared memory, x[col] from global/L1,
//                three row products written back over the slot's entries of planes 0..2; phase 2 — one thread per
//                scalar row sums its segment, stores y and accumulates the CG dot product p'Ap in a register.
// One partial per CTA (≤148) reaches the deterministic last-block reduction.
static const int SPMV_CONS = 256;          // threads of one consumer group
static const int SPMV_GROUPS = 2;          // consumer groups working on alternate chunks (hides the x-gather latency of one behind the other)
static const int SPMV_PTHREADS = SPMV_GROUPS * SPMV_CONS + 32;
static const int SPMV_STAGES = 3;
static const int SPMV_LDC = SPMV_CAP + 8;  // column-index slots per stage: a chunk is fetched from a 4-aligned slot, so up to CAP+4 entries land
static const size_t SPMV_STAGE_BYTES = (size_t)9 * SPMV_LDS * sizeof(double) + SPMV_LDC * sizeof(int);
static const size_t SPMV_PSMEM = SPMV_STAGES * SPMV_STAGE_BYTES + 2 * SPMV_STAGES * sizeof(u64) + 32 * sizeof(double) + 16;

template <bool CG>
__global__ void __launch_bounds__(SPMV_PTHREADS, 1) k_spmv_bsr_pipe(const int* __restrict__ blk_ptr, const int* __restrict__ blk_col,
                                                                    const double* __restrict__ val, i64 ldv,
                                                                    const double* __restrict__ x, double* __restrict__ y, int nq, int R,
                                                                    const int* done_flag, CGScalars* cg, double* partials, unsigned int* counter,
                                                                    double* dot_out) {
    TOE_DYN_SMEM(unsigned char, smem_raw, 128);
    u64* bars = reinterpret_cast<u64*>(smem_raw + SPMV_STAGES * SPMV_STAGE_BYTES);      // full[0..S), empty[0..S)
    double* red = reinterpret_cast<double*>(bars + 2 * SPMV_STAGES);
    if (done_flag && *done_flag) return;
    const int tid = threadIdx.x;
    if (tid == 0) {
        // full[s]: one arrival (the producer's expect_tx) + the bytes; empty[s]: one arrival per consumer GROUP — see the consumer loop
        for (int s = 0; s < SPMV_STAGES; s++) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + SPMV_STAGES + s), SPMV_GROUPS); }
    }
    __syncthreads();
    const int nchunks = (nq + R - 1) / R;
    double dot = 0.0;
    if (tid >= SPMV_GROUPS * SPMV_CONS) {
        // ---------------- producer ----------------
        if (tid == SPMV_GROUPS * SPMV_CONS) {
            const u64 pol = l2_evict_first_policy();
            int i = 0;
            for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, i++) {
                const int st = i % SPMV_STAGES;
                const unsigned ph = (unsigned)(i / SPMV_STAGES) & 1u;
                mbar_wait(smem_u32(bars + SPMV_STAGES + st), ph ^ 1u);          // stage free (passes at once on the first lap)
                const int r0 = chunk * R, r1 = min(r0 + R, nq);
                const int s0 = __ldg(&blk_ptr[r0]), s1 = __ldg(&blk_ptr[r1]);
                const int base = s0 & ~3;
                const int cnt = ((s1 - base) + 3) & ~3;
                double* sval = reinterpret_cast<double*>(smem_raw + st * SPMV_STAGE_BYTES);
                int* scol = reinterpret_cast<int*>(sval + 9 * SPMV_LDS);
                const smem_ptr_t full = smem_u32(bars + st);
                if (cnt > 0) {
                    mbar_expect_tx(full, (unsigned)cnt * (9 * 8 + 4));
#pragma unroll
                    for (int k = 0; k < 9; k++) bulk_g2s(smem_u32(sval + k * SPMV_LDS), val + (size_t)k * ldv + base, (unsigned)cnt * 8, full, pol);
                    bulk_g2s(smem_u32(scol), blk_col + base, (unsigned)cnt * 4, full, pol);
                } else {
                    mbar_arrive(full);
                }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int group = tid / SPMV_CONS, gt = tid - group * SPMV_CONS;
        const int lr = gt / 3, c = gt - 3 * lr;
        int i = 0;
        // Every consumer group is a consumer of EVERY fill of every stage: the chunks it does not process it still waits for and
        // releases (one thread).  A parity wait means "the phase of this parity has completed", which is also true of the phase BEFORE
        // the one the waiter means while that one is still in flight — so a waiter may never be more than one phase away from the
        // barrier.  With the groups taking alternate chunks, the fill of a stage before a group's own was the OTHER group's; waiting for
        // lap L+1 without having seen lap L complete let the wait pass on lap L-1 whenever the copies of two consecutive chunks landed
        // more than a chunk's processing time out of order (≈2 in 10^6 launches at 10M tets: stale rows in y, a wrong p'Ap — the
        // round-1 "transient CG breakdown", and an illegal address when the stage still held no data at all).  The refill of a stage
        // needs the release of both groups, so no wait can be overtaken by two fills either.
        for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, i++) {
            const int st = i % SPMV_STAGES;
            const unsigned ph = (unsigned)(i / SPMV_STAGES) & 1u;
            if ((i % SPMV_GROUPS) != group) {
                if (gt == 0) { mbar_wait(smem_u32(bars + st), ph); mbar_arrive(smem_u32(bars + SPMV_STAGES + st)); }
                continue;
            }
            const int r0 = chunk * R, r1 = min(r0 + R, nq);
            const int s0 = __ldg(&blk_ptr[r0]), s1 = __ldg(&blk_ptr[r1]);
            const bool has_row = (lr < r1 - r0) && (gt < 3 * R);
            int my_lo = 0, my_hi = 0;
            double xrow = 0.0;
            const size_t row = 3 * (size_t)r0 + gt;
            if (has_row) {
                my_lo = __ldg(&blk_ptr[r0 + lr]); my_hi = __ldg(&blk_ptr[r0 + lr + 1]);
                if (CG) xrow = __ldg(&x[row]);
            }
            const int base = s0 & ~3;
            double* sval = reinterpret_cast<double*>(smem_raw + st * SPMV_STAGE_BYTES);
            const int* scol = reinterpret_cast<const int*>(sval + 9 * SPMV_LDS);
            mbar_wait(smem_u32(bars + st), ph);                                  // bytes of this stage have landed
            const int hi_s = s1 - base;
            // all x gathers of this thread's (up to 4) slots are issued before any product is formed
            const int k0 = s0 - base + gt;
            double xs[4][3];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = k0 + j * SPMV_CONS;
                const int col = k < hi_s ? scol[k] : 0;
                const double* xp = x + 3 * (size_t)col;
                xs[j][0] = __ldg(xp); xs[j][1] = __ldg(xp + 1); xs[j][2] = __ldg(xp + 2);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = k0 + j * SPMV_CONS;
                if (k < hi_s) {
                    const double* v = sval + k;
                    const double p0 = v[0] * xs[j][0] + v[SPMV_LDS] * xs[j][1] + v[2 * SPMV_LDS] * xs[j][2];
                    const double p1 = v[3 * SPMV_LDS] * xs[j][0] + v[4 * SPMV_LDS] * xs[j][1] + v[5 * SPMV_LDS] * xs[j][2];
                    const double p2 = v[6 * SPMV_LDS] * xs[j][0] + v[7 * SPMV_LDS] * xs[j][1] + v[8 * SPMV_LDS] * xs[j][2];
                    sval[k] = p0; sval[SPMV_LDS + k] = p1; sval[2 * SPMV_LDS + k] = p2;   // own slot only: no cross-thread hazard
                }
            }
            consumer_sync(group);
            if (has_row) {
                const double* pr = sval + c * SPMV_LDS - base;
                double acc = 0.0;
                for (int k = my_lo; k < my_hi; k++) acc += pr[k];
                y[row] = acc;
                if (CG) dot += acc * xrow;
            }
            // every thread orders its own generic-proxy accesses to the stage (slot products written, row sums read) before the
            // async-proxy refill; only then does the group meet and release the stage
            fence_proxy_async();
            consumer_sync(group);
            if (gt == 0) mbar_arrive(smem_u32(bars + SPMV_STAGES + st));   // stage may be refilled
        }
    }
    if (CG) {
        double d = block_sum(dot, red);
        double tot;
        if (grid_sum_last_block(d, partials, counter, red, &tot)) { if (dot_out) *dot_out = tot; else cg_after_pAp(cg, tot); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// matrix-free operator, gather form: one thread per node sums [Ke xe]_a over the cells of the node (ascending),
// recomputing the element operator from the gradients held in registers — Ke is never formed:
//   H = Σ_b x_b ⊗ g_b,  σ = λ tr(ε) I + 2μ ε,  y_a += w σ g_a
// No atomics; deterministic.  Prescribed rows return m·x (constrained operator), prescribed columns are masked.
// ---------------------------------------------------------------------------------------------------------
template <int NPC, bool CG, bool MASK>
__global__ void __launch_bounds__(128) k_ebe_gather(const int* __restrict__ inc_ptr, const int* __restrict__ inc, const int* __restrict__ cq,
                                                    const double* __restrict__ xq, Material mat,
                                                    const unsigned char* __restrict__ dflag, const double* __restrict__ dval, int any_dirichlet,
                                                    const unsigned char* __restrict__ owned, const double* __restrict__ x, double* __restrict__ y, int nq,
                                                    const int* done_flag, CGScalars* cg, double* partials, unsigned int* counter, double* dot_out) {
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    int qn = blockIdx.x * blockDim.x + threadIdx.x;
    double dotv = 0.0;
    if (qn < nq) {
        double ya[3] = {0, 0, 0};
        int lo = __ldg(&inc_ptr[qn]), hi = __ldg(&inc_ptr[qn + 1]);
        for (int i = lo; i < hi; i++) {
            int ea = __ldg(&inc[i]);
            int e = ea / NPC, a = ea - e * NPC;
            double lam, mu; material_at(mat, e, lam, mu);
            if (NPC == 4) {
                int q[4]; double X[4][3], g[4][3];
                tet_load(cq, xq, e, q, X);
                double det = tet_grads(X, g);
                double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double xb[3]; load3(x, q[b], xb);
                    if (MASK) {
#pragma unroll
                        for (int c2 = 0; c2 < 3; c2++) if (dflag[3 * (size_t)q[b] + c2]) xb[c2] = 0.0;
                    }
#pragma unroll
                    for (int c2 = 0; c2 < 3; c2++)
#pragma unroll
                        for (int i2 = 0; i2 < 3; i2++) H[c2][i2] += xb[c2] * g[b][i2];
                }
                double S[3][3]; hooke_from_grad(H, lam, mu, S);
                double ga[3] = {g[0][0], g[0][1], g[0][2]};
#pragma unroll
                for (int k = 1; k < 4; k++) if (k == a) { ga[0] = g[k][0]; ga[1] = g[k][1]; ga[2] = g[k][2]; }
                double w = det * (1.0 / 6.0);
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) ya[c2] += w * (S[c2][0] * ga[0] + S[c2][1] * ga[1] + S[c2][2] * ga[2]);
            } else {
                int q[8]; double X[8][3], xe[8][3];
                hex_load(cq, xq, e, q, X);
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    load3(x, q[b], xe[b]);
                    if (MASK) {
#pragma unroll
                        for (int c2 = 0; c2 < 3; c2++) if (dflag[3 * (size_t)q[b] + c2]) xe[b][c2] = 0.0;
                    }
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
                    double ga[3] = {g[0][0], g[0][1], g[0][2]};
#pragma unroll
                    for (int k = 1; k < 8; k++) if (k == a) { ga[0] = g[k][0]; ga[1] = g[k][1]; ga[2] = g[k][2]; }
#pragma unroll
                    for (int c2 = 0; c2 < 3; c2++) ya[c2] += det * (S[c2][0] * ga[0] + S[c2][1] * ga[1] + S[c2][2] * ga[2]);
                }
            }
        }
        double xs[3]; load3(x, qn, xs);
        if (any_dirichlet) {
#pragma unroll
            for (int c2 = 0; c2 < 3; c2++) { size_t d = 3 * (size_t)qn + c2; if (dflag[d]) ya[c2] = (!owned || owned[qn]) ? dval[d] * xs[c2] : 0.0; }
        }
#pragma unroll
        for (int c2 = 0; c2 < 3; c2++) y[3 * (size_t)qn + c2] = ya[c2];
        dotv = ya[0] * xs[0] + ya[1] * xs[1] + ya[2] * xs[2];
    }
    if (CG) {
        double d = block_sum(dotv, red);
        double tot;
        if (grid_sum_last_block(d, partials, counter, red, &tot)) { if (dot_out) *dot_out = tot; else cg_after_pAp(cg, tot); }
    }
}

double op_bytes(toe_ctx* ctx, int matrix_free) {
    double n = 3.0 * ctx->nq;
    if (matrix_free)   // conn + material per cell, coordinates per node, x read + y write
        return (double)ctx->ne * (4.0 * ctx->npc + 8.0) + ctx->nq * 24.0 + n * 16.0;
    // block-CSR: 8 B per value + 4 B per 3x3 block index + row pointers + x read + y write
    return 9.0 * ctx->nnzb * 8.0 + ctx->nnzb * 4.0 + (ctx->nq + 1) * 4.0 + n * 16.0;
}

// y = A x.  cg != null: PCG mode (early exit on cg->done, fused p'Ap).  assume_masked: x is zero on prescribed dofs.
static int op_launch(toe_ctx* ctx, const double* x, double* y, int matrix_free, CGScalars* cg, bool assume_masked, const int* done_flag = nullptr, double* dot_out = nullptr) {
    if (cg && !done_flag) done_flag = &cg->done;
    if (!matrix_free) {
        if (!ctx->have_K) return toe_fail(ctx, TOE_ERR_STATE, "assembled operator requested but K is not assembled");
        if (ctx->max_deg > SPMV_CAP) return toe_fail(ctx, TOE_ERR_MESH, "a node has %d neighbours; the SpMV stages at most %d blocks per row", ctx->max_deg, SPMV_CAP);
        int R = SPMV_CAP / (ctx->max_deg > 0 ? ctx->max_deg : 1);
        if (R > SPMV_ROWS) R = SPMV_ROWS;
        int nchunks = (ctx->nq + R - 1) / R;
        unsigned grid = (unsigned)(nchunks < N_SM ? nchunks : N_SM);          // persistent: one CTA per SM
        if (const char* eg = getenv("TOE_SPMV_GRID")) { int v = atoi(eg); if (v >= 1 && (unsigned)v < grid) grid = (unsigned)v; }   // tests: many chunks per CTA on small meshes
        if (!ctx->spmv_attr_set) {
            CU(cudaFuncSetAttribute(k_spmv_bsr_pipe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMV_PSMEM));
            CU(cudaFuncSetAttribute(k_spmv_bsr_pipe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMV_PSMEM));
            ctx->spmv_attr_set = true;
        }
        if (cg) LAUNCH(ctx, k_spmv_bsr_pipe<true>, grid, SPMV_PTHREADS, SPMV_PSMEM, (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, (const double*)ctx->val.p,
                       ctx->ldv, x, y, ctx->nq, R, done_flag, cg, ctx->partials.p, ctx->counters.p + 1, dot_out);
        else    LAUNCH(ctx, k_spmv_bsr_pipe<false>, grid, SPMV_PTHREADS, SPMV_PSMEM, (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, (const double*)ctx->val.p,
                       ctx->ldv, x, y, ctx->nq, R, done_flag, (CGScalars*)nullptr, (double*)nullptr, (unsigned int*)nullptr, (double*)nullptr);
        return TOE_OK;
    }
    if (ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "matrix-free operator requested but no material is set");
    bool mask = ctx->any_dirichlet && !assume_masked;
    static const bool use_gather = getenv("TOE_EBE_GATHER") != nullptr;       // older node-gather form, kept for comparison
    if (!use_gather) return ebe_tile_launch(ctx, x, y, cg, mask, done_flag, dot_out);
    if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "matrix-free operator needs the incidence lists (toe_build_pattern)");
    unsigned grid = div_up(ctx->nq, 128);
#define EBE_ARGS (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p, (const double*)ctx->xq.p, ctx->mat, \
        (const unsigned char*)ctx->dflag.p, (const double*)ctx->dval.p, (int)ctx->any_dirichlet, ctx->owned, x, y, ctx->nq, done_flag, cg, ctx->partials.p, ctx->counters.p + 1, dot_out
    if (ctx->npc == 4) {
        if (cg) { if (mask) LAUNCH(ctx, (k_ebe_gather<4, true, true>), grid, 128, 0, EBE_ARGS); else LAUNCH(ctx, (k_ebe_gather<4, true, false>), grid, 128, 0, EBE_ARGS); }
        else    { if (mask) LAUNCH(ctx, (k_ebe_gather<4, false, true>), grid, 128, 0, EBE_ARGS); else LAUNCH(ctx, (k_ebe_gather<4, false, false>), grid, 128, 0, EBE_ARGS); }
    } else {
        if (cg) { if (mask) LAUNCH(ctx, (k_ebe_gather<8, true, true>), grid, 128, 0, EBE_ARGS); else LAUNCH(ctx, (k_ebe_gather<8, true, false>), grid, 128, 0, EBE_ARGS); }
        else    { if (mask) LAUNCH(ctx, (k_ebe_gather<8, false, true>), grid, 128, 0, EBE_ARGS); else LAUNCH(ctx, (k_ebe_gather<8, false, false>), grid, 128, 0, EBE_ARGS); }
    }
#undef EBE_ARGS
    return TOE_OK;
}

int op_apply(toe_ctx* ctx, const double* x, double* y, int matrix_free, double* /*unused*/, bool assume_masked) {
    TRY(op_launch(ctx, x, y, matrix_free, nullptr, assume_masked));
    return dist_post_spmv(ctx, y);
}

// ---------------------------------------------------------------------------------------------------------
// PCG vector kernels
// ---------------------------------------------------------------------------------------------------------
static const int VEC_THREADS = 256;
static unsigned vec_grid(size_t n) { return min_u(div_up((i64)n, VEC_THREADS), (unsigned)(N_SM * 8)); }

// x = 0, r = f, Minv = 1 ./ D with D[abs(D) < 1e-12] = 1 (RobustSolver.jl:231-236), z = M r, p = z, γ = r'z
// L2: the stopping test uses ‖r‖₂ instead of Krylov.jl's √(r'Mr) (TOE_PCG_L2_NORM)
template <bool L2>
__global__ void __launch_bounds__(VEC_THREADS) k_cg_init(const double* __restrict__ f, const double* __restrict__ diag, double* __restrict__ Minv,
                                                         double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, size_t n,
                                                         CGScalars* cg, double atol, double rtol, i64 itmax, double* hist, i64 hist_cap,
                                                         double* partials, unsigned int* counter) {
    __shared__ double red[32];
    double s = 0.0, s2 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double d = diag[i];
        if (fabs(d) < 1e-12) d = 1.0;
        double mi = 1.0 / d, ri = f[i], zi = mi * ri;
        Minv[i] = mi; x[i] = 0.0; r[i] = ri; p[i] = zi;
        s += ri * zi;
        if (L2) s2 += ri * ri;
    }
    s = block_sum(s, red);
    if (L2) s2 = block_sum(s2, red);
    double tot, tot2 = 0.0;
    if (L2 ? grid_sum2_last_block(s, s2, partials, counter, red, &tot, &tot2) : grid_sum_last_block(s, partials, counter, red, &tot)) {
        cg->gamma = tot; cg->pAp = 0.0; cg->beta = 0.0;
        cg->res0 = sqrt(L2 ? tot2 : tot);
        cg->res = cg->res0;
        cg->eps = atol + rtol * cg->res0;
        cg->iter = 0; cg->itmax = itmax;
        cg->converged = (cg->res0 <= cg->eps) ? 1 : 0;
        cg->done = (cg->converged || itmax <= 0) ? 1 : 0;
        cg->breakdown = 0;
        if (hist_cap > 0) hist[0] = cg->res0;
    }
}

// α = γ/p'Ap, x += α p, r -= α Ap, γ' = r' M r; the last block closes the iteration (β, convergence test)
template <bool L2>
__global__ void __launch_bounds__(VEC_THREADS) k_cg_xr(const double* __restrict__ p, const double* __restrict__ Ap, const double* __restrict__ Minv,
                                                       double* __restrict__ x, double* __restrict__ r, size_t n, CGScalars* cg,
                                                       double* hist, i64 hist_cap, double* partials, unsigned int* counter) {
    __shared__ double red[32];
    if (cg->done) return;
    const double alpha = cg->gamma / cg->pAp;
    double s = 0.0, s2 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double pi = p[i];
        x[i] += alpha * pi;
        double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        s += ri * ri * Minv[i];
        if (L2) s2 += ri * ri;
    }
    s = block_sum(s, red);
    if (L2) s2 = block_sum(s2, red);
    double tot, tot2 = -1.0;
    if (L2 ? grid_sum2_last_block(s, s2, partials, counter, red, &tot, &tot2) : grid_sum_last_block(s, partials, counter, red, &tot))
        cg_after_gamma(cg, tot, hist, hist_cap, tot2);
}

// p = M r + β p
__global__ void __launch_bounds__(VEC_THREADS) k_cg_p(const double* __restrict__ r, const double* __restrict__ Minv, double* __restrict__ p,
                                                      size_t n, const CGScalars* cg) {
    if (cg->done) return;
    const double beta = cg->beta;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = Minv[i] * r[i] + beta * p[i];
}

// out[0] = Σ (a-b)^2, out[1] = Σ a^2, out[2] = Σ a*b
__global__ void __launch_bounds__(VEC_THREADS) k_norms(const double* __restrict__ a, const double* __restrict__ b, size_t n, double* out,
                                                       double* partials, unsigned int* counter, const unsigned char* __restrict__ owned) {
    __shared__ double red[32];
    double s0 = 0, s1 = 0, s2 = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (owned && !owned[i / 3]) continue;
        double ai = a[i], bi = b[i], d = ai - bi;
        s0 += d * d; s1 += ai * ai; s2 += ai * bi;
    }
    double tot;
    s0 = block_sum(s0, red);
    if (grid_sum_last_block(s0, partials, counter, red, &tot)) out[0] = tot;
    s1 = block_sum(s1, red);
    if (grid_sum_last_block(s1, partials + gridDim.x, counter + 1, red, &tot)) out[1] = tot;
    s2 = block_sum(s2, red);
    if (grid_sum_last_block(s2, partials + 2 * gridDim.x, counter + 2, red, &tot)) out[2] = tot;
}

// out = Σ_owned a·b
__global__ void __launch_bounds__(VEC_THREADS) k_dot_masked(const double* __restrict__ a, const double* __restrict__ b, size_t n,
                                                            const unsigned char* __restrict__ owned, const int* done_flag,
                                                            double* partials, unsigned int* counter, double* out) {
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (!owned || owned[i / 3]) s += a[i] * b[i];
    s = block_sum(s, red);
    double tot;
    if (grid_sum_last_block(s, partials, counter, red, &tot)) *out = tot;
}

// ---------------------------------------------------------------------------------------------------------
// Partitioned runs: Chronopoulos–Gear form of the same preconditioned CG — one merged reduction {γ=r'u, δ=w'u} and hence
// ONE ncclAllReduce per iteration (standard CG needs two).  Mathematically identical iterates; x₀=0, M=diag(K)⁻¹,
// same stopping rule on √γ.  Scalars are double-buffered by iteration parity so that no thread reads a value another
// thread of the same launch rewrites.
//   k_cgcg_vec(j):  β=γ_j/γ_{j-1}, α=γ_j/(δ_j-βγ_j/α_{j-1});  p=u+βp; s=w+βs; x+=αp; r-=αs; u=Mr   (one pass over 7 vectors)
//   operator:       w = A u  (+ interface sum)
//   γ_{j+1} partial (owner-masked) is formed inside k_cgcg_vec, the δ_{j+1} partial inside the operator kernel (all local rows);
//   both land in slot (j+1)&1 and are summed over the ranks by the exchange step (peer-memory kernel or NCCL group)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VEC_THREADS) k_cgcg_init(const double* __restrict__ f, const double* __restrict__ diag, double* __restrict__ Minv,
                                                           double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                           double* __restrict__ p, double* __restrict__ s, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double d = diag[i];
        if (fabs(d) < 1e-12) d = 1.0;
        double mi = 1.0 / d, ri = f[i];
        Minv[i] = mi; x[i] = 0.0; r[i] = ri; z[i] = mi * ri; p[i] = 0.0; s[i] = 0.0;
    }
}
__global__ void k_cgcg_fin_init(CGScalars* cg, double atol, double rtol, i64 itmax, double* hist, i64 hist_cap, int l2) {
    if (threadIdx.x || blockIdx.x) return;
    double g = cg->gd[0][0];
    cg->gamma = g; cg->res0 = sqrt(l2 ? cg->gd[0][2] : g); cg->res = cg->res0; cg->eps = atol + rtol * cg->res0;
    cg->iter = 0; cg->itmax = itmax; cg->breakdown = 0; cg->beta = 0.0; cg->pAp = 0.0;
    cg->converged = 0; cg->done = 0;
    cg->alpha[0] = cg->alpha[1] = 1.0;
    cg->gd[1][0] = g; cg->gd[1][1] = 0.0;
}
__global__ void __launch_bounds__(VEC_THREADS) k_cgcg_vec(const double* __restrict__ Minv, const double* __restrict__ w, double* __restrict__ z,
                                                          double* __restrict__ p, double* __restrict__ s, double* __restrict__ x, double* __restrict__ r,
                                                          size_t n, CGScalars* cg, int par, double* hist, i64 hist_cap,
                                                          const unsigned char* __restrict__ owned, double* partials, unsigned int* counter, int l2,
                                                          double* trace) {
    __shared__ double red[32];
    if (cg->done) return;
    const i64 j = cg->iter;                        // advanced once, by the last block of this launch, after every block has read it
    const double g = cg->gd[par][0], dl = cg->gd[par][1];
    const double gprev = cg->gd[par ^ 1][0], aprev = cg->alpha[par ^ 1];
    const bool first = (j == 0);
    const double res = sqrt(l2 ? cg->gd[par][2] : g);
    const bool conv = res <= cg->eps, tired = j >= cg->itmax;
    const double beta = first ? 0.0 : g / gprev;
    const double denom = first ? dl : dl - beta * g / aprev;       // = p'Ap in exact arithmetic
    const bool brk = !(denom > 0.0) && !conv && !tired;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (j < hist_cap) hist[j] = res;
        if (trace && j < hist_cap) { trace[4 * j] = g; trace[4 * j + 1] = dl; }
        cg->gamma = g; cg->res = res;
        if (conv) { cg->converged = 1; cg->done = 1; }
        else if (tired) cg->done = 1;
        else if (brk) { cg->breakdown = 1; cg->done = 1; }
        else cg->alpha[par] = g / denom;
    }
    if (conv || tired || brk) return;
    const double alpha = g / denom;
    double gs = 0.0, ns = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double pi = z[i] + beta * p[i];
        double si = w[i] + beta * s[i];
        p[i] = pi; s[i] = si;
        x[i] += alpha * pi;
        double ri = r[i] - alpha * si;
        r[i] = ri;
        double zi = Minv[i] * ri;
        z[i] = zi;
        if (!owned || owned[i / 3]) { gs += ri * zi; ns += ri * ri; }
    }
    // γ_{j+1} = r'Mr (and ν_{j+1} = r'r) over owned dofs: local partials into the other parity slot (summed over the ranks after the operator)
    gs = block_sum(gs, red);
    ns = block_sum(ns, red);
    double tot, tot2;
    if (grid_sum2_last_block(gs, ns, partials, counter, red, &tot, &tot2)) {
        cg->gd[par ^ 1][0] = tot; cg->gd[par ^ 1][2] = tot2; cg->iter = j + 1;
        if (trace && j + 1 < hist_cap) trace[4 * (j + 1) + 2] = tot;
    }
}
// diagnostic: this rank's δ partial of the iteration that is being closed (after the operator, before the exchange)
__global__ void k_trace_delta(const CGScalars* cg, int par, double* trace, i64 cap) {
    if (threadIdx.x || blockIdx.x || cg->done) return;
    const i64 j = cg->iter;                      // already advanced by k_cgcg_vec
    if (j < cap) trace[4 * j + 3] = cg->gd[par ^ 1][1];
}
int dist_exchange_allreduce(toe_ctx* ctx, double* y, double* scal, int count);
int dist_check_exchange(toe_ctx* ctx);     // dist.cu: interface sum of y + allreduce of scal in ONE NCCL group

// one iteration of the partitioned single-reduction CG; `par` = iteration parity (baked into captured graphs).
// 4 kernels + one NCCL group: the δ partial is the operator's own fused dot over ALL local rows (Σ_ranks w_local·z equals
// w·z because z is interface-consistent), so the allreduce does not have to wait for the interface sum.
static int cgcg_iteration(toe_ctx* ctx, int matrix_free, size_t n, i64 hist_cap, int par, int l2) {
    CGScalars* cg = ctx->cgs.p;
    double* z = ctx->cg_z.p; double* w = ctx->Ap.p;
    LAUNCH(ctx, k_cgcg_vec, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->Minv.p, (const double*)w, z, ctx->p.p, ctx->cg_s.p, ctx->u.p, ctx->r.p,
           n, cg, par, ctx->hist.p, hist_cap, ctx->owned, ctx->partials.p, ctx->counters.p + 8, l2, ctx->cg_trace.p);
    TRY(op_launch(ctx, z, w, matrix_free, cg, true, &cg->done, &cg->gd[par ^ 1][1]));
    if (ctx->cg_trace.p) LAUNCH(ctx, k_trace_delta, 1, 32, 0, (const CGScalars*)cg, par, ctx->cg_trace.p, hist_cap);
    TRY(dist_exchange_allreduce(ctx, w, &cg->gd[par ^ 1][0], l2 ? 3 : 2));
    return TOE_OK;
}

static int cg_iteration(toe_ctx* ctx, int matrix_free, size_t n, i64 hist_cap, bool two_level = false, int l2 = 0) {
    CGScalars* cg = ctx->cgs.p;
    if (two_level) {
        if (ctx->dist) {     // partitioned: the local p'Ap partial (all local rows: p is interface-consistent) rides with the interface sum
            TRY(op_launch(ctx, ctx->p.p, ctx->Ap.p, matrix_free, cg, true, &cg->done, &cg->gd[0][0]));
            TRY(dist_exchange_allreduce(ctx, ctx->Ap.p, &cg->gd[0][0], 1));
        } else {
            TRY(op_launch(ctx, ctx->p.p, ctx->Ap.p, matrix_free, cg, true));
        }
        return tl_cg_after_operator(ctx, hist_cap);
    }
    {
        TRY(op_launch(ctx, ctx->p.p, ctx->Ap.p, matrix_free, cg, true));
#define XR_ARGS (const double*)ctx->p.p, (const double*)ctx->Ap.p, (const double*)ctx->Minv.p, ctx->u.p, ctx->r.p, n, cg, ctx->hist.p, hist_cap, ctx->partials.p, ctx->counters.p + 2
        if (l2) LAUNCH(ctx, k_cg_xr<true>, vec_grid(n), VEC_THREADS, 0, XR_ARGS);
        else    LAUNCH(ctx, k_cg_xr<false>, vec_grid(n), VEC_THREADS, 0, XR_ARGS);
#undef XR_ARGS
    }
    LAUNCH(ctx, k_cg_p, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->r.p, (const double*)ctx->Minv.p, ctx->p.p, n, (const CGScalars*)cg);
    return TOE_OK;
}

static const int CG_BATCH_DEFAULT = 50;       // iterations per host check / captured graph (always even: parity-indexed scalars)
static const i64 HIST_CAP = 1LL << 20;

// Σ_owned w_i (a_i - b_i)^2  (w = null: plain sum of squares)
__global__ void __launch_bounds__(VEC_THREADS) k_wdiff2(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ w, size_t n,
                                                        const unsigned char* __restrict__ owned, double* partials, unsigned int* counter, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (owned && !owned[i / 3]) continue;
        double d = a[i] - b[i];
        s += (w ? w[i] : 1.0) * d * d;
    }
    s = block_sum(s, red);
    double tot;
    if (grid_sum_last_block(s, partials, counter, red, &tot)) *out = tot;
}

int solve_pcg(toe_ctx* ctx, double atol, double rtol, i64 itmax, int flags, toe_pcg_stats* stats, double* history, i64 history_cap) {
    int matrix_free = (flags & TOE_PCG_MATRIX_FREE) ? 1 : 0;
    const int l2 = (flags & TOE_PCG_L2_NORM) ? 1 : 0;
    if (!matrix_free && !ctx->have_K) return toe_fail(ctx, TOE_ERR_STATE, "solve: K is not assembled (or pass TOE_PCG_MATRIX_FREE)");
    if (matrix_free && ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "solve: no material set");
    if (itmax < 0) return toe_fail(ctx, TOE_ERR_ARG, "solve: itmax must be >= 0");
    const bool dist = ctx->dist != nullptr;
    const bool two_level = (flags & TOE_PCG_TWO_LEVEL) != 0;
    if (l2 && two_level) return toe_fail(ctx, TOE_ERR_ARG, "solve: TOE_PCG_L2_NORM is implemented for the Jacobi preconditioner only");
    TRY(ensure_vectors(ctx));
    TRY(compute_diag(ctx));
    if (matrix_free && !getenv("TOE_EBE_GATHER")) TRY(mesh_build_tiles(ctx));     // allocations must not happen inside graph capture
    size_t n = 3 * (size_t)ctx->nq;
    const i64 hist_cap = HIST_CAP;            // fixed: pointer and capacity are baked into the captured graph
    CU(ctx->hist.alloc(hist_cap));
    if (dist) { CU(ctx->cg_s.alloc(n)); CU(ctx->cg_z.alloc(n)); }
    if (dist && getenv("TOE_CG_TRACE")) { CU(ctx->cg_trace.alloc(4 * (size_t)hist_cap)); CU(cudaMemsetAsync(ctx->cg_trace.p, 0, 4 * (size_t)hist_cap * sizeof(double), ctx->stream)); }
    if (!ctx->cgs_host) CU(cudaMallocHost((void**)&ctx->cgs_host, sizeof(CGScalars)));
    i64 launches0 = ctx->launches;
    const int per_iter = two_level ? 6 : (dist ? 4 : 3);
    int coarse_dofs = 0; double precond_seconds = 0.0;
    int CG_BATCH = CG_BATCH_DEFAULT;
    if (const char* eb = getenv("TOE_CG_BATCH")) { int v = atoi(eb); if (v >= 2) CG_BATCH = v & ~1; }

    if (dist) TRY(dist_align(ctx));            // the ranks start the iteration's exchange sequence in lock-step (outside the timed solve)
    EventPair ev; CU(ev.create());
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    CU(cudaEventRecord(e0, ctx->stream));
    if (two_level) TRY(tl_prepare(ctx, matrix_free, &coarse_dofs, &precond_seconds));      // ZᵀKZ and its inverse: part of the solve time
    if (two_level) {
        TRY(tl_cg_init(ctx, atol, rtol, itmax, hist_cap));
    } else if (!dist) {
#define INIT_ARGS (const double*)ctx->f.p, (const double*)ctx->diag.p, ctx->Minv.p, ctx->u.p, ctx->r.p, ctx->p.p, n, ctx->cgs.p, atol, rtol, itmax, ctx->hist.p, hist_cap, ctx->partials.p, ctx->counters.p
        if (l2) LAUNCH(ctx, k_cg_init<true>, vec_grid(n), VEC_THREADS, 0, INIT_ARGS);
        else    LAUNCH(ctx, k_cg_init<false>, vec_grid(n), VEC_THREADS, 0, INIT_ARGS);
#undef INIT_ARGS
    } else {
        CU(cudaMemsetAsync(ctx->errflag.p + 2, 0, sizeof(int), ctx->stream));
        CU(cudaMemsetAsync(ctx->cgs.p, 0, sizeof(CGScalars), ctx->stream));
        LAUNCH(ctx, k_cgcg_init, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->f.p, (const double*)ctx->diag.p, ctx->Minv.p, ctx->u.p, ctx->r.p, ctx->cg_z.p,
               ctx->p.p, ctx->cg_s.p, n);
        LAUNCH(ctx, k_dot_masked, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->r.p, (const double*)ctx->cg_z.p, n, ctx->owned, (const int*)nullptr,
               ctx->partials.p, ctx->counters.p + 8, &ctx->cgs.p->gd[0][0]);
        if (l2) LAUNCH(ctx, k_dot_masked, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->r.p, (const double*)ctx->r.p, n, ctx->owned, (const int*)nullptr,
                       ctx->partials.p, ctx->counters.p + 8, &ctx->cgs.p->gd[0][2]);
        TRY(op_launch(ctx, ctx->cg_z.p, ctx->Ap.p, matrix_free, ctx->cgs.p, true, &ctx->cgs.p->done, &ctx->cgs.p->gd[0][1]));
        TRY(dist_exchange_allreduce(ctx, ctx->Ap.p, &ctx->cgs.p->gd[0][0], l2 ? 3 : 2));
        LAUNCH(ctx, k_cgcg_fin_init, 1, 32, 0, ctx->cgs.p, atol, rtol, itmax, ctx->hist.p, hist_cap, l2);
    }

    auto one_iteration = [&](int k) -> int {
        return (dist && !two_level) ? cgcg_iteration(ctx, matrix_free, n, hist_cap, k & 1, l2) : cg_iteration(ctx, matrix_free, n, hist_cap, two_level, l2);
    };
    bool use_graph = !(flags & TOE_PCG_NO_GRAPH);
    if (dist && !getenv("TOE_DIST_GRAPH")) use_graph = false;     // NCCL inside stream capture: opt-in (TOE_DIST_GRAPH=1) — 2-4 % faster, but processes did not exit cleanly in two runs
    i64 key = ((ctx->op_generation * 2 + l2) * 8 + (two_level ? 4 : 0) + matrix_free * 2 + 1) * 4096 + CG_BATCH;
    if (use_graph && (ctx->graph_key != key || !ctx->graph_exec)) {
        if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(ctx->stream, dist ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal));
        i64 l0 = ctx->launches;
        int st = TOE_OK;
        for (int k = 0; k < CG_BATCH && st == TOE_OK; k++) st = one_iteration(k);
        ctx->launches = l0;
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
        if (st != TOE_OK) { if (g) cudaGraphDestroy(g); return st; }
        if (ce != cudaSuccess) return toe_fail(ctx, TOE_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return toe_fail(ctx, TOE_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
        ctx->graph_key = key;
    }
    i64 max_batches = (itmax + CG_BATCH - 1) / CG_BATCH + 1;
    for (i64 bt = 0; bt < max_batches; bt++) {
        if (use_graph) { CU(cudaGraphLaunch(ctx->graph_exec, ctx->stream)); ctx->launches += per_iter * CG_BATCH; }
        else for (int k = 0; k < CG_BATCH; k++) TRY(one_iteration(k));
        CU(cudaMemcpyAsync(ctx->cgs_host, ctx->cgs.p, sizeof(CGScalars), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->cgs_host->done) break;
    }
    CU(cudaEventRecord(e1, ctx->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    CU(cudaGetLastError());
    CGScalars h = *ctx->cgs_host;
    ctx->tm.solve = ms * 1e-3;
    if (dist) TRY(dist_check_exchange(ctx));
    ctx->have_solution = true;

    // true residual f - K u, in the l2 norm (RobustSolver.jl:468) and in the norm of the stopping test.  A converged recurrence whose true
    // residual is orders of magnitude away means the iterates were corrupted on the way (a faulty exchange cannot make p'Ap <= 0 by
    // itself): that is an error, never a result.
    TRY(op_launch(ctx, ctx->u.p, ctx->tmp.p, matrix_free, nullptr, true));
    TRY(dist_post_spmv(ctx, ctx->tmp.p));
    double* out = ctx->partials.p + 3 * (N_SM * 8) + 8;
    LAUNCH(ctx, k_norms, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->f.p, (const double*)ctx->tmp.p, n, out, ctx->partials.p, ctx->counters.p + 4, ctx->owned);
    LAUNCH(ctx, k_wdiff2, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->f.p, (const double*)ctx->tmp.p, (const double*)(two_level ? nullptr : ctx->Minv.p), n,
           ctx->owned, ctx->partials.p, ctx->counters.p + 7, out + 3);
    TRY(dist_allreduce(ctx, out, 4));
    double hn[4];
    CU(cudaMemcpyAsync(hn, out, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const double rel_res_l2 = hn[1] > 0 ? sqrt(hn[0] / hn[1]) : sqrt(hn[0]);
    const double true_res = (l2 || two_level) ? sqrt(hn[0]) : sqrt(hn[3]);
    const double final_res = two_level ? sqrt(h.gamma) : h.res;
    if (stats) {
        stats->niter = h.iter; stats->converged = h.converged; stats->breakdown = h.breakdown;
        stats->res0_M = h.res0; stats->res_M = final_res;
        stats->solve_seconds = ms * 1e-3;
        stats->rel_res_l2 = rel_res_l2;
        stats->true_res = true_res;
        // operator time: a few isolated launches (local product only)
        const int reps = 5;
        EventPair ev2; CU(ev2.create());
        cudaEvent_t a = ev2.a, b = ev2.b;
        TRY(op_launch(ctx, ctx->u.p, ctx->tmp.p, matrix_free, nullptr, true));
        CU(cudaEventRecord(a, ctx->stream));
        for (int k = 0; k < reps; k++) TRY(op_launch(ctx, ctx->u.p, ctx->tmp.p, matrix_free, nullptr, true));
        CU(cudaEventRecord(b, ctx->stream));
        CU(cudaEventSynchronize(b));
        float oms = 0; cudaEventElapsedTime(&oms, a, b);
        stats->spmv_seconds = (double)h.iter * (oms * 1e-3 / reps);
        stats->spmv_bytes = op_bytes(ctx, matrix_free);
        stats->kernel_launches = ctx->launches - launches0;
        stats->restarts = 0;
        stats->coarse_dofs = coarse_dofs; stats->precond_seconds = precond_seconds;
    }
    if (history && history_cap > 0) {
        i64 cnt = h.iter + 1 < history_cap ? h.iter + 1 : history_cap;
        if (cnt > hist_cap) cnt = hist_cap;
        CU(cudaMemcpy(history, ctx->hist.p, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (h.converged && !two_level && !(true_res <= 1e4 * fmax(h.eps, final_res)))
        return toe_fail(ctx, TOE_ERR_NUMERIC, "solve: the recurrence reports convergence (residual %.3e <= %.3e after %lld iterations) but the true residual "
                        "of the returned u is %.3e in the same norm: the iterates were corrupted", final_res, h.eps, (long long)h.iter, true_res);
    return TOE_OK;
}

// number of entries of a that differ from b (bitwise)
__global__ void __launch_bounds__(VEC_THREADS) k_count_diff(const double* __restrict__ a, const double* __restrict__ b, size_t n, unsigned long long* count) {
    unsigned long long c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (__double_as_longlong(a[i]) != __double_as_longlong(b[i])) c++;
    if (c) atomicAdd(count, c);
}

// Soak test of the operator (diagnostic): the product of the stored vector u is formed once, then `reps` more times, each result
// compared bit for bit with the first on the device.  Every kernel of the path is deterministic, so any mismatch is a fault of the
// machinery (pipeline protocol, exchange, hardware), not rounding.  what: 1 = local product only, 2 = interface exchange only (the
// same local product is re-exchanged), 3 = product + exchange (what one PCG iteration does); 2 and 3 need a partitioned ctx.
int spmv_soak(toe_ctx* ctx, int matrix_free, int what, i64 reps, i64* mismatching_batches, i64* mismatching_entries) {
    TRY(ensure_vectors(ctx));
    if (!ctx->have_solution) return toe_fail(ctx, TOE_ERR_STATE, "toe_spmv_soak: needs a stored vector (solve or toe_set_solution first)");
    if (what < 1 || what > 3) return toe_fail(ctx, TOE_ERR_ARG, "toe_spmv_soak: what must be 1, 2 or 3");
    size_t n = 3 * (size_t)ctx->nq;
    DevBuf<unsigned long long> cnt; CU(cnt.alloc(2));
    CU(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
    TRY(dist_align(ctx));
    double* local_ref = ctx->r.p; double* full_ref = ctx->tmp.p; double* work = ctx->Ap.p;
    TRY(op_launch(ctx, ctx->u.p, local_ref, matrix_free, nullptr, false));
    CU(cudaMemcpyAsync(full_ref, local_ref, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    TRY(dist_post_spmv(ctx, full_ref));
    const double* ref = (what == 1) ? local_ref : full_ref;
    i64 bad = 0;
    unsigned long long prev = 0, h[2];
    for (i64 k = 0; k < reps; k++) {
        if (what & 1) TRY(op_launch(ctx, ctx->u.p, work, matrix_free, nullptr, false));
        else CU(cudaMemcpyAsync(work, local_ref, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        if (what & 2) TRY(dist_post_spmv(ctx, work));
        LAUNCH(ctx, k_count_diff, vec_grid(n), VEC_THREADS, 0, (const double*)work, ref, n, cnt.p);
        if ((k & 255) == 255 || k == reps - 1) {              // host check every 256 applications
            CU(cudaMemcpyAsync(h, cnt.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (h[0] != prev) { bad++; prev = h[0]; }
        }
    }
    if (mismatching_batches) *mismatching_batches = bad;
    if (mismatching_entries) *mismatching_entries = (i64)prev;
    return TOE_OK;
}

int cg_trace(toe_ctx* ctx, double* out, i64 iterations) {
    if (!ctx->cg_trace.p) return toe_fail(ctx, TOE_ERR_STATE, "no CG trace: set TOE_CG_TRACE=1 before a partitioned solve");
    if (iterations > HIST_CAP) iterations = HIST_CAP;
    CU(cudaMemcpy(out, ctx->cg_trace.p, 4 * (size_t)iterations * sizeof(double), cudaMemcpyDeviceToHost));
    return TOE_OK;
}

int time_spmv(toe_ctx* ctx, int matrix_free, int reps, double* seconds_out, double* bytes_out) {
    TRY(ensure_vectors(ctx));
    if (reps < 1) reps = 1;
    EventPair ev; CU(ev.create());
    cudaEvent_t a = ev.a, b = ev.b;
    TRY(op_launch(ctx, ctx->u.p, ctx->tmp.p, matrix_free, nullptr, false));
    CU(cudaEventRecord(a, ctx->stream));
    for (int k = 0; k < reps; k++) TRY(op_launch(ctx, ctx->u.p, ctx->tmp.p, matrix_free, nullptr, false));
    CU(cudaEventRecord(b, ctx->stream));
    CU(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    CU(cudaGetLastError());
    if (seconds_out) *seconds_out = ms * 1e-3 / reps;
    if (bytes_out) *bytes_out = op_bytes(ctx, matrix_free);
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// energy: eₑ = ½ uₑᵀ Kₑ uₑ = ½ Σ_q (λ tr(ε)² + 2μ ε:ε) dΩ, Σ eₑ, compliance f'u
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double strain_energy_density(const double H[3][3], double lam, double mu) {
    double tr = H[0][0] + H[1][1] + H[2][2];
    double exy = 0.5 * (H[0][1] + H[1][0]), eyz = 0.5 * (H[1][2] + H[2][1]), exz = 0.5 * (H[0][2] + H[2][0]);
    double ee = H[0][0] * H[0][0] + H[1][1] * H[1][1] + H[2][2] * H[2][2] + 2.0 * (exy * exy + eyz * eyz + exz * exz);
    return 0.5 * (lam * tr * tr + 2.0 * mu * ee);
}

template <int NPC>
__global__ void __launch_bounds__(128) k_elem_energy(const int* __restrict__ cq, const double* __restrict__ xq, Material mat,
                                                     const double* __restrict__ u, double* __restrict__ ee_out, int ne,
                                                     double* partials) {
    __shared__ double red[32];
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    double en = 0.0;
    if (e < ne) {
        double lam, mu; material_at(mat, e, lam, mu);
        if (NPC == 4) {
            int q[4]; double X[4][3], g[4][3];
            tet_load(cq, xq, e, q, X);
            double det = tet_grads(X, g);
            double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int b = 0; b < 4; b++) {
                double ub[3]; load3(u, q[b], ub);
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int i = 0; i < 3; i++) H[c][i] += ub[c] * g[b][i];
            }
            en = strain_energy_density(H, lam, mu) * det * (1.0 / 6.0);
        } else {
            int q[8]; double X[8][3], ue[8][3];
            hex_load(cq, xq, e, q, X);
#pragma unroll
            for (int b = 0; b < 8; b++) load3(u, q[b], ue[b]);
            for (int gp = 0; gp < 8; gp++) {
                double g[8][3], N[8];
                double det = hex_grads_at(X, gp, g, N);
                double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
                for (int b = 0; b < 8; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++)
#pragma unroll
                        for (int i = 0; i < 3; i++) H[c][i] += ue[b][c] * g[b][i];
                en += strain_energy_density(H, lam, mu) * det;
            }
        }
        if (ee_out) ee_out[e] = en;
    }
    double s = block_sum(en, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// fixed-order two-level sum of n partials (n may exceed one block's reach)
__global__ void k_sum_strided(const double* __restrict__ in, i64 n, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
}

int energy(toe_ctx* ctx, double* half_uKu, double* compliance, double* per_elem_host) {
    if (!ctx->have_dofs || ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "energy: mesh/material not set");
    TRY(ensure_vectors(ctx));
    size_t n = 3 * (size_t)ctx->nq;
    int ne = (int)ctx->ne;
    StageTimer T(ctx, &ctx->tm.energy);
    unsigned grid = div_up(ne, 128);
    TmpBuf<double> part(ctx->stream); CU(part.alloc(grid + 4));
    TmpBuf<double> ee(ctx->stream);
    if (per_elem_host) CU(ee.alloc(ne));
    double* eep = per_elem_host ? ee.p : nullptr;
    if (ctx->npc == 4) LAUNCH(ctx, k_elem_energy<4>, grid, 128, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, ctx->mat, (const double*)ctx->u.p, eep, ne, part.p);
    else               LAUNCH(ctx, k_elem_energy<8>, grid, 128, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, ctx->mat, (const double*)ctx->u.p, eep, ne, part.p);
    LAUNCH(ctx, k_sum_strided, 1, 1024, 0, (const double*)part.p, (i64)grid, part.p + grid);
    double* out = ctx->partials.p + 3 * (N_SM * 8) + 8;
    LAUNCH(ctx, k_norms, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->f.p, (const double*)ctx->u.p, n, out, ctx->partials.p, ctx->counters.p + 4, ctx->owned);
    TRY(dist_allreduce(ctx, part.p + grid, 1));
    TRY(dist_allreduce(ctx, out, 3));
    double h_e = 0, hn[3];
    CU(cudaMemcpyAsync(&h_e, part.p + grid, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hn, out, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (per_elem_host) {
        if (ctx->dist) TRY(dist_sum_per_element(ctx, ee.p, per_elem_host));
        else CU(cudaMemcpyAsync(per_elem_host, ee.p, ne * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRY(T.finish());
    if (half_uKu) *half_uKu = h_e;
    if (compliance) *compliance = hn[2];
    return TOE_OK;
}

int energy_assembled(toe_ctx* ctx, double* half_uKu) {
    if (!ctx->have_K) return toe_fail(ctx, TOE_ERR_STATE, "energy_assembled: K not assembled");
    TRY(ensure_vectors(ctx));
    size_t n = 3 * (size_t)ctx->nq;
    TRY(op_apply(ctx, ctx->u.p, ctx->tmp.p, 0, nullptr, false));
    double* out = ctx->partials.p + 3 * (N_SM * 8) + 8;
    LAUNCH(ctx, k_norms, vec_grid(n), VEC_THREADS, 0, (const double*)ctx->u.p, (const double*)ctx->tmp.p, n, out, ctx->partials.p, ctx->counters.p + 4, ctx->owned);
    TRY(dist_allreduce(ctx, out, 3));
    double hn[3];
    CU(cudaMemcpyAsync(hn, out, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (half_uKu) *half_uKu = 0.5 * hn[2];
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stress recovery (calculate_stresses, FiniteElementAnalysis.jl:440-509 / :730-801)
// ---------------------------------------------------------------------------------------------------------
template <int NPC>
__global__ void __launch_bounds__(128) k_stress(const int* __restrict__ cq, const double* __restrict__ xq, Material mat,
                                                const double* __restrict__ u, double* __restrict__ sigma, double* __restrict__ vm_out, int ne,
                                                double* __restrict__ bmax, int* __restrict__ barg) {
    const int NQ = NPC == 4 ? 4 : 8;
    __shared__ double smax[128];
    __shared__ int sarg[128];
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    double vm = -1.0;
    if (e < ne) {
        double lam, mu; material_at(mat, e, lam, mu);
        double avg[6] = {0, 0, 0, 0, 0, 0};
        if (NPC == 4) {
            int q[4]; double X[4][3], g[4][3];
            tet_load(cq, xq, e, q, X);
            tet_grads(X, g);
            double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int b = 0; b < 4; b++) {
                double ub[3]; load3(u, q[b], ub);
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int i = 0; i < 3; i++) H[c][i] += ub[c] * g[b][i];
            }
            double S[3][3]; hooke_from_grad(H, lam, mu, S);
            double s6[6] = {S[0][0], S[1][1], S[2][2], S[0][1], S[1][2], S[0][2]};
            for (int gp = 0; gp < NQ; gp++) {
#pragma unroll
                for (int k = 0; k < 6; k++) { if (sigma) sigma[((size_t)e * NQ + gp) * 6 + k] = s6[k]; avg[k] += s6[k]; }
            }
        } else {
            int q[8]; double X[8][3], ue[8][3];
            hex_load(cq, xq, e, q, X);
#pragma unroll
            for (int b = 0; b < 8; b++) load3(u, q[b], ue[b]);
            for (int gp = 0; gp < 8; gp++) {
                double g[8][3], N[8];
                hex_grads_at(X, gp, g, N);
                double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
                for (int b = 0; b < 8; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++)
#pragma unroll
                        for (int i = 0; i < 3; i++) H[c][i] += ue[b][c] * g[b][i];
                double S[3][3]; hooke_from_grad(H, lam, mu, S);
                double s6[6] = {S[0][0], S[1][1], S[2][2], S[0][1], S[1][2], S[0][2]};
#pragma unroll
                for (int k = 0; k < 6; k++) { if (sigma) sigma[((size_t)e * NQ + gp) * 6 + k] = s6[k]; avg[k] += s6[k]; }
            }
        }
#pragma unroll
        for (int k = 0; k < 6; k++) avg[k] /= (double)NQ;
        // sqrt(3/2 dev(σ)⊡dev(σ))
        double mean = (avg[0] + avg[1] + avg[2]) / 3.0;
        double d0 = avg[0] - mean, d1 = avg[1] - mean, d2 = avg[2] - mean;
        double dd = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * (avg[3] * avg[3] + avg[4] * avg[4] + avg[5] * avg[5]);
        vm = sqrt(1.5 * dd);
        if (vm_out) vm_out[e] = vm;
    }
    // block arg-max, first maximum wins (strict `>` in the reference loop, :492)
    smax[threadIdx.x] = vm; sarg[threadIdx.x] = e;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double v2 = smax[threadIdx.x + o]; int a2 = sarg[threadIdx.x + o];
            if (v2 > smax[threadIdx.x] || (v2 == smax[threadIdx.x] && a2 < sarg[threadIdx.x])) { smax[threadIdx.x] = v2; sarg[threadIdx.x] = a2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { bmax[blockIdx.x] = smax[0]; barg[blockIdx.x] = sarg[0]; }
}

__global__ void k_argmax_final(const double* __restrict__ bmax, const int* __restrict__ barg, int n, double* omax, int* oarg) {
    __shared__ double smax[1024];
    __shared__ int sarg[1024];
    double v = -1.0; int a = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v2 = bmax[i]; int a2 = barg[i];
        if (v2 > v || (v2 == v && a2 < a)) { v = v2; a = a2; }
    }
    smax[threadIdx.x] = v; sarg[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double v2 = smax[threadIdx.x + o]; int a2 = sarg[threadIdx.x + o];
            if (v2 > smax[threadIdx.x] || (v2 == smax[threadIdx.x] && a2 < sarg[threadIdx.x])) { smax[threadIdx.x] = v2; sarg[threadIdx.x] = a2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *omax = smax[0]; *oarg = sarg[0]; }
}

int dist_sum_per_element_w(toe_ctx* ctx, const double* local_dev, double* global_host, int width);
int dist_argmax(toe_ctx* ctx, double* max_inout, i64* cell_inout);

int stresses_with(toe_ctx* ctx, const Material& mat, const double* u_dev, double* sigma_host, double* vm_host, double* max_vm, int64_t* max_cell);

int stresses(toe_ctx* ctx, double* sigma_host, double* vm_host, double* max_vm, int64_t* max_cell) {
    if (!ctx->have_dofs || ctx->mat.mode == MAT_NONE) return toe_fail(ctx, TOE_ERR_STATE, "stresses: mesh/material not set");
    TRY(ensure_vectors(ctx));
    return stresses_with(ctx, ctx->mat, ctx->u.p, sigma_host, vm_host, max_vm, max_cell);
}

// stress recovery for ANY displacement vector and material (calculate_stresses(u, dh, cv, λ, μ) is a free function in the reference:
// FiniteElementAnalysis.jl:440 / :730); reads the mesh only — K, constraints, material and solution of the ctx stay as they are
int stresses_with(toe_ctx* ctx, const Material& mat, const double* u_dev, double* sigma_host, double* vm_host, double* max_vm, int64_t* max_cell) {
    if (!ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "stresses: DOFs not built");
    int ne = (int)ctx->ne;
    int nqp = ctx->npc == 4 ? 4 : 8;
    unsigned grid = div_up(ne, 128);
    DevBuf<double> sig, vm, bmax; DevBuf<int> barg;
    if (sigma_host) CU(sig.alloc((size_t)ne * nqp * 6));
    if (vm_host) CU(vm.alloc(ne));
    CU(bmax.alloc(grid + 1)); CU(barg.alloc(grid + 1));
    double* sp = sigma_host ? sig.p : nullptr; double* vp = vm_host ? vm.p : nullptr;
    if (ctx->npc == 4) LAUNCH(ctx, k_stress<4>, grid, 128, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, mat, u_dev, sp, vp, ne, bmax.p, barg.p);
    else               LAUNCH(ctx, k_stress<8>, grid, 128, 0, (const int*)ctx->cq.p, (const double*)ctx->xq.p, mat, u_dev, sp, vp, ne, bmax.p, barg.p);
    LAUNCH(ctx, k_argmax_final, 1, 1024, 0, (const double*)bmax.p, (const int*)barg.p, (int)grid, bmax.p + grid, barg.p + grid);
    double hm = 0; int ha = 0;
    CU(cudaMemcpyAsync(&hm, bmax.p + grid, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&ha, barg.p + grid, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->dist) {
        if (sigma_host) TRY(dist_sum_per_element_w(ctx, sig.p, sigma_host, nqp * 6));
        if (vm_host) TRY(dist_sum_per_element_w(ctx, vm.p, vm_host, 1));
    } else {
        if (sigma_host) CU(cudaMemcpyAsync(sigma_host, sig.p, (size_t)ne * nqp * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (vm_host) CU(cudaMemcpyAsync(vm_host, vm.p, (size_t)ne * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (ctx->dist) { i64 cell = ha; TRY(dist_argmax(ctx, &hm, &cell)); ha = (int)cell; }
    // reference semantics: max starts at 0.0 and cell id 0, updated only on strict > (:446-447, :492)
    if (hm > 0.0) { if (max_vm) *max_vm = hm; if (max_cell) *max_cell = (int64_t)ha + 1; }
    else { if (max_vm) *max_vm = 0.0; if (max_cell) *max_cell = 0; }
    return TOE_OK;
}
