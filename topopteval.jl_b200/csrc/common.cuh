// common.cuh — context, device buffers, error plumbing and reduction helpers of libtopopt_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>

#include "../../include/topopt_b200.h"

typedef long long i64;
typedef unsigned long long u64;

#define TOE_OK 0
#define TOE_ERR_CUDA     -1
#define TOE_ERR_ARG      -2
#define TOE_ERR_STATE    -3
#define TOE_ERR_MESH     -4
#define TOE_ERR_COMM     -5
#define TOE_ERR_NUMERIC  -6

static const int N_SM = 148;   // B200: 2 dies x 74 SMs

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    cudaError_t resize_exact(size_t count) { release(); return alloc(count); }
    size_t bytes() const { return n * sizeof(T); }
};

// Temporary device buffer of ONE API call: stream-ordered allocation from the device's default memory pool (release threshold set to
// "never" in toe_create), so that after the first call of a kind no allocation reaches the driver and — unlike cudaFree — freeing does
// not synchronise the device.  Everything that touches the buffer must be enqueued on the same stream.
template <typename T>
struct TmpBuf {
    T* p = nullptr;
    cudaStream_t s;
    explicit TmpBuf(cudaStream_t stream) : s(stream) {}
    TmpBuf(const TmpBuf&) = delete;
    TmpBuf& operator=(const TmpBuf&) = delete;
    ~TmpBuf() { if (p) cudaFreeAsync(p, s); }
    cudaError_t alloc(size_t count) {
        if (p) { cudaFreeAsync(p, s); p = nullptr; }
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocAsync((void**)&p, count * sizeof(T), s);
        if (e != cudaSuccess) p = nullptr;
        return e;
    }
};

enum MaterialMode { MAT_NONE = 0, MAT_UNIFORM = 1, MAT_SIMP = 2, MAT_PERCELL = 3 };

// device-visible material description (passed by value to kernels)
struct Material {
    int mode;
    double lambda, mu;            // MAT_UNIFORM
    double E0, nu, Emin, p;       // MAT_SIMP
    const double* density;        // MAT_SIMP (ne)
    const double* lam_e;          // MAT_PERCELL
    const double* mu_e;
};

// scalars of the PCG recurrence, resident on the device (no host round trip per iteration)
struct CGScalars {
    double gamma;      // r'z
    double pAp;
    double beta;
    double eps;        // atol + rtol*sqrt(gamma0)
    double res0;
    double aux;        // scratch for extra reductions
    i64 iter;
    i64 itmax;
    int done;
    int converged;
    int breakdown;
    int pad;
    // single-reduction (Chronopoulos–Gear) recurrence used on partitioned runs: {γ, δ, ν = r'r (l2 criterion only), unused}
    // double-buffered by iteration parity; the first 2 (3 with the l2 criterion) are summed over the ranks by the exchange
    double gd[2][4];
    double alpha[2];
    double res;        // residual in the norm of the stopping test at the last completed iteration (√(r'Mr) or ‖r‖₂)
};

struct DistState;   // dist.cu
struct TwoLevel;    // twolevel.cu

struct toe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // CUDA errors found in the runtime's last-error slot BEFORE a launch: left behind by an earlier call nobody checked (an event record,
    // a free, a library probing peer access …).  They are not this launch's, so they do not fail it — but they are not dropped either.
    i64 stale_cuda_errors = 0;
    std::string stale_err;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    toe_timings tm = {};
    i64 launches = 0;

    // mesh as given (global ids), sizes
    i64 nn = 0, ne = 0;
    int npc = 0;
    bool have_mesh = false, have_dofs = false, have_pattern = false, have_contrib = false;
    bool have_K = false, have_solution = false;

    DevBuf<int> conn0;        // ne*npc, 0-based node ids (as given)
    DevBuf<double> xyz;       // 3*nn as given
    DevBuf<int> node_q;       // nn: dof-node id (first-touch rank) or -1
    int nq = 0;               // number of referenced nodes; ndofs = 3*nq
    DevBuf<int> cq;           // ne*npc connectivity in dof-node ids
    DevBuf<double> xq;        // 3*nq coordinates in dof-node order
    // node -> element incidence (entries e*npc+a, ascending)
    DevBuf<int> inc_ptr, inc;
    // block pattern (dof-node adjacency, sorted)
    DevBuf<int> blk_ptr, blk_col, diag_slot;
    i64 nnzb = 0;
    int max_deg = 0;          // largest number of blocks in a row
    i64 ldv = 0;              // stride between the 9 value planes (>= nnzb, multiple of 16)
    // block -> contributing (e,a,b) lists, off-diagonal blocks only (entries e*64 + a*8 + b, ascending e)
    DevBuf<int> ctr_ptr, ctr;
    // the same lists relative to the row (Tet4, ROWS assembly): (position of the cell in the row node's incidence list) << 4 | a << 2 | b
    DevBuf<unsigned short> rctr;
    bool have_rctr = false;
    int max_inc = 0;          // largest number of cells around a node
    // tiles of consecutive cells for the matrix-free operator (ebe_tile.cu)
    bool have_tiles = false;
    int ntiles = 0, tile_elems = 0, tile_max_nodes = 0;
    i64 tile_slots = 0;                                   // Σ local nodes over tiles = rows of the staging array
    DevBuf<int> tile_off, tile_nodes;                     // [ntiles+1] (each tile's rows padded to a multiple of 8), [tile_slots]: dof-node id of each local node, -1 = padding
    DevBuf<int> tile_m;                                   // [ntiles]: local nodes of each tile
    DevBuf<unsigned short> tile_lconn, tile_inc, tile_nstart;   // local connectivity, node-sorted ref ids, first ref of each local node
    DevBuf<int> nst_ptr, nst;                             // node -> staging rows (ascending)
    DevBuf<double> tile_stage;                            // 3 doubles per staging row
    // K values: 9 planes of nnzb doubles, plane k = 3*c+d holds K[3q+c, 3q'+d] of block slot s at val[k*nnzb+s]
    DevBuf<double> val;

    // boundary-node selection (surface.cu): 1 = node lies on a face that belongs to exactly one cell
    DevBuf<int> surf_flag;
    bool have_surface = false;

    // material
    Material mat = {};
    DevBuf<double> density, lam_e, mu_e;
    DevBuf<double> lamw, muw;          // per-cell Lamé parameters of the current SIMP assembly (E(ρ) evaluated once per cell)

    // dof vectors
    DevBuf<double> f, u, r, p, Ap, Minv, diag, tmp;
    DevBuf<double> cg_s, cg_z;    // extra vectors of the single-reduction CG (partitioned runs only)
    DevBuf<unsigned char> dflag;  // 1 = prescribed
    DevBuf<double> dval;          // m of the handler that prescribed the dof (diag entry of the constrained K)
    bool have_diag = false;       // diag holds diag of the current operator
    bool any_dirichlet = false;

    // reduction scratch
    DevBuf<double> partials;
    DevBuf<unsigned int> counters;
    DevBuf<CGScalars> cgs;
    DevBuf<double> hist;
    DevBuf<double> cg_trace;      // diagnostic (TOE_CG_TRACE=1, partitioned CG): per iteration {γ, δ summed over ranks; this rank's γ, δ partials}
    DevBuf<int> errflag;

    // CUDA graph of a batch of PCG iterations
    cudaGraphExec_t graph_exec = nullptr;
    i64 graph_key = -1;
    i64 op_generation = 0;    // bumped whenever the operator (mesh, material, K, constraints) changes
    CGScalars* cgs_host = nullptr;   // pinned readback buffer

    // cudaFuncSetAttribute is per device: the opt-in to large dynamic shared memory is remembered per ctx, not per process
    bool spmv_attr_set = false, rows_attr_set = false;
    size_t ebe_attr_smem[2] = {0, 0}, pipe_attr_smem[2] = {0, 0};

    DistState* dist = nullptr;
    TwoLevel* tl = nullptr;       // two-level preconditioner state (twolevel.cu), built on first use
    // views set by the multi-GPU layer (null / 0 on a single GPU)
    const unsigned char* owned = nullptr;   // per local dof-node: 1 if this rank owns it (masked reductions, Dirichlet diagonal)
    const int* glob2loc = nullptr;          // global dof-node id -> local id, -1 if not on this rank
    i64 n_global = 0;                       // global number of DOFs (0 = same as local)
};

inline int toe_fail(toe_ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    return code;
}

// a reported error is also cleared from the runtime's last-error slot (sticky ones stay), so that it is not blamed on the next launch
#define CU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cudaGetLastError(); \
    return toe_fail(ctx, TOE_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); } } while (0)
#define TRY(call) do { int _s = (call); if (_s != TOE_OK) return _s; } while (0)

inline void toe_note_stale(toe_ctx* c, cudaError_t e, const char* before) {
    c->stale_cuda_errors++;
    char buf[512];
    snprintf(buf, sizeof buf, "%s (%s) was pending before the launch of %s", cudaGetErrorName(e), cudaGetErrorString(e), before);
    c->stale_err = buf;
    static const bool verbose = getenv("TOE_VERBOSE") != nullptr;
    if (verbose) fprintf(stderr, "libtopopt_b200: stale CUDA error: %s\n", buf);
}

#ifndef TOE_EMU
// a launch that the runtime refuses (bad configuration, shared-memory opt-in missing, sticky error) fails HERE, by kernel name, not
// at some later synchronisation; cudaPeekAtLastError is a host-side read (also legal during stream capture)
// (the slot is cleared first; what was in it is counted and kept: toe_debug_stale_cuda_errors)
#define LAUNCH(ctx, kern, grid, block, smem, ...) do { \
    { cudaError_t _pe = cudaGetLastError(); if (_pe != cudaSuccess) toe_note_stale((ctx), _pe, #kern); } \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); (ctx)->launches++; \
    cudaError_t _le = cudaPeekAtLastError(); \
    if (_le != cudaSuccess) return toe_fail((ctx), TOE_ERR_CUDA, "launch of %s <<<%u, %u, %zu B>>> failed: %s (%s)", #kern, (unsigned)(grid), (unsigned)(block), \
                                            (size_t)(smem), cudaGetErrorName(_le), cudaGetErrorString(_le)); } while (0)
// dynamic shared memory of the running block
#define TOE_DYN_SMEM(type, name, align) extern __shared__ __align__(align) type name[]
typedef unsigned smem_ptr_t;       // 32-bit shared-window address, as the mbarrier / bulk-copy PTX wants it
#else
// tests/cuda_emu (host-side logic check of these very sources, test infrastructure only): a launch runs the grid on fibers
#define LAUNCH(ctx, kern, grid, block, smem, ...) do { \
    auto _args = std::make_tuple(__VA_ARGS__); \
    emu::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [_args]() { std::apply([](auto... a) { kern(a...); }, _args); }, \
                reinterpret_cast<const void*>(&kern)); \
    (ctx)->launches++; \
    cudaError_t _le = cudaPeekAtLastError(); \
    if (_le != cudaSuccess) return toe_fail((ctx), TOE_ERR_CUDA, "launch of %s <<<%u, %u, %zu B>>> failed: %s (%s)", #kern, (unsigned)(grid), (unsigned)(block), \
                                            (size_t)(smem), cudaGetErrorName(_le), cudaGetErrorString(_le)); } while (0)
#define TOE_DYN_SMEM(type, name, align) type* name = reinterpret_cast<type*>(emu::dyn_smem())
typedef size_t smem_ptr_t;
#endif

static inline unsigned int div_up(i64 a, i64 b) { return (unsigned int)((a + b - 1) / b); }
static inline unsigned int min_u(unsigned int a, unsigned int b) { return a < b ? a : b; }

// ---- timing helper -------------------------------------------------------------------------------------
struct StageTimer {
    toe_ctx* c; double* slot;
    StageTimer(toe_ctx* ctx, double* s) : c(ctx), slot(s) { cudaEventRecord(c->ev0, c->stream); }
    int finish() {
        cudaEventRecord(c->ev1, c->stream);
        cudaError_t e = cudaEventSynchronize(c->ev1);
        if (e != cudaSuccess) return toe_fail(c, TOE_ERR_CUDA, "CUDA error %s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
        float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        *slot = ms * 1e-3;
        e = cudaGetLastError();
        if (e != cudaSuccess) return toe_fail(c, TOE_ERR_CUDA, "CUDA error %s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
        return TOE_OK;
    }
};

// a pair of events that is destroyed on every exit path (error returns included)
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    EventPair() {}
    EventPair(const EventPair&) = delete;
    EventPair& operator=(const EventPair&) = delete;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
};

// ---- device reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the block; result valid in thread 0.  `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    if (w == 0) {
        v = lane < nw ? sh[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// Deterministic grid reduction: every block deposits its partial, the last block to arrive (ticket counter)
// sums the partials in a fixed order.  Returns true in thread 0 of that last block, with *total set.
// `partials` must hold gridDim.x doubles.
__device__ __forceinline__ bool grid_sum_last_block(double block_val /*thread 0*/, double* partials, unsigned int* counter,
                                                    double* sh, double* total) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_val;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double s = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(&partials[i]);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) { *total = s; *counter = 0u; return true; }
    return false;
}

// two sums with one ticket: the same (last) block holds both totals.  `partials` must hold 2*gridDim.x doubles.
__device__ __forceinline__ bool grid_sum2_last_block(double v0, double v1 /*thread 0*/, double* partials, unsigned int* counter,
                                                     double* sh, double* tot0, double* tot1) {
    __shared__ bool is_last2;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v0; partials[gridDim.x + blockIdx.x] = v1;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        is_last2 = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last2) return false;
    __threadfence();
    double s0 = 0.0, s1 = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) { s0 += __ldcg(&partials[i]); s1 += __ldcg(&partials[gridDim.x + i]); }
    s0 = block_sum(s0, sh);
    s1 = block_sum(s1, sh);
    if (threadIdx.x == 0) { *tot0 = s0; *tot1 = s1; *counter = 0u; return true; }
    return false;
}

// ---- PCG scalar recurrences (device side) -------------------------------------------------------------------
__device__ __forceinline__ void cg_after_pAp(CGScalars* s, double pAp) {
    s->pAp = pAp;
    if (!(pAp > 0.0)) { s->done = 1; s->breakdown = 1; }       // Krylov.jl stops on non-positive curvature
}
// gnew = r'Mr of the new residual; res2 = square of the residual norm the stopping test uses (= gnew for Krylov.jl's M-norm rule,
// r'r for the plain l2 rule)
__device__ __forceinline__ void cg_after_gamma(CGScalars* s, double gnew, double* hist, i64 hist_cap, double res2 = -1.0) {
    s->beta = gnew / s->gamma;
    s->gamma = gnew;
    s->iter += 1;
    double res = sqrt(res2 >= 0.0 ? res2 : gnew);
    s->res = res;
    if (s->iter < hist_cap) hist[s->iter] = res;
    if (res <= s->eps) { s->done = 1; s->converged = 1; }
    else if (s->iter >= s->itmax) s->done = 1;
}

// (cell, a, b) packed into one int for the block -> contribution lists
template <int NPC> __host__ __device__ __forceinline__ int ctr_pack(int e, int a, int b) {
    return NPC == 4 ? ((e << 4) | (a << 2) | b) : ((e << 6) | (a << 3) | b);
}
template <int NPC> __host__ __device__ __forceinline__ void ctr_unpack(int v, int& e, int& a, int& b) {
    if (NPC == 4) { e = v >> 4; a = (v >> 2) & 3; b = v & 3; } else { e = v >> 6; a = (v >> 3) & 7; b = v & 7; }
}

// ---- prototypes across translation units -----------------------------------------------------------------
int scan_exclusive_i32(toe_ctx* ctx, const int* in, int* out, i64 n, i64* total_out);   // out may alias in; out has n+1 entries
int mesh_upload(toe_ctx* ctx, i64 nn, const double* xyz, i64 ne, int npc, const int64_t* conn);
int mesh_build_dofs(toe_ctx* ctx);
int mesh_build_pattern(toe_ctx* ctx);
int mesh_build_contrib(toe_ctx* ctx);
int ensure_vectors(toe_ctx* ctx);
int compute_diag(toe_ctx* ctx);   // diag of the current operator (assembled K or EbE + dirichlet overrides)
int op_apply(toe_ctx* ctx, const double* x, double* y, int matrix_free, double* dot_out_dev /*or null*/, bool assume_masked);
double op_bytes(toe_ctx* ctx, int matrix_free);
int dist_post_spmv(toe_ctx* ctx, double* y);      // interface sum (no-op without dist)
int dist_allreduce(toe_ctx* ctx, double* dev_vals, int count);
int dist_align(toe_ctx* ctx);                     // ranks' streams meet (no-op without dist); see dist.cu
static const int PARTIALS_ALIGN_SLOT = 3 * (N_SM * 8) + 16;   // scratch double of dist_align inside ctx->partials (after the k_norms outputs)
int mesh_build_tiles(toe_ctx* ctx);
int ebe_tile_launch(toe_ctx* ctx, const double* x, double* y, CGScalars* cg, bool mask, const int* done_flag, double* dot_out);
int dist_sum_per_element(toe_ctx* ctx, const double* local_dev, double* global_host);   // per-cell output, global cell order
void dist_destroy(toe_ctx* ctx);
// two-level preconditioner (twolevel.cu)
int tl_prepare(toe_ctx* ctx, int matrix_free, int* coarse_dofs, double* setup_seconds);
int tl_cg_init(toe_ctx* ctx, double atol, double rtol, i64 itmax, i64 hist_cap);
int tl_cg_after_operator(toe_ctx* ctx, i64 hist_cap);
void tl_invalidate(toe_ctx* ctx);
void tl_destroy(toe_ctx* ctx);
