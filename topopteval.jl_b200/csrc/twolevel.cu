// twolevel.cu — two-level preconditioner for the CG solve ("stronger preconditioning behind
// SolverConfig.preconditioner"; the reference's hooks are create_preconditioner, RobustSolver.jl:223-271).
//
//     M⁻¹ = D⁻¹ + Z (ZᵀKZ)⁻¹ Zᵀ            additive: Jacobi + a coarse-space correction
//
// Z spans the six rigid-body modes (3 translations, 3 rotations about the box centre) of each box of a bx×by×bz grid laid
// over the mesh's bounding box: 6·bx·by·bz coarse unknowns (≤ 6144).  Jacobi-CG on a slender solid needs O(L/h) iterations
// because it cannot move the low-energy bending / stretching modes; the coarse space carries exactly those.  Measured with
// the CPU oracle on the synthetic cantilever (96×32×12 cubes, 125 k DOFs): 3124 Jacobi iterations → 72 with 512 boxes
// (3460 → 85 with SIMP-like densities), same solution to 2e-10.
//
//   setup (per mesh)       box of every dof-node, nodes listed per box in ascending order (stable counting sort)
//   coarse operator        A_c = ZᵀKZ by coloured probing: boxes two apart never share a cell, so all boxes of one of the ≤27
//   (per assembled K)      colours are probed with ONE operator application per mode: ≤162 SpMVs + a per-box reduction, no atomics
//                          → works for the assembled and the matrix-free operator alike; then in-place Gauss–Jordan → A_c⁻¹
//   apply (per iteration)  w = Zᵀr (one CTA per box, fixed-order sums) → y = A_c⁻¹w (dense GEMV, ≤302 MB) → z = D⁻¹r + Zy
// Everything is deterministic (fixed summation orders), prescribed DOFs are masked out of Z.
#include "common.cuh"
#include <cstdlib>
#include <cmath>

struct TwoLevel {
    bool have_setup = false;
    i64 built_generation = -1;          // op_generation the inverse was built for
    int built_matrix_free = -1;
    int b[3] = {1, 1, 1};
    int m = 0, nc = 0;
    double lo[3] = {0, 0, 0}, inv_h[3] = {0, 0, 0}, h[3] = {0, 0, 0};
    DevBuf<int> agg, agg_ptr, agg_nodes;
    DevBuf<double> A, w, y, z, tv, ty;
    double setup_seconds = 0.0;
};

void tl_destroy(toe_ctx* ctx) { delete ctx->tl; ctx->tl = nullptr; }

struct TLGeom { double lo[3], inv_h[3], h[3]; int b[3]; };

__device__ __forceinline__ int tl_box_of(const TLGeom& g, const double* x, int ib[3]) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int i = (int)floor((x[c] - g.lo[c]) * g.inv_h[c]);
        ib[c] = i < 0 ? 0 : (i >= g.b[c] ? g.b[c] - 1 : i);
    }
    return ib[0] + g.b[0] * (ib[1] + g.b[1] * ib[2]);
}
__device__ __forceinline__ void tl_centre(const TLGeom& g, int box, double c[3]) {
    int ix = box % g.b[0], iy = (box / g.b[0]) % g.b[1], iz = box / (g.b[0] * g.b[1]);
    c[0] = g.lo[0] + (ix + 0.5) * g.h[0]; c[1] = g.lo[1] + (iy + 0.5) * g.h[1]; c[2] = g.lo[2] + (iz + 0.5) * g.h[2];
}

__global__ void k_tl_agg(const double* __restrict__ xq, int nq, TLGeom g, int* __restrict__ agg) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double x[3] = {xq[3 * (size_t)q], xq[3 * (size_t)q + 1], xq[3 * (size_t)q + 2]};
    int ib[3];
    agg[q] = tl_box_of(g, x, ib);
}

// cells never span more than two neighbouring boxes per axis? (the probing colours rely on it)
__global__ void k_tl_check_adjacent(const int* __restrict__ blk_ptr, const int* __restrict__ blk_col, const int* __restrict__ agg, int nq, TLGeom g, int* bad) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int a = agg[q];
    const int ax = a % g.b[0], ay = (a / g.b[0]) % g.b[1], az = a / (g.b[0] * g.b[1]);
    for (int s = blk_ptr[q]; s < blk_ptr[q + 1]; s++) {
        const int c = agg[blk_col[s]];
        const int cx = c % g.b[0], cy = (c / g.b[0]) % g.b[1], cz = c / (g.b[0] * g.b[1]);
        if (abs(cx - ax) > 1 || abs(cy - ay) > 1 || abs(cz - az) > 1) { atomicExch(bad, 1); return; }
    }
}

// stable counting sort of the dof-nodes by box: chunk = 256 consecutive nodes
static const int TL_CHUNK = 256;
// partitioned runs: only the nodes this rank OWNS are listed, so that the per-box sums of all ranks add up to the global sums
__global__ void __launch_bounds__(TL_CHUNK) k_tl_count(const int* __restrict__ agg, const unsigned char* __restrict__ owned, int nq, int nchunks,
                                                       int* __restrict__ cnt /* [m][nchunks] */) {
    int q = blockIdx.x * TL_CHUNK + threadIdx.x;
    if (q < nq && (!owned || owned[q])) atomicAdd(&cnt[(size_t)agg[q] * nchunks + blockIdx.x], 1);
}
__global__ void __launch_bounds__(TL_CHUNK) k_tl_fill(const int* __restrict__ agg, const unsigned char* __restrict__ owned, int nq, int nchunks,
                                                      const int* __restrict__ start /* scanned cnt */, int* __restrict__ agg_nodes) {
    __shared__ int sa[TL_CHUNK];
    const int q = blockIdx.x * TL_CHUNK + threadIdx.x;
    const int a = (q < nq && (!owned || owned[q])) ? agg[q] : -1;
    sa[threadIdx.x] = a;
    __syncthreads();
    if (a < 0) return;
    int rank = 0;
    for (int t = 0; t < (int)threadIdx.x; t++) rank += (sa[t] == a);
    agg_nodes[start[(size_t)a * nchunks + blockIdx.x] + rank] = q;
}
__global__ void k_tl_box_ptr(const int* __restrict__ start, int nchunks, int m, int total, int* __restrict__ agg_ptr) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < m) agg_ptr[a] = start[(size_t)a * nchunks];
    if (a == m) agg_ptr[m] = total;
}

// w[6I+a] = Σ_{i in box I} P_i[:,a] · r_i      (P_i = [I | -[d_i]x], d_i = x_i - centre(I)); one CTA per box, fixed order
static const int TL_RT = 256;
__global__ void __launch_bounds__(TL_RT) k_tl_restrict(const int* __restrict__ agg_ptr, const int* __restrict__ agg_nodes, const double* __restrict__ xq,
                                                       const unsigned char* __restrict__ dflag, const double* __restrict__ r, TLGeom g,
                                                       double* __restrict__ w, const int* done_flag) {
    __shared__ double red[32];
    if (done_flag && *done_flag) return;
    const int I = blockIdx.x;
    double c[3]; tl_centre(g, I, c);
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int k = agg_ptr[I] + threadIdx.x; k < agg_ptr[I + 1]; k += TL_RT) {
        const int q = agg_nodes[k];
        double rv[3], d[3];
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
            const size_t dof = 3 * (size_t)q + cc;
            rv[cc] = dflag[dof] ? 0.0 : r[dof];
            d[cc] = xq[dof] - c[cc];
        }
        s[0] += rv[0]; s[1] += rv[1]; s[2] += rv[2];
        s[3] += d[1] * rv[2] - d[2] * rv[1];          // (d × r): the rotation modes u = ω × d pair with r as ω · (d × r)
        s[4] += d[2] * rv[0] - d[0] * rv[2];
        s[5] += d[0] * rv[1] - d[1] * rv[0];
    }
#pragma unroll
    for (int a = 0; a < 6; a++) {
        double v = block_sum(s[a], red);
        if (threadIdx.x == 0) w[6 * (size_t)I + a] = v;
    }
}

// y = A w (dense, row-major), one warp per row
__global__ void __launch_bounds__(256) k_tl_gemv(const double* __restrict__ A, const double* __restrict__ w, double* __restrict__ y, int nc, const int* done_flag) {
    if (done_flag && *done_flag) return;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nc) return;                      // whole warps leave together
    const double* a = A + (size_t)row * nc;
    double s = 0.0;
    for (int c = lane; c < nc; c += 32) s += a[c] * w[c];
    s = warp_sum(s);
    if (lane == 0) y[row] = s;
}

// z = D⁻¹ r + Z y, γ' = r·z; the last block closes the CG iteration (β, convergence) or, with init != 0, opens the solve
__global__ void __launch_bounds__(256) k_tl_z(const double* __restrict__ Minv, const double* __restrict__ r, const int* __restrict__ agg,
                                              const double* __restrict__ xq, const double* __restrict__ y, const unsigned char* __restrict__ dflag, TLGeom g,
                                              double* __restrict__ z, int nq, CGScalars* cg, int init, double atol, double rtol, i64 itmax,
                                              double* hist, i64 hist_cap, double* partials, unsigned int* counter,
                                              const unsigned char* __restrict__ owned, double* local_out) {
    __shared__ double red[32];
    if (!init && cg->done) return;
    double s = 0.0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const int I = agg[q];
        double c[3]; tl_centre(g, I, c);
        const double* yy = y + 6 * (size_t)I;
        const double d0 = xq[3 * (size_t)q] - c[0], d1 = xq[3 * (size_t)q + 1] - c[1], d2 = xq[3 * (size_t)q + 2] - c[2];
        const double u[3] = {yy[0] + yy[4] * d2 - yy[5] * d1, yy[1] + yy[5] * d0 - yy[3] * d2, yy[2] + yy[3] * d1 - yy[4] * d0};   // t + ω × d
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
            const size_t dof = 3 * (size_t)q + cc;
            const double ri = r[dof];
            const double zi = Minv[dof] * ri + (dflag[dof] ? 0.0 : u[cc]);
            z[dof] = zi;
            if (!owned || owned[q]) s += ri * zi;
        }
    }
    s = block_sum(s, red);
    double tot;
    if (grid_sum_last_block(s, partials, counter, red, &tot)) {
        if (local_out) { *local_out = tot; return; }          // partitioned: the caller allreduces and closes (k_tl_close)
        if (init) {
            cg->gamma = tot; cg->pAp = 0.0; cg->beta = 0.0;
            cg->res0 = sqrt(tot > 0.0 ? tot : 0.0);
            cg->eps = atol + rtol * cg->res0;
            cg->iter = 0; cg->itmax = itmax;
            cg->converged = (cg->res0 <= cg->eps) ? 1 : 0;
            cg->done = (cg->converged || itmax <= 0) ? 1 : 0;
            cg->breakdown = 0;
            if (hist_cap > 0) hist[0] = cg->res0;
        } else {
            cg_after_gamma(cg, tot, hist, hist_cap);
        }
    }
}

// x = 0, r = f, Minv = 1 ./ D with D[abs(D) < 1e-12] = 1 (RobustSolver.jl:231-236)
__global__ void __launch_bounds__(256) k_tl_init(const double* __restrict__ f, const double* __restrict__ diag, double* __restrict__ Minv,
                                                 double* __restrict__ x, double* __restrict__ r, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double d = diag[i];
        if (fabs(d) < 1e-12) d = 1.0;
        Minv[i] = 1.0 / d; x[i] = 0.0; r[i] = f[i];
    }
}
// α = γ/p'Ap, x += α p, r -= α Ap
__global__ void __launch_bounds__(256) k_tl_xr(const double* __restrict__ p, const double* __restrict__ Ap, double* __restrict__ x, double* __restrict__ r,
                                               size_t n, const CGScalars* cg) {
    if (cg->done) return;
    const double alpha = cg->gamma / cg->pAp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        r[i] -= alpha * Ap[i];
    }
}
// p = z + β p   (first: p = z)
__global__ void __launch_bounds__(256) k_tl_p(const double* __restrict__ z, double* __restrict__ p, size_t n, const CGScalars* cg, int first) {
    if (!first && cg->done) return;
    const double beta = first ? 0.0 : cg->beta;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = first ? z[i] : z[i] + beta * p[i];
}

// partitioned runs: the scalars come back from an allreduce; one thread closes what the last block closes on a single GPU
__global__ void k_tl_set_pAp(CGScalars* cg, const double* pAp) {
    if (threadIdx.x || blockIdx.x || cg->done) return;
    cg_after_pAp(cg, *pAp);
}
__global__ void k_tl_close(CGScalars* cg, const double* gamma, int init, double atol, double rtol, i64 itmax, double* hist, i64 hist_cap) {
    if (threadIdx.x || blockIdx.x) return;
    if (!init && cg->done) return;
    const double tot = *gamma;
    if (init) {
        cg->gamma = tot; cg->pAp = 0.0; cg->beta = 0.0;
        cg->res0 = sqrt(tot > 0.0 ? tot : 0.0);
        cg->eps = atol + rtol * cg->res0;
        cg->iter = 0; cg->itmax = itmax;
        cg->converged = (cg->res0 <= cg->eps) ? 1 : 0;
        cg->done = (cg->converged || itmax <= 0) ? 1 : 0;
        cg->breakdown = 0;
        if (hist_cap > 0) hist[0] = cg->res0;
    } else {
        cg_after_gamma(cg, tot, hist, hist_cap);
    }
}

// probing vector of (colour, mode): v_i = P_i[:,mode] for nodes whose box has that colour, masked at prescribed dofs
__global__ void k_tl_probe(const int* __restrict__ agg, const double* __restrict__ xq, const unsigned char* __restrict__ dflag, TLGeom g,
                           int cx, int cy, int cz, int mode, double* __restrict__ v, int nq) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int I = agg[q];
    const int ix = I % g.b[0], iy = (I / g.b[0]) % g.b[1], iz = I / (g.b[0] * g.b[1]);
    double u[3] = {0, 0, 0};
    if (ix % 3 == cx && iy % 3 == cy && iz % 3 == cz) {
        if (mode < 3) u[mode] = 1.0;
        else {
            double c[3]; tl_centre(g, I, c);
            const double d0 = xq[3 * (size_t)q] - c[0], d1 = xq[3 * (size_t)q + 1] - c[1], d2 = xq[3 * (size_t)q + 2] - c[2];
            if (mode == 3) { u[1] = -d2; u[2] = d1; }            // e_x × d
            else if (mode == 4) { u[0] = d2; u[2] = -d0; }       // e_y × d
            else { u[0] = -d1; u[1] = d0; }                      // e_z × d
        }
    }
#pragma unroll
    for (int cc = 0; cc < 3; cc++) { const size_t dof = 3 * (size_t)q + cc; v[dof] = dflag[dof] ? 0.0 : u[cc]; }
}

// column (J*, mode) of A_c for every box row I: J* = the box of this colour inside I's 3x3x3 neighbourhood (if any)
__global__ void k_tl_scatter(const double* __restrict__ w, TLGeom g, int cx, int cy, int cz, int mode, double* __restrict__ A, int m, int nc) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * m) return;
    const int I = t / 6;
    const int ix = I % g.b[0], iy = (I / g.b[0]) % g.b[1], iz = I / (g.b[0] * g.b[1]);
    const int dx = (cx - ix % 3 + 3) % 3, dy = (cy - iy % 3 + 3) % 3, dz = (cz - iz % 3 + 3) % 3;
    const int jx = ix + (dx == 2 ? -1 : dx), jy = iy + (dy == 2 ? -1 : dy), jz = iz + (dz == 2 ? -1 : dz);
    if (jx < 0 || jx >= g.b[0] || jy < 0 || jy >= g.b[1] || jz < 0 || jz >= g.b[2]) return;
    const int J = jx + g.b[0] * (jy + g.b[1] * jz);
    A[(size_t)t * nc + 6 * J + mode] = w[t];
}

// A_c(I,J) = Σ_{i∈I} Σ_{j∈J} P_iᵀ K_ij P_j straight from the assembled blocks (single GPU, assembled K): one CTA per (box I,
// neighbour slot), threads stride over the nodes of I in list order, 36 partial sums each, fixed-order block reduction.
// Replaces the ≤162 probing operator applications when K is in memory; the probing path stays for the matrix-free
// operator and for partitioned runs.
__device__ __forceinline__ void tl_mode(int b, const double d[3], double u[3]) {      // displacement of rigid-body mode b at offset d
    u[0] = u[1] = u[2] = 0.0;
    if (b < 3) u[b] = 1.0;
    else if (b == 3) { u[1] = -d[2]; u[2] = d[1]; }
    else if (b == 4) { u[0] = d[2]; u[2] = -d[0]; }
    else { u[0] = -d[1]; u[1] = d[0]; }
}
__global__ void __launch_bounds__(128) k_tl_coarse_direct(const int* __restrict__ agg_ptr, const int* __restrict__ agg_nodes, const int* __restrict__ agg,
                                                          const int* __restrict__ blk_ptr, const int* __restrict__ blk_col, const double* __restrict__ val, i64 ldv,
                                                          const double* __restrict__ xq, const unsigned char* __restrict__ dflag, TLGeom g,
                                                          double* __restrict__ A, int nc) {
    __shared__ double red[32];
    const int I = blockIdx.x / 27, slot = blockIdx.x - 27 * I;
    const int ix = I % g.b[0], iy = (I / g.b[0]) % g.b[1], iz = I / (g.b[0] * g.b[1]);
    const int jx = ix + slot % 3 - 1, jy = iy + (slot / 3) % 3 - 1, jz = iz + slot / 9 - 1;
    if (jx < 0 || jx >= g.b[0] || jy < 0 || jy >= g.b[1] || jz < 0 || jz >= g.b[2]) return;     // uniform per CTA
    const int J = jx + g.b[0] * (jy + g.b[1] * jz);
    double cI[3], cJ[3]; tl_centre(g, I, cI); tl_centre(g, J, cJ);
    double acc[36];
#pragma unroll
    for (int k = 0; k < 36; k++) acc[k] = 0.0;
    for (int k = agg_ptr[I] + threadIdx.x; k < agg_ptr[I + 1]; k += blockDim.x) {
        const int i = agg_nodes[k];
        double di[3], mi[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { di[c] = xq[3 * (size_t)i + c] - cI[c]; mi[c] = dflag[3 * (size_t)i + c] ? 0.0 : 1.0; }
        double T[3][6];                                        // Σ_{j∈J} K_ij P_j  (rows of prescribed dofs of i are dropped below)
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int b = 0; b < 6; b++) T[c][b] = 0.0;
        bool any = false;
        for (int s = blk_ptr[i]; s < blk_ptr[i + 1]; s++) {
            const int j = blk_col[s];
            if (agg[j] != J) continue;
            any = true;
            double B[3][3], dj[3], mj[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                dj[c] = xq[3 * (size_t)j + c] - cJ[c]; mj[c] = dflag[3 * (size_t)j + c] ? 0.0 : 1.0;
#pragma unroll
                for (int d = 0; d < 3; d++) B[c][d] = val[(size_t)(3 * c + d) * ldv + s];
            }
#pragma unroll
            for (int b = 0; b < 6; b++) {
                double u[3]; tl_mode(b, dj, u);
#pragma unroll
                for (int c = 0; c < 3; c++) T[c][b] += B[c][0] * (mj[0] * u[0]) + B[c][1] * (mj[1] * u[1]) + B[c][2] * (mj[2] * u[2]);
            }
        }
        if (!any) continue;
#pragma unroll
        for (int a = 0; a < 6; a++) {
            double u[3]; tl_mode(a, di, u);
#pragma unroll
            for (int b = 0; b < 6; b++) acc[6 * a + b] += (mi[0] * u[0]) * T[0][b] + (mi[1] * u[1]) * T[1][b] + (mi[2] * u[2]) * T[2][b];
        }
    }
#pragma unroll
    for (int k = 0; k < 36; k++) {
        const double v = block_sum(acc[k], red);
        if (threadIdx.x == 0) A[(size_t)(6 * I + k / 6) * nc + 6 * J + k % 6] = v;
    }
}

// A := (A + Aᵀ)/2; empty modes (zero diagonal: boxes without free nodes) are decoupled with a unit diagonal
__global__ void k_tl_symmetrise(double* __restrict__ A, int nc) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nc * nc) return;
    const int i = (int)(t / nc), j = (int)(t - (size_t)i * nc);
    if (j <= i) return;
    const double v = 0.5 * (A[t] + A[(size_t)j * nc + i]);
    A[t] = v; A[(size_t)j * nc + i] = v;
}
__global__ void k_tl_diag_scale(const double* __restrict__ A, int nc, double* __restrict__ dmax_part) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nc; i += blockDim.x) { double d = fabs(A[(size_t)i * nc + i]); s = d > s ? d : s; }
    // max over the block via repeated shuffles on the (non-negative) values
    for (int o = 16; o > 0; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, s, o); s = t > s ? t : s; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double mx = 0.0; for (int k = 0; k < (int)(blockDim.x >> 5); k++) mx = red[k] > mx ? red[k] : mx; *dmax_part = mx; }
}

// in-place Gauss–Jordan inverse of the SPD matrix A (no pivoting).  Step k: row k is scaled by 1/pivot, every other row i gets
// A[i,:] -= A[i,k]·A[k,:] with the k-th column replaced by -A[i,k]/pivot.  A pivot below thr·dmax decouples that mode.
__global__ void __launch_bounds__(256) k_gj_row(double* __restrict__ A, int nc, int k, const double* __restrict__ dmax, double thr, double* __restrict__ colk, int* __restrict__ skipped) {
    __shared__ double s_p;
    __shared__ int s_skip;
    double* rowk = A + (size_t)k * nc;
    if (threadIdx.x == 0) { s_p = rowk[k]; s_skip = !(s_p > thr * (*dmax)); if (s_skip) atomicAdd(skipped, 1); }
    __syncthreads();
    const double p = s_p;
    const bool skip = s_skip != 0;
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        colk[j] = (j == k || skip) ? 0.0 : A[(size_t)j * nc + k];          // column k as it was (rows != k); a skipped mode eliminates nothing
        if (skip) { if (j != k) A[(size_t)j * nc + k] = 0.0; }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        if (skip) rowk[j] = (j == k) ? 1.0 : 0.0;
        else rowk[j] = (j == k) ? 1.0 / p : rowk[j] / p;
    }
}
__global__ void __launch_bounds__(256) k_gj_update(double* __restrict__ A, int nc, int k, const double* __restrict__ colk) {
    const int i = blockIdx.x;
    if (i == k) return;
    const double f = colk[i];
    if (f == 0.0) return;
    const double* rowk = A + (size_t)k * nc;
    double* rowi = A + (size_t)i * nc;
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        const double base = (j == k) ? 0.0 : rowi[j];
        rowi[j] = base - f * rowk[j];
    }
}

// ---- blocked Gauss–Jordan (32-wide pivot blocks): the same in-place inverse with 1/30 of the memory traffic -----------------
//   step K:  Dinv = A_KK⁻¹ (in shared memory);  R = Dinv·A_K: (row panel, R_K = Dinv);  C = A_:K (column panel, saved);
//            A_ij ← (j∈K ? 0 : A_ij) − C_i R_j for rows i∉K;  A_K: ← R
static const int GJB = 32;

// empty modes (boxes without free nodes: zero diagonal) are decoupled before the elimination starts
__global__ void k_gjb_flag_empty(const double* __restrict__ A, int nc, const double* __restrict__ dmax, double thr, int* __restrict__ skip, int* __restrict__ nskipped) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    const int e = !(A[(size_t)k * nc + k] > thr * (*dmax));
    skip[k] = e;
    if (e) atomicAdd(nskipped, 1);
}
__global__ void k_gjb_clear_empty(double* __restrict__ A, int nc, const int* __restrict__ skip) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nc * nc) return;
    const int i = (int)(t / nc), j = (int)(t - (size_t)i * nc);
    if (skip[i] || skip[j]) A[t] = (i == j) ? 1.0 : 0.0;
}

// inverse of the bs×bs diagonal block in shared memory; a pivot that has become tiny decouples its mode (flagged in skip[])
__global__ void __launch_bounds__(256) k_gjb_diag(const double* __restrict__ A, int nc, int k0, int bs, const double* __restrict__ dmax, double thr,
                                                  double* __restrict__ Dinv /* GJB×GJB */, int* __restrict__ skip, int* __restrict__ nskipped) {
    __shared__ double D[GJB][GJB + 1];
    __shared__ double cp[GJB], rp[GJB];
    __shared__ int s_skip;
    for (int t = threadIdx.x; t < GJB * GJB; t += blockDim.x) {
        const int i = t / GJB, j = t - i * GJB;
        D[i][j] = (i < bs && j < bs) ? A[(size_t)(k0 + i) * nc + k0 + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int p = 0; p < bs; p++) {
        if (threadIdx.x == 0) {
            const bool already = skip[k0 + p] != 0;
            const bool tiny = !(D[p][p] > thr * (*dmax));
            s_skip = (already || tiny) ? 1 : 0;
            if (tiny && !already) { skip[k0 + p] = 1; atomicAdd(nskipped, 1); }
        }
        __syncthreads();
        const bool sk = s_skip != 0;
        const double piv = D[p][p];
        if (threadIdx.x < GJB) {
            const int j = threadIdx.x;
            cp[j] = (j == p || sk) ? 0.0 : D[j][p];
            rp[j] = sk ? (j == p ? 1.0 : 0.0) : (j == p ? 1.0 / piv : D[p][j] / piv);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < GJB * GJB; t += blockDim.x) {
            const int i = t / GJB, j = t - i * GJB;
            if (i == p) D[i][j] = rp[j];
            else {
                const double base = (j == p) ? 0.0 : D[i][j];
                D[i][j] = base - cp[i] * rp[j];
            }
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < GJB * GJB; t += blockDim.x) { const int i = t / GJB, j = t - i * GJB; Dinv[t] = D[i][j]; }
}

// column panel C[i][q] = A[i][k0+q] (0 for decoupled modes), row panel R[p][j] = Σ_q Dinv[p][q] A[k0+q][j] (j∉K), R[p][k0+q] = Dinv[p][q]
__global__ void __launch_bounds__(256) k_gjb_panels(const double* __restrict__ A, int nc, int k0, int bs, const double* __restrict__ Dinv,
                                                    const int* __restrict__ skip, double* __restrict__ C, double* __restrict__ R) {
    __shared__ double Ds[GJB][GJB + 1];
    for (int t = threadIdx.x; t < GJB * GJB; t += blockDim.x) Ds[t / GJB][t % GJB] = Dinv[t];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;          // one column of R and one row of C per thread
    if (j >= nc) return;
    for (int q = 0; q < GJB; q++) C[(size_t)j * GJB + q] = (q < bs && !skip[k0 + q]) ? A[(size_t)j * nc + k0 + q] : 0.0;
    const bool inK = j >= k0 && j < k0 + bs;
    double a[GJB];
#pragma unroll
    for (int q = 0; q < GJB; q++) a[q] = (q < bs && !inK) ? A[(size_t)(k0 + q) * nc + j] : 0.0;
    for (int p = 0; p < GJB; p++) {
        double v = 0.0;
        if (p < bs) {
            if (inK) v = Ds[p][j - k0];
            else if (!skip[k0 + p]) {
#pragma unroll
                for (int q = 0; q < GJB; q++) v += Ds[p][q] * a[q];
            }
        }
        R[(size_t)p * nc + j] = v;
    }
}

// A_ij ← (j∈K ? 0 : A_ij) − Σ_q C[i][q] R[q][j]  for rows i∉K; 64×64 tile per CTA, 4×4 outputs per thread
__global__ void __launch_bounds__(256) k_gjb_update(double* __restrict__ A, int nc, int k0, int bs, const double* __restrict__ C, const double* __restrict__ R, int tiles) {
    __shared__ double Cs[64][GJB + 1];
    __shared__ double Rs[GJB][64 + 2];
    const int ti = blockIdx.x / tiles, tj = blockIdx.x - ti * tiles;
    const int i0 = ti * 64, j0 = tj * 64;
    for (int t = threadIdx.x; t < 64 * GJB; t += blockDim.x) {
        const int r = t / GJB, q = t - r * GJB;
        Cs[r][q] = (i0 + r < nc) ? C[(size_t)(i0 + r) * GJB + q] : 0.0;
    }
    for (int t = threadIdx.x; t < GJB * 64; t += blockDim.x) {
        const int q = t / 64, c = t - q * 64;
        Rs[q][c] = (j0 + c < nc) ? R[(size_t)q * nc + j0 + c] : 0.0;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
    for (int q = 0; q < GJB; q++) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) a[u] = Cs[ty * 4 + u][q];
#pragma unroll
        for (int v = 0; v < 4; v++) b[v] = Rs[q][tx * 4 + v];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[u][v] += a[u] * b[v];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int i = i0 + ty * 4 + u;
        if (i >= nc || (i >= k0 && i < k0 + bs)) continue;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const int j = j0 + tx * 4 + v;
            if (j >= nc) continue;
            const size_t at = (size_t)i * nc + j;
            const double base = (j >= k0 && j < k0 + bs) ? 0.0 : A[at];
            A[at] = base - acc[u][v];
        }
    }
}
__global__ void k_gjb_write_rows(double* __restrict__ A, int nc, int k0, int bs, const double* __restrict__ R) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)bs * nc) return;
    const int p = (int)(t / nc), j = (int)(t - (size_t)p * nc);
    A[(size_t)(k0 + p) * nc + j] = R[(size_t)p * nc + j];
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static TLGeom tl_geom(const TwoLevel* t) {
    TLGeom g;
    for (int c = 0; c < 3; c++) { g.lo[c] = t->lo[c]; g.inv_h[c] = t->inv_h[c]; g.h[c] = t->h[c]; g.b[c] = t->b[c]; }
    return g;
}

__global__ void k_tl_bbox(const double* __restrict__ xq, int nq, double* __restrict__ part /* 6 per block */) {
    __shared__ double sh[6][32];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x)
        for (int c = 0; c < 3; c++) { double v = xq[3 * (size_t)q + c]; lo[c] = v < lo[c] ? v : lo[c]; hi[c] = v > hi[c] ? v : hi[c]; }
    for (int c = 0; c < 3; c++)
        for (int o = 16; o > 0; o >>= 1) {
            double a = __shfl_xor_sync(0xffffffffu, lo[c], o), b = __shfl_xor_sync(0xffffffffu, hi[c], o);
            lo[c] = a < lo[c] ? a : lo[c]; hi[c] = b > hi[c] ? b : hi[c];
        }
    if ((threadIdx.x & 31) == 0) for (int c = 0; c < 3; c++) { sh[c][threadIdx.x >> 5] = lo[c]; sh[3 + c][threadIdx.x >> 5] = hi[c]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; c++) {
            double a = 1e300, b = -1e300;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) { a = sh[c][k] < a ? sh[c][k] : a; b = sh[3 + c][k] > b ? sh[3 + c][k] : b; }
            part[6 * blockIdx.x + c] = a; part[6 * blockIdx.x + 3 + c] = b;
        }
    }
}

static int tl_setup(toe_ctx* ctx) {
    if (!ctx->tl) ctx->tl = new TwoLevel();
    TwoLevel* t = ctx->tl;
    if (t->have_setup) return TOE_OK;
    if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "two-level preconditioner: call setup_problem first");
    const int nq = ctx->nq;
    // bounding box: of the referenced nodes — on a partitioned ctx of ALL nodes of the global mesh (every rank holds the global
    // coordinate array), so that every rank lays the same box grid without talking to the others
    const int nb = 64;
    DevBuf<double> part; CU(part.alloc(6 * nb));
    if (ctx->dist) LAUNCH(ctx, k_tl_bbox, nb, 256, 0, (const double*)ctx->xyz.p, (int)ctx->nn, part.p);
    else           LAUNCH(ctx, k_tl_bbox, nb, 256, 0, (const double*)ctx->xq.p, nq, part.p);
    std::vector<double> hp(6 * nb);
    CU(cudaMemcpyAsync(hp.data(), part.p, 6 * nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < nb; k++) for (int c = 0; c < 3; c++) { lo[c] = std::fmin(lo[c], hp[6 * k + c]); hi[c] = std::fmax(hi[c], hp[6 * k + 3 + c]); }
    int target = 512;
    if (const char* e = getenv("TOE_TL_BOXES_TARGET")) { int v = atoi(e); if (v >= 1) target = v; }
    if (target > 1024) target = 1024;
    int forced[3] = {0, 0, 0};
    if (const char* e = getenv("TOE_TL_BOXES")) sscanf(e, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
    CU(t->agg.alloc(nq));
    DevBuf<int> bad; CU(bad.alloc(1));
    DevBuf<double> badd; CU(badd.alloc(1));
    for (;; target /= 2) {
        int b[3] = {1, 1, 1};
        if (forced[0] > 0 && forced[1] > 0 && forced[2] > 0 && (i64)forced[0] * forced[1] * forced[2] <= 1024) { b[0] = forced[0]; b[1] = forced[1]; b[2] = forced[2]; forced[0] = 0; }
        else while ((i64)b[0] * b[1] * b[2] < target) {            // halve the longest box edge until there are enough boxes
            int ax = 0;
            for (int c = 1; c < 3; c++) if ((hi[c] - lo[c]) / b[c] > (hi[ax] - lo[ax]) / b[ax]) ax = c;
            b[ax] *= 2;
        }
        for (int c = 0; c < 3; c++) {
            double ext = hi[c] - lo[c];
            if (!(ext > 0.0)) { ext = 1.0; b[c] = 1; }
            t->b[c] = b[c]; t->lo[c] = lo[c]; t->h[c] = ext / b[c]; t->inv_h[c] = b[c] / ext;
        }
        t->m = t->b[0] * t->b[1] * t->b[2]; t->nc = 6 * t->m;
        const TLGeom g = tl_geom(t);
        LAUNCH(ctx, k_tl_agg, div_up(nq, 256), 256, 0, (const double*)ctx->xq.p, nq, g, t->agg.p);
        CU(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
        LAUNCH(ctx, k_tl_check_adjacent, div_up(nq, 256), 256, 0, (const int*)ctx->blk_ptr.p, (const int*)ctx->blk_col.p, (const int*)t->agg.p, nq, g, bad.p);
        int hb = 0;
        CU(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->dist) {                                           // every rank must take the same decision
            double v = hb ? 1.0 : 0.0;
            CU(cudaMemcpyAsync(badd.p, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            TRY(dist_allreduce(ctx, badd.p, 1));
            CU(cudaMemcpyAsync(&v, badd.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            hb = v > 0.0;
        }
        if (!hb || t->m == 1) break;                               // boxes at least as wide as the cells: the 27 probing colours are valid
        if (target < 2) target = 2;
    }
    // nodes per box, ascending (stable counting sort over chunks of 256 nodes)
    const int nchunks = (nq + TL_CHUNK - 1) / TL_CHUNK;
    const size_t ncnt = (size_t)t->m * nchunks;
    DevBuf<int> cnt; CU(cnt.alloc(ncnt + 1));
    CU(cudaMemsetAsync(cnt.p, 0, (ncnt + 1) * sizeof(int), ctx->stream));
    LAUNCH(ctx, k_tl_count, nchunks, TL_CHUNK, 0, (const int*)t->agg.p, ctx->owned, nq, nchunks, cnt.p);
    i64 total = 0;
    TRY(scan_exclusive_i32(ctx, cnt.p, cnt.p, (i64)ncnt, &total));
    if (!ctx->dist && total != nq) return toe_fail(ctx, TOE_ERR_STATE, "two-level preconditioner: box lists hold %lld of %d nodes", (long long)total, nq);
    CU(t->agg_nodes.alloc(nq)); CU(t->agg_ptr.alloc(t->m + 1));
    LAUNCH(ctx, k_tl_fill, nchunks, TL_CHUNK, 0, (const int*)t->agg.p, ctx->owned, nq, nchunks, (const int*)cnt.p, t->agg_nodes.p);
    LAUNCH(ctx, k_tl_box_ptr, div_up(t->m + 1, 256), 256, 0, (const int*)cnt.p, nchunks, t->m, (int)total, t->agg_ptr.p);
    CU(t->A.alloc((size_t)t->nc * t->nc)); CU(t->w.alloc(t->nc + 8)); CU(t->y.alloc(t->nc + 8));
    const size_t n = 3 * (size_t)nq;
    CU(t->z.alloc(n)); CU(t->tv.alloc(n)); CU(t->ty.alloc(n));
    CU(cudaStreamSynchronize(ctx->stream));
    t->have_setup = true; t->built_generation = -1;
    return TOE_OK;
}

// A_c = ZᵀKZ by coloured probing with the current operator, then A_c⁻¹ in place
static int tl_build_inverse(toe_ctx* ctx, int matrix_free) {
    TwoLevel* t = ctx->tl;
    if (t->built_generation == ctx->op_generation && t->built_matrix_free == matrix_free) return TOE_OK;
    const TLGeom g = tl_geom(t);
    const int nq = ctx->nq, nc = t->nc, m = t->m;
    EventPair ev; CU(ev.create());
    cudaEvent_t a = ev.a, b = ev.b;
    CU(cudaEventRecord(a, ctx->stream));
    CU(cudaMemsetAsync(t->A.p, 0, (size_t)nc * nc * sizeof(double), ctx->stream));
    const int ncx = t->b[0] < 3 ? t->b[0] : 3, ncy = t->b[1] < 3 ? t->b[1] : 3, ncz = t->b[2] < 3 ? t->b[2] : 3;
    const bool direct = !matrix_free && ctx->have_K && !ctx->dist && !getenv("TOE_TL_PROBE");
    if (direct)
        LAUNCH(ctx, k_tl_coarse_direct, 27 * m, 128, 0, (const int*)t->agg_ptr.p, (const int*)t->agg_nodes.p, (const int*)t->agg.p, (const int*)ctx->blk_ptr.p,
               (const int*)ctx->blk_col.p, (const double*)ctx->val.p, ctx->ldv, (const double*)ctx->xq.p, (const unsigned char*)ctx->dflag.p, g, t->A.p, nc);
    else for (int cz = 0; cz < ncz; cz++) for (int cy = 0; cy < ncy; cy++) for (int cx = 0; cx < ncx; cx++)
        for (int mode = 0; mode < 6; mode++) {
            LAUNCH(ctx, k_tl_probe, div_up(nq, 256), 256, 0, (const int*)t->agg.p, (const double*)ctx->xq.p, (const unsigned char*)ctx->dflag.p, g, cx, cy, cz, mode, t->tv.p, nq);
            TRY(op_apply(ctx, t->tv.p, t->ty.p, matrix_free, nullptr, true));
            LAUNCH(ctx, k_tl_restrict, m, TL_RT, 0, (const int*)t->agg_ptr.p, (const int*)t->agg_nodes.p, (const double*)ctx->xq.p, (const unsigned char*)ctx->dflag.p,
                   (const double*)t->ty.p, g, t->w.p, (const int*)nullptr);
            TRY(dist_allreduce(ctx, t->w.p, nc));                 // partitioned: per-box sums of the owned nodes of every rank
            LAUNCH(ctx, k_tl_scatter, div_up(6 * m, 256), 256, 0, (const double*)t->w.p, g, cx, cy, cz, mode, t->A.p, m, nc);
        }
    LAUNCH(ctx, k_tl_symmetrise, div_up((i64)nc * nc, 256), 256, 0, t->A.p, nc);
    double* dmax = t->y.p + nc;                       // scratch behind y
    LAUNCH(ctx, k_tl_diag_scale, 1, 256, 0, (const double*)t->A.p, nc, dmax);
    DevBuf<int> skipped; CU(skipped.alloc(1));
    CU(cudaMemsetAsync(skipped.p, 0, sizeof(int), ctx->stream));
    if (getenv("TOE_TL_GJ_UNBLOCKED")) {                     // reference form: one pivot per step (kept for cross-checks)
        DevBuf<double> colk; CU(colk.alloc(nc));
        for (int k = 0; k < nc; k++) {
            LAUNCH(ctx, k_gj_row, 1, 256, 0, t->A.p, nc, k, (const double*)dmax, 1e-13, colk.p, skipped.p);
            LAUNCH(ctx, k_gj_update, nc, 256, 0, t->A.p, nc, k, (const double*)colk.p);
        }
        CU(cudaStreamSynchronize(ctx->stream));
    } else {
        DevBuf<int> skip; CU(skip.alloc(nc));
        DevBuf<double> Dinv, Cp, Rp; CU(Dinv.alloc(GJB * GJB)); CU(Cp.alloc((size_t)nc * GJB)); CU(Rp.alloc((size_t)GJB * nc));
        LAUNCH(ctx, k_gjb_flag_empty, div_up(nc, 256), 256, 0, (const double*)t->A.p, nc, (const double*)dmax, 1e-13, skip.p, skipped.p);
        LAUNCH(ctx, k_gjb_clear_empty, div_up((i64)nc * nc, 256), 256, 0, t->A.p, nc, (const int*)skip.p);
        const int tiles = (nc + 63) / 64;
        for (int k0 = 0; k0 < nc; k0 += GJB) {
            const int bs = nc - k0 < GJB ? nc - k0 : GJB;
            LAUNCH(ctx, k_gjb_diag, 1, 256, 0, (const double*)t->A.p, nc, k0, bs, (const double*)dmax, 1e-13, Dinv.p, skip.p, skipped.p);
            LAUNCH(ctx, k_gjb_panels, div_up(nc, 256), 256, 0, (const double*)t->A.p, nc, k0, bs, (const double*)Dinv.p, (const int*)skip.p, Cp.p, Rp.p);
            LAUNCH(ctx, k_gjb_update, tiles * tiles, 256, 0, t->A.p, nc, k0, bs, (const double*)Cp.p, (const double*)Rp.p, tiles);
            LAUNCH(ctx, k_gjb_write_rows, div_up((i64)bs * nc, 256), 256, 0, t->A.p, nc, k0, bs, (const double*)Rp.p);
        }
        CU(cudaStreamSynchronize(ctx->stream));               // the panels go out of scope
    }
    LAUNCH(ctx, k_tl_symmetrise, div_up((i64)nc * nc, 256), 256, 0, t->A.p, nc);
    CU(cudaEventRecord(b, ctx->stream));
    CU(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    CU(cudaGetLastError());
    t->setup_seconds = ms * 1e-3;
    t->built_generation = ctx->op_generation; t->built_matrix_free = matrix_free;
    return TOE_OK;
}

int tl_prepare(toe_ctx* ctx, int matrix_free, int* coarse_dofs, double* setup_seconds) {
    if (ctx->tl && ctx->tl->have_setup && (size_t)ctx->tl->z.n < 3 * (size_t)ctx->nq) ctx->tl->have_setup = false;
    TRY(tl_setup(ctx));
    ctx->tl->setup_seconds = 0.0;
    TRY(tl_build_inverse(ctx, matrix_free));
    if (coarse_dofs) *coarse_dofs = ctx->tl->nc;
    if (setup_seconds) *setup_seconds = ctx->tl->setup_seconds;
    return TOE_OK;
}

void tl_invalidate(toe_ctx* ctx) { if (ctx->tl) { ctx->tl->have_setup = false; ctx->tl->built_generation = -1; } }

static unsigned tl_vec_grid(size_t n) { return min_u(div_up((i64)n, 256), (unsigned)(N_SM * 8)); }

// opens the solve: x = 0, r = f, z = M⁻¹ r, p = z, γ = r·z
int tl_cg_init(toe_ctx* ctx, double atol, double rtol, i64 itmax, i64 hist_cap) {
    TwoLevel* t = ctx->tl;
    const TLGeom g = tl_geom(t);
    const size_t n = 3 * (size_t)ctx->nq;
    LAUNCH(ctx, k_tl_init, tl_vec_grid(n), 256, 0, (const double*)ctx->f.p, (const double*)ctx->diag.p, ctx->Minv.p, ctx->u.p, ctx->r.p, n);
    LAUNCH(ctx, k_tl_restrict, t->m, TL_RT, 0, (const int*)t->agg_ptr.p, (const int*)t->agg_nodes.p, (const double*)ctx->xq.p, (const unsigned char*)ctx->dflag.p,
           (const double*)ctx->r.p, g, t->w.p, (const int*)nullptr);
    TRY(dist_allreduce(ctx, t->w.p, t->nc));
    LAUNCH(ctx, k_tl_gemv, div_up(t->nc, 8), 256, 0, (const double*)t->A.p, (const double*)t->w.p, t->y.p, t->nc, (const int*)nullptr);
    double* gscal = ctx->dist ? &ctx->cgs.p->gd[1][0] : nullptr;       // partitioned: local γ partial → allreduce → k_tl_close
    LAUNCH(ctx, k_tl_z, tl_vec_grid(ctx->nq), 256, 0, (const double*)ctx->Minv.p, (const double*)ctx->r.p, (const int*)t->agg.p, (const double*)ctx->xq.p,
           (const double*)t->y.p, (const unsigned char*)ctx->dflag.p, g, t->z.p, ctx->nq, ctx->cgs.p, 1, atol, rtol, itmax, ctx->hist.p, hist_cap,
           ctx->partials.p, ctx->counters.p + 9, ctx->owned, gscal);
    if (ctx->dist) {
        TRY(dist_allreduce(ctx, gscal, 1));
        LAUNCH(ctx, k_tl_close, 1, 32, 0, ctx->cgs.p, (const double*)gscal, 1, atol, rtol, itmax, ctx->hist.p, hist_cap);
    }
    LAUNCH(ctx, k_tl_p, tl_vec_grid(n), 256, 0, (const double*)t->z.p, ctx->p.p, n, (const CGScalars*)ctx->cgs.p, 1);
    return TOE_OK;
}

// the part of one CG iteration after the operator (which already left p'Ap in the scalars): 5 kernels
int tl_cg_after_operator(toe_ctx* ctx, i64 hist_cap) {
    TwoLevel* t = ctx->tl;
    const TLGeom g = tl_geom(t);
    const size_t n = 3 * (size_t)ctx->nq;
    CGScalars* cg = ctx->cgs.p;
    if (ctx->dist) LAUNCH(ctx, k_tl_set_pAp, 1, 32, 0, cg, (const double*)&cg->gd[0][0]);      // p'Ap summed over the ranks by the caller's exchange
    LAUNCH(ctx, k_tl_xr, tl_vec_grid(n), 256, 0, (const double*)ctx->p.p, (const double*)ctx->Ap.p, ctx->u.p, ctx->r.p, n, (const CGScalars*)cg);
    LAUNCH(ctx, k_tl_restrict, t->m, TL_RT, 0, (const int*)t->agg_ptr.p, (const int*)t->agg_nodes.p, (const double*)ctx->xq.p, (const unsigned char*)ctx->dflag.p,
           (const double*)ctx->r.p, g, t->w.p, (const int*)&cg->done);
    TRY(dist_allreduce(ctx, t->w.p, t->nc));
    LAUNCH(ctx, k_tl_gemv, div_up(t->nc, 8), 256, 0, (const double*)t->A.p, (const double*)t->w.p, t->y.p, t->nc, (const int*)&cg->done);
    double* gscal = ctx->dist ? &cg->gd[1][0] : nullptr;
    LAUNCH(ctx, k_tl_z, tl_vec_grid(ctx->nq), 256, 0, (const double*)ctx->Minv.p, (const double*)ctx->r.p, (const int*)t->agg.p, (const double*)ctx->xq.p,
           (const double*)t->y.p, (const unsigned char*)ctx->dflag.p, g, t->z.p, ctx->nq, cg, 0, 0.0, 0.0, (i64)0, ctx->hist.p, hist_cap,
           ctx->partials.p, ctx->counters.p + 9, ctx->owned, gscal);
    if (ctx->dist) {
        TRY(dist_allreduce(ctx, gscal, 1));
        LAUNCH(ctx, k_tl_close, 1, 32, 0, cg, (const double*)gscal, 0, 0.0, 0.0, (i64)0, ctx->hist.p, hist_cap);
    }
    LAUNCH(ctx, k_tl_p, tl_vec_grid(n), 256, 0, (const double*)t->z.p, ctx->p.p, n, (const CGScalars*)cg, 0);
    return TOE_OK;
}
