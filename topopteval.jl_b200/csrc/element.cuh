// element.cuh — per-element device math: Tet4 / Hex8 geometry, isotropic Hooke, SIMP modulus.
//
// Restates what the reference's inner loops compute (FiniteElementAnalysis.jl:214-243 / :664-700):
//   Ke[(a,c),(b,d)] = Σ_q [ λ g_ac g_bd + μ( δ_cd g_a·g_b + g_ad g_bc ) ] detJ_q w_q ,  g_a = ∇N_a
// (closed form of sym(∇N_i) ⊡ (λ tr(ε_j) I + 2μ ε_j) for the basis i=(a,c) ↦ e_c ⊗ ∇N_a).
// Tet4: gradients are constant, Σ_q detJ w_q = detJ/6.  Hex8: 2x2x2 Gauss, points ±1/√3, weights 1.
#pragma once
#include "common.cuh"

__device__ __forceinline__ void material_at(const Material& m, int e, double& lam, double& mu) {
    if (m.mode == MAT_UNIFORM) { lam = m.lambda; mu = m.mu; return; }
    if (m.mode == MAT_SIMP) {
        // create_simp_material_model closure, FiniteElementAnalysis.jl:622-631
        double rho = __ldg(&m.density[e]);
        double E = m.Emin + (m.E0 - m.Emin) * pow(rho, m.p);
        lam = E * m.nu / ((1.0 + m.nu) * (1.0 - 2.0 * m.nu));
        mu = E / (2.0 * (1.0 + m.nu));
        return;
    }
    lam = __ldg(&m.lam_e[e]); mu = __ldg(&m.mu_e[e]);
}

__device__ __forceinline__ void load3(const double* __restrict__ x, int q, double* o) {
    const double* p = x + 3 * (size_t)q;
    o[0] = __ldg(p); o[1] = __ldg(p + 1); o[2] = __ldg(p + 2);
}

// 3x3 inverse via adjugate. J[i][j]; returns det, Jinv[j][i] such that Jinv*J = I
__device__ __forceinline__ double inv3(const double J[3][3], double Ji[3][3]) {
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    double id = 1.0 / det;
    Ji[0][0] = c00 * id;
    Ji[1][0] = c01 * id;
    Ji[2][0] = c02 * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    return det;
}

// Tet4: g[a][i] = ∂N_a/∂x_i, returns detJ (volume = detJ/6).  Reference vertices (0,0,0),(1,0,0),(0,1,0),(0,0,1).
__device__ __forceinline__ double tet_grads(const double X[4][3], double g[4][3]) {
    double J[3][3], Ji[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) J[i][j] = X[j + 1][i] - X[0][i];
    double det = inv3(J, Ji);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        g[1][i] = Ji[0][i]; g[2][i] = Ji[1][i]; g[3][i] = Ji[2][i];
        g[0][i] = -(Ji[0][i] + Ji[1][i] + Ji[2][i]);
    }
    return det;
}

__device__ __forceinline__ void tet_load(const int* __restrict__ cq, const double* __restrict__ xq, int e, int q[4], double X[4][3]) {
    int4 c = __ldg(reinterpret_cast<const int4*>(cq) + e);
    q[0] = c.x; q[1] = c.y; q[2] = c.z; q[3] = c.w;
#pragma unroll
    for (int a = 0; a < 4; a++) load3(xq, q[a], X[a]);
}

// 3x3 block (a,b) of Ke given gradients and coefficient w = detJ*weight: B[c][d]
__device__ __forceinline__ void block_ab(const double* ga, const double* gb, double lam, double mu, double w, double B[9], bool accumulate) {
    double dot = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double v = lam * (ga[c] * gb[d]) + mu * ((c == d ? dot : 0.0) + ga[d] * gb[c]);
            v *= w;
            if (accumulate) B[3 * c + d] += v; else B[3 * c + d] = v;
        }
}

// Hex8 (VTK order) reference signs
__device__ __forceinline__ void hex_sign(int a, double& sx, double& sy, double& sz) {
    sx = ((a & 3) == 1 || (a & 3) == 2) ? 1.0 : -1.0;
    sy = (a & 2) ? 1.0 : -1.0;
    sz = (a & 4) ? 1.0 : -1.0;
}

// Gauss point gp (0..7), x fastest: ξ = ±1/√3
__device__ __forceinline__ void hex_qp(int gp, double& xi, double& eta, double& zeta) {
    const double g = 0.57735026918962576451;
    xi = (gp & 1) ? g : -g; eta = (gp & 2) ? g : -g; zeta = (gp & 4) ? g : -g;
}

// shape values and physical gradients of all 8 nodes at one Gauss point; returns detJ (weight 1)
__device__ __forceinline__ double hex_grads_at(const double X[8][3], int gp, double g[8][3], double N[8]) {
    double xi, eta, zeta; hex_qp(gp, xi, eta, zeta);
    double dn[8][3];
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Ji[3][3];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        double sx, sy, sz; hex_sign(a, sx, sy, sz);
        double fx = 1.0 + xi * sx, fy = 1.0 + eta * sy, fz = 1.0 + zeta * sz;
        N[a] = 0.125 * fx * fy * fz;
        dn[a][0] = 0.125 * sx * fy * fz;
        dn[a][1] = 0.125 * sy * fx * fz;
        dn[a][2] = 0.125 * sz * fx * fy;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) J[i][j] += X[a][i] * dn[a][j];
    }
    double det = inv3(J, Ji);
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int i = 0; i < 3; i++) g[a][i] = dn[a][0] * Ji[0][i] + dn[a][1] * Ji[1][i] + dn[a][2] * Ji[2][i];
    return det;
}

__device__ __forceinline__ void hex_load(const int* __restrict__ cq, const double* __restrict__ xq, int e, int q[8], double X[8][3]) {
    const int4* c = reinterpret_cast<const int4*>(cq) + 2 * (size_t)e;
    int4 c0 = __ldg(c), c1 = __ldg(c + 1);
    q[0] = c0.x; q[1] = c0.y; q[2] = c0.z; q[3] = c0.w; q[4] = c1.x; q[5] = c1.y; q[6] = c1.z; q[7] = c1.w;
#pragma unroll
    for (int a = 0; a < 8; a++) load3(xq, q[a], X[a]);
}

// σ·g for isotropic Hooke from a displacement gradient H[c][i] = Σ_b u_b[c] g_b[i]:
//   ε = sym(H), σ = λ tr(ε) I + 2μ ε   (constitutive_relation, FiniteElementAnalysis.jl:126-129)
__device__ __forceinline__ void hooke_from_grad(const double H[3][3], double lam, double mu, double S[3][3]) {
    double tr = H[0][0] + H[1][1] + H[2][2];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int d = 0; d < 3; d++) S[c][d] = mu * (H[c][d] + H[d][c]) + (c == d ? lam * tr : 0.0);
}
