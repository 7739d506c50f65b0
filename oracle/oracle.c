/*
 * oracle.c — CPU ORACLE, TEST / BASELINE INFRASTRUCTURE ONLY.  Never linked into or called by the product path.
 *
 * Plain-C, single-threaded restatement of the reference's hot path with the reference's own loop structure, used
 * (a) as a second, independent checker next to oracle/fea_oracle.py and (b) as the timed CPU baseline of bench.py
 * (`cpu_baseline.kind = "port"`): the genuine reference is Julia (+ Ferrite/Tensors/Krylov), which is not installed
 * in this image, so it cannot be compiled into oracle/_ref.  PARITY UNPINNED (see oracle/fea_oracle.py header).
 *
 *   first-touch DOF numbering    Ferrite close!(dh)            call site FiniteElementAnalysis.jl:174-176
 *   sorted CSC pattern           Ferrite allocate_matrix(dh)   :181
 *   Ke by the (q,i,j) loops      FiniteElementAnalysis.jl:218-243 / :677-699  (sym, Hooke, double contraction)
 *   sorted-merge assembly        Ferrite assemble!             :246 / :703
 *   Dirichlet                    Ferrite apply!(K,f,ch)        :540-542
 *   Jacobi-PCG                   Krylov.jl cg with M           RobustSolver.jl:231-236, 294-305, 337
 *   energy                       0.5*dot(u,K*u)                :550
 * All indices 0-based inside this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ---- reference element (Ferrite Lagrange{Ref*,1}, QuadratureRule{Ref*}(2)) ------------------------------- */
static const double HEX_S[8][3] = {{-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1}};

static int ref_element(int npc, double N[8][8], double dN[8][8][3], double w[8]) {
    if (npc == 4) {
        double a = (5.0 - sqrt(5.0)) / 20.0, b = (5.0 + 3.0 * sqrt(5.0)) / 20.0;
        double qp[4][3] = {{a,a,a},{a,a,b},{a,b,a},{b,a,a}};
        for (int q = 0; q < 4; q++) {
            w[q] = 1.0 / 24.0;
            N[q][0] = 1 - qp[q][0] - qp[q][1] - qp[q][2]; N[q][1] = qp[q][0]; N[q][2] = qp[q][1]; N[q][3] = qp[q][2];
            double d[4][3] = {{-1,-1,-1},{1,0,0},{0,1,0},{0,0,1}};
            memcpy(dN[q], d, sizeof d);
        }
        return 4;
    }
    double g = 1.0 / sqrt(3.0);
    int q = 0;
    for (int kz = -1; kz <= 1; kz += 2) for (int ky = -1; ky <= 1; ky += 2) for (int kx = -1; kx <= 1; kx += 2, q++) {
        double x = kx * g, y = ky * g, z = kz * g;
        w[q] = 1.0;
        for (int a = 0; a < 8; a++) {
            const double* s = HEX_S[a];
            N[q][a] = 0.125 * (1 + x * s[0]) * (1 + y * s[1]) * (1 + z * s[2]);
            dN[q][a][0] = 0.125 * s[0] * (1 + y * s[1]) * (1 + z * s[2]);
            dN[q][a][1] = 0.125 * s[1] * (1 + x * s[0]) * (1 + z * s[2]);
            dN[q][a][2] = 0.125 * s[2] * (1 + x * s[0]) * (1 + y * s[1]);
        }
    }
    return 8;
}

/* Ferrite reinit!: J = Σ x_a ⊗ dN_a/dξ, dNdx = dNdξ · J⁻¹, returns det J (caller rejects <= 0) */
static double reinit_qp(int npc, const double X[8][3], const double dN[8][3], double dNdx[8][3]) {
    double J[3][3] = {{0}};
    for (int a = 0; a < npc; a++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[i][j] += X[a][i] * dN[a][j];
    double c00 = J[1][1]*J[2][2]-J[1][2]*J[2][1], c01 = J[1][2]*J[2][0]-J[1][0]*J[2][2], c02 = J[1][0]*J[2][1]-J[1][1]*J[2][0];
    double det = J[0][0]*c00 + J[0][1]*c01 + J[0][2]*c02, id = 1.0 / det;
    double Ji[3][3];
    Ji[0][0]=c00*id; Ji[1][0]=c01*id; Ji[2][0]=c02*id;
    Ji[0][1]=(J[0][2]*J[2][1]-J[0][1]*J[2][2])*id; Ji[1][1]=(J[0][0]*J[2][2]-J[0][2]*J[2][0])*id; Ji[2][1]=(J[0][1]*J[2][0]-J[0][0]*J[2][1])*id;
    Ji[0][2]=(J[0][1]*J[1][2]-J[0][2]*J[1][1])*id; Ji[1][2]=(J[0][2]*J[1][0]-J[0][0]*J[1][2])*id; Ji[2][2]=(J[0][0]*J[1][1]-J[0][1]*J[1][0])*id;
    for (int a = 0; a < npc; a++) for (int i = 0; i < 3; i++)
        dNdx[a][i] = dN[a][0]*Ji[0][i] + dN[a][1]*Ji[1][i] + dN[a][2]*Ji[2][i];
    return det;
}

/* ---- setup_problem ----------------------------------------------------------------------------------------- */
/* returns ndofs; node_first_dof[nn] (-1 = none), cell_dofs[ne*3*npc] */
i64 oracle_first_touch(i64 ne, int npc, const i64* cells1, i64 nn, i64* node_first_dof, i64* cell_dofs) {
    for (i64 g = 0; g < nn; g++) node_first_dof[g] = -1;
    i64 next = 0;
    for (i64 e = 0; e < ne; e++) for (int v = 0; v < npc; v++) {
        i64 g = cells1[e * npc + v] - 1;
        if (node_first_dof[g] < 0) { node_first_dof[g] = next; next += 3; }
        for (int c = 0; c < 3; c++) cell_dofs[(e * npc + v) * 3 + c] = node_first_dof[g] + c;
    }
    return next;
}

static int cmp_i64(const void* a, const void* b) { i64 x = *(const i64*)a, y = *(const i64*)b; return (x > y) - (x < y); }

/* sorted CSC pattern: colptr[n+1] filled; returns nnz and a malloc'ed rowval in *rowval_out (caller frees with oracle_free) */
i64 oracle_pattern(i64 ne, int nb, const i64* cell_dofs, i64 n, i64* colptr, i64** rowval_out) {
    i64 total = ne * nb * nb;
    i64* key = (i64*)malloc(sizeof(i64) * total);
    i64 k = 0;
    for (i64 e = 0; e < ne; e++) { const i64* cd = cell_dofs + e * nb;
        for (int j = 0; j < nb; j++) for (int i = 0; i < nb; i++) key[k++] = cd[j] * n + cd[i]; }
    qsort(key, total, sizeof(i64), cmp_i64);
    i64 nnz = 0;
    for (i64 t = 0; t < total; t++) if (t == 0 || key[t] != key[t - 1]) key[nnz++] = key[t];
    i64* rowval = (i64*)malloc(sizeof(i64) * nnz);
    memset(colptr, 0, sizeof(i64) * (n + 1));
    for (i64 t = 0; t < nnz; t++) { i64 col = key[t] / n; rowval[t] = key[t] - col * n; colptr[col + 1]++; }
    for (i64 j = 0; j < n; j++) colptr[j + 1] += colptr[j];
    free(key);
    *rowval_out = rowval;
    return nnz;
}
void oracle_free(void* p) { free(p); }

/* ---- element stiffness by the reference's loop nest (:218-243) ------------------------------------------------ */
static void sym_grad(const double gN[3], int c, double eps[3][3]) {     /* symmetric(e_c ⊗ ∇N) */
    double G[3][3] = {{0}};
    for (int i = 0; i < 3; i++) G[c][i] = gN[i];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) eps[i][j] = 0.5 * (G[i][j] + G[j][i]);
}

/* ke: nb*nb row-major ke[i*nb+j]; returns min det J */
double oracle_ke(int npc, const double X[8][3], double lam, double mu, double* ke) {
    double N[8][8], dN[8][8][3], w[8], dNdx[8][3];
    int nq = ref_element(npc, N, dN, w), nb = 3 * npc;
    double mindet = 1e300;
    memset(ke, 0, sizeof(double) * nb * nb);
    for (int q = 0; q < nq; q++) {
        double det = reinit_qp(npc, X, dN[q], dNdx);
        if (det < mindet) mindet = det;
        double dOm = det * w[q];                                        /* getdetJdV */
        for (int i = 0; i < nb; i++) {
            for (int j = 0; j < nb; j++) {
                double ei[3][3], ej[3][3], sig[3][3];
                sym_grad(dNdx[i / 3], i % 3, ei);                       /* εi = symmetric(∇Ni) */
                sym_grad(dNdx[j / 3], j % 3, ej);                       /* εj = symmetric(∇Nj) */
                double tr = ej[0][0] + ej[1][1] + ej[2][2];
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) /* σ = λ tr(ε) I + 2μ ε */
                    sig[a][b] = (a == b ? lam * tr : 0.0) + 2.0 * mu * ej[a][b];
                double dc = 0.0;
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) dc += ei[a][b] * sig[a][b];   /* εi ⊡ σ */
                ke[i * nb + j] += dc * dOm;
            }
        }
    }
    return mindet;
}

/* material per cell: mode 0 = (lam,mu) uniform in par[0..1]; mode 1 = SIMP par = (E0,nu,Emin,p) with density */
static void material(int mode, const double* par, const double* density, i64 e, double* lam, double* mu) {
    if (mode == 0) { *lam = par[0]; *mu = par[1]; return; }
    double E = par[2] + (par[0] - par[2]) * pow(density[e], par[3]);    /* :624 */
    *lam = E * par[1] / ((1 + par[1]) * (1 - 2 * par[1]));
    *mu = E / (2 * (1 + par[1]));
}

/* Ke of a range of cells, for parity checks: out[(e-first)*nb*nb + i*nb + j] */
int oracle_ke_batch(i64 first, i64 count, int npc, const i64* cells1, const double* xyz, int mode, const double* par,
                    const double* density, double* out) {
    int nb = 3 * npc;
    for (i64 e = first; e < first + count; e++) {
        double X[8][3], lam, mu;
        for (int a = 0; a < npc; a++) memcpy(X[a], xyz + 3 * (cells1[e * npc + a] - 1), 3 * sizeof(double));
        material(mode, par, density, e, &lam, &mu);
        if (oracle_ke(npc, X, lam, mu, out + (e - first) * nb * nb) <= 0) return -1;
    }
    return 0;
}

/* assemble_stiffness_matrix(_simp)!: zero K, per cell Ke then Ferrite assemble! (sort dofs, merge into the sorted column) */
int oracle_assemble(i64 ne, int npc, const i64* cells1, const double* xyz, const i64* cell_dofs, int mode, const double* par,
                    const double* density, i64 n, const i64* colptr, const i64* rowval, double* nzval) {
    int nb = 3 * npc;
    double* ke = (double*)malloc(sizeof(double) * nb * nb);
    memset(nzval, 0, sizeof(double) * colptr[n]);
    for (i64 e = 0; e < ne; e++) {
        double X[8][3], lam, mu;
        for (int a = 0; a < npc; a++) memcpy(X[a], xyz + 3 * (cells1[e * npc + a] - 1), 3 * sizeof(double));
        material(mode, par, density, e, &lam, &mu);
        if (oracle_ke(npc, X, lam, mu, ke) <= 0) { free(ke); return -1; }
        const i64* cd = cell_dofs + e * nb;
        int perm[24];
        for (int i = 0; i < nb; i++) perm[i] = i;
        for (int i = 1; i < nb; i++) { int p = perm[i], j = i - 1; while (j >= 0 && cd[perm[j]] > cd[p]) { perm[j + 1] = perm[j]; j--; } perm[j + 1] = p; }
        for (int jj = 0; jj < nb; jj++) {
            int j = perm[jj];
            i64 col = cd[j], r = colptr[col];
            for (int ii = 0; ii < nb; ii++) {            /* sorted local rows vs sorted rowval: linear merge */
                int i = perm[ii];
                while (rowval[r] < cd[i]) r++;
                nzval[r] += ke[i * nb + j];
            }
        }
    }
    free(ke);
    return 0;
}

/* Ferrite apply!(K,f,ch), zero-valued constraints; flag[n] marks prescribed dofs; returns m */
double oracle_apply_dirichlet(i64 n, const i64* colptr, const i64* rowval, double* nzval, double* f, const unsigned char* flag) {
    double m = 0.0;
    for (i64 j = 0; j < n; j++) for (i64 k = colptr[j]; k < colptr[j + 1]; k++) if (rowval[k] == j) m += fabs(nzval[k]);
    m /= (double)n;
    for (i64 j = 0; j < n; j++) for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
        i64 i = rowval[k];
        if (flag[i] || flag[j]) nzval[k] = (i == j) ? m : 0.0;
    }
    for (i64 j = 0; j < n; j++) if (flag[j]) f[j] = 0.0;
    return m;
}

/* y = K x, CSC (what SparseArrays does for K*p) */
void oracle_spmv(i64 n, const i64* colptr, const i64* rowval, const double* nzval, const double* x, double* y) {
    memset(y, 0, sizeof(double) * n);
    for (i64 j = 0; j < n; j++) { double xj = x[j]; for (i64 k = colptr[j]; k < colptr[j + 1]; k++) y[rowval[k]] += nzval[k] * xj; }
}

/* Krylov.jl cg with M = Diagonal(1 ./ D), D[|D|<1e-12] = 1; returns niter, *solved; residuals (may be NULL) gets niter+1 entries */
i64 oracle_pcg(i64 n, const i64* colptr, const i64* rowval, const double* nzval, const double* b, double atol, double rtol,
               i64 itmax, double* x, int* solved, double* residuals) {
    double *r = malloc(8 * n), *z = malloc(8 * n), *p = malloc(8 * n), *Ap = malloc(8 * n), *Mi = malloc(8 * n);
    for (i64 j = 0; j < n; j++) { double d = 0.0; for (i64 k = colptr[j]; k < colptr[j + 1]; k++) if (rowval[k] == j) d = nzval[k];
        if (fabs(d) < 1e-12) d = 1.0; Mi[j] = 1.0 / d; }
    double gamma = 0.0;
    for (i64 i = 0; i < n; i++) { x[i] = 0.0; r[i] = b[i]; z[i] = Mi[i] * r[i]; p[i] = z[i]; gamma += r[i] * z[i]; }
    double eps = atol + rtol * sqrt(gamma);
    i64 k = 0;
    if (residuals) residuals[0] = sqrt(gamma);
    while (sqrt(gamma) > eps && k < itmax) {
        oracle_spmv(n, colptr, rowval, nzval, p, Ap);
        double pAp = 0.0; for (i64 i = 0; i < n; i++) pAp += p[i] * Ap[i];
        if (pAp <= 0) break;
        double alpha = gamma / pAp, gnew = 0.0;
        for (i64 i = 0; i < n; i++) x[i] += alpha * p[i];
        for (i64 i = 0; i < n; i++) r[i] -= alpha * Ap[i];
        for (i64 i = 0; i < n; i++) z[i] = Mi[i] * r[i];
        for (i64 i = 0; i < n; i++) gnew += r[i] * z[i];
        double beta = gnew / gamma;
        for (i64 i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        gamma = gnew; k++;
        if (residuals) residuals[k] = sqrt(gamma);
    }
    *solved = sqrt(gamma) <= eps;
    free(r); free(z); free(p); free(Ap); free(Mi);
    return k;
}

double oracle_energy(i64 n, const i64* colptr, const i64* rowval, const double* nzval, const double* u) {
    double* y = malloc(8 * n); oracle_spmv(n, colptr, rowval, nzval, u, y);
    double s = 0.0; for (i64 i = 0; i < n; i++) s += u[i] * y[i];
    free(y); return 0.5 * s;
}
