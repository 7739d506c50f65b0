/* cantilever.c — the hot path through the C ABI alone (no Python, no Julia): the call sequence a host in any language makes.
 *
 *     gcc -std=c99 -O2 examples/cantilever.c -Iinclude -Ltopopteval.jl_b200 -ltopopt_b200 -Wl,-rpath,$PWD/topopteval.jl_b200 -lm -o cantilever
 *     ./cantilever 24 8 4          # nx ny nz cubes of the 60 x 20 x 4 beam, 6 tets per cube; needs a B200 (there is no CPU fallback)
 *
 * Mirrors test/runtests.jl:28-41 of the reference: setup_problem → assemble_stiffness_matrix! → apply_fixed_boundary! (x = 0) →
 * apply_force! (total [0,0,-1] on x = 60) → solve_system → deformation energy, max von Mises.  The mesh is the structured 6-tet split
 * of tests (topopteval.jl_b200/meshgen.py): node i + (nx+1)(j + (ny+1)k), tets (1,2,4,8),(1,5,2,8),(2,3,4,8),(2,7,3,8),(2,5,6,8),(2,6,7,8)
 * of each cube's VTK-hexahedron-ordered corners. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "topopt_b200.h"

#define CHECK(call) do { int st_ = (call); if (st_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, st_, toe_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int nx = argc > 1 ? atoi(argv[1]) : 24, ny = argc > 2 ? atoi(argv[2]) : 8, nz = argc > 3 ? atoi(argv[3]) : 4;
    const double L[3] = {60.0, 20.0, 4.0};
    const int64_t nn = (int64_t)(nx + 1) * (ny + 1) * (nz + 1), ne = 6LL * nx * ny * nz;
    double* xyz = (double*)malloc(sizeof(double) * 3 * nn);
    int64_t* conn = (int64_t*)malloc(sizeof(int64_t) * 4 * ne);
    int64_t* fixed = (int64_t*)malloc(sizeof(int64_t) * (ny + 1) * (nz + 1));
    int64_t* load = (int64_t*)malloc(sizeof(int64_t) * (ny + 1) * (nz + 1));
    int64_t nfixed = 0, nload = 0, e = 0;
    static const int tets[6][4] = {{0, 1, 3, 7}, {0, 4, 1, 7}, {1, 2, 3, 7}, {1, 6, 2, 7}, {1, 4, 5, 7}, {1, 5, 6, 7}};
    for (int k = 0; k <= nz; k++) for (int j = 0; j <= ny; j++) for (int i = 0; i <= nx; i++) {
        const int64_t g = i + (int64_t)(nx + 1) * (j + (int64_t)(ny + 1) * k);
        xyz[3 * g] = L[0] * i / nx; xyz[3 * g + 1] = L[1] * j / ny; xyz[3 * g + 2] = L[2] * k / nz;
        if (i == 0) fixed[nfixed++] = g + 1;              /* node ids are 1-based at the ABI, like the reference's */
        if (i == nx) load[nload++] = g + 1;
    }
    for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
        int64_t c[8];
        const int di[8] = {0, 1, 1, 0, 0, 1, 1, 0}, dj[8] = {0, 0, 1, 1, 0, 0, 1, 1}, dk[8] = {0, 0, 0, 0, 1, 1, 1, 1};
        for (int a = 0; a < 8; a++) c[a] = (i + di[a]) + (int64_t)(nx + 1) * ((j + dj[a]) + (int64_t)(ny + 1) * (k + dk[a])) + 1;
        for (int t = 0; t < 6; t++, e++) for (int a = 0; a < 4; a++) conn[4 * e + a] = c[tets[t][a]];
    }

    toe_ctx* ctx = NULL;
    if (toe_create(0, &ctx) != 0) { fprintf(stderr, "toe_create: %s\n", toe_last_error(NULL)); return 1; }
    int64_t ndofs = 0, nnz = 0;
    CHECK(toe_set_mesh(ctx, nn, xyz, ne, 4, conn));                       /* setup_problem */
    CHECK(toe_build_dofs(ctx, &ndofs));
    CHECK(toe_build_pattern(ctx, &nnz));
    const double E = 1.0, nu = 0.3;                                       /* create_material_model */
    const double lambda = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
    CHECK(toe_assemble_lame(ctx, lambda, mu, 0));                         /* assemble_stiffness_matrix! */
    const double F[3] = {0.0, 0.0, -1.0};
    CHECK(toe_add_nodal_force(ctx, load, nload, F));                      /* apply_force! */
    int64_t* node_dof = (int64_t*)malloc(sizeof(int64_t) * nn);           /* apply_fixed_boundary!: all 3 components of the x = 0 nodes */
    CHECK(toe_get_node_dofs(ctx, node_dof));
    int64_t* pres = (int64_t*)malloc(sizeof(int64_t) * 3 * nfixed);
    for (int64_t i = 0; i < nfixed; i++) for (int c = 0; c < 3; c++) pres[3 * i + c] = node_dof[fixed[i] - 1] + c;
    double mean_diag = 0.0;
    CHECK(toe_apply_dirichlet(ctx, pres, 3 * nfixed, &mean_diag));        /* the single apply!(K, f, ch) of solve_system */
    toe_pcg_stats st;
    CHECK(toe_solve_pcg(ctx, 1e-8, 1e-8, 100000, 0, &st, NULL, 0));       /* solve (Jacobi-PCG, Krylov.jl stopping rule) */
    double energy = 0.0, compliance = 0.0, max_vm = 0.0;
    int64_t max_cell = 0;
    CHECK(toe_energy(ctx, &energy, &compliance, NULL));
    CHECK(toe_stresses(ctx, NULL, NULL, &max_vm, &max_cell));
    printf("%lld tets, %lld DOFs, nnz %lld | PCG %lld iterations (converged %d, %.3f s) | deformation energy %.10g, compliance %.10g, "
           "max von Mises %.6g in cell %lld\n", (long long)ne, (long long)ndofs, (long long)nnz, (long long)st.niter, (int)st.converged,
           st.solve_seconds, energy, compliance, max_vm, (long long)max_cell);
    toe_destroy(ctx);
    free(xyz); free(conn); free(fixed); free(load); free(node_dof); free(pres);
    return st.converged ? 0 : 2;
}
