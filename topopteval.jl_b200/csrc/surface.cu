// surface.cu — boundary-node selection and surface traction on the GPU.
//
// Reference loops replaced:
//   extract_surface_nodes!            SelectNodesForBC.jl:59-123   faces (sorted node tuples) counted in a Dict; count == 1 → surface
//   select_surface_nodes_by_plane     :146-185                     abs(dot(x - p, n/‖n‖)) < tol over the surface nodes
//   select_surface_nodes_by_circle    :207-266                     plane selection, then in-plane distance <= radius + tol
//   get_boundary_facets               SurfaceTraction.jl:45-66     (cell, local face) whose vertices all lie in a node set
//   compute_boundary_area             :88-122                      Σ dΓ over FacetQuadratureRule{Ref*}(2)
//   apply_surface_traction!           :160-225                     f[celldofs] += Σ_q (N_i · t(x_q)) dΓ_q
// Face tables = get_face_nodes (FiniteElementAnalysis.jl:42-56), Ferrite's facet order.
//
// The O(ne) hashing pass of the reference becomes a search: a face of cell e is interior iff another cell of the incidence
// list of its first node holds all of its nodes — no hash table, no sort; the incidence lists already exist (mesh.cu).
#include "common.cuh"
#include <vector>

__constant__ int c_tet_face[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};
__constant__ int c_hex_face[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {4, 5, 6, 7}};

template <int NPC>
__device__ __forceinline__ int face_local(int f, int k) { return NPC == 4 ? c_tet_face[f][k] : c_hex_face[f][k]; }

// one thread per (cell, face): flags the dof-nodes of faces that no other cell shares
template <int NPC>
__global__ void k_surface_faces(const int* __restrict__ inc_ptr, const int* __restrict__ inc, const int* __restrict__ cq, i64 ne, int* __restrict__ qflag) {
    const int NF = NPC == 4 ? 4 : 6, FN = NPC == 4 ? 3 : 4;
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ne * NF) return;
    const i64 e = t / NF; const int f = (int)(t - e * NF);
    int fn[4];
#pragma unroll
    for (int k = 0; k < FN; k++) fn[k] = cq[e * NPC + face_local<NPC>(f, k)];
    bool shared = false;
    for (int i = inc_ptr[fn[0]]; i < inc_ptr[fn[0] + 1] && !shared; i++) {
        const int e2 = inc[i] / NPC;
        if (e2 == (int)e) continue;
        const int* c2 = cq + (size_t)e2 * NPC;
        bool all = true;
#pragma unroll
        for (int k = 1; k < FN; k++) {
            bool has = false;
#pragma unroll
            for (int a = 0; a < NPC; a++) has |= (c2[a] == fn[k]);
            all &= has;
        }
        shared = all;
    }
    if (!shared) {
#pragma unroll
        for (int k = 0; k < FN; k++) qflag[fn[k]] = 1;
    }
}

__global__ void k_surface_node_flags(const int* __restrict__ node_q, const int* __restrict__ qflag, i64 nn, int* __restrict__ flag) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nn) flag[g] = (node_q[g] >= 0 && qflag[node_q[g]]) ? 1 : 0;
}

// mode 0: all surface nodes; 1: plane; 2: circle.  n = unit normal.
__global__ void k_select(const int* __restrict__ surf, const double* __restrict__ xyz, i64 nn, int mode, double px, double py, double pz,
                         double nx, double ny, double nz, double tol, double radius, int* __restrict__ out) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nn) return;
    int sel = surf[g];
    if (sel && mode >= 1) {
        const double vx = xyz[3 * g] - px, vy = xyz[3 * g + 1] - py, vz = xyz[3 * g + 2] - pz;
        const double d = vx * nx + vy * ny + vz * nz;
        sel = fabs(d) < tol;                                              // strict, :172
        if (sel && mode == 2) {
            const double qx = vx - d * nx, qy = vy - d * ny, qz = vz - d * nz;
            sel = sqrt(qx * qx + qy * qy + qz * qz) <= radius + tol;      // :252-258
        }
    }
    out[g] = sel ? 1 : 0;
}

__global__ void k_compact_ids(const int* __restrict__ flag, const int* __restrict__ pos, i64 nn, int64_t* __restrict__ out) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nn && flag[g]) out[pos[g]] = g + 1;
}

__global__ void k_flag_nodes(const int64_t* __restrict__ nodes, i64 n, i64 nn, unsigned char* __restrict__ flag, int* err) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t g = nodes[i];
    if (g < 1 || g > nn) { atomicExch(err + 2, 1); return; }
    flag[g - 1] = 1;
}

// one thread per (cell, face): 1 if every vertex of the face is flagged (interior faces qualify too, like the reference)
template <int NPC>
__global__ void k_facet_flags(const int* __restrict__ conn0, i64 ne, const unsigned char* __restrict__ nflag, int* __restrict__ out) {
    const int NF = NPC == 4 ? 4 : 6, FN = NPC == 4 ? 3 : 4;
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ne * NF) return;
    const i64 e = t / NF; const int f = (int)(t - e * NF);
    bool all = true;
#pragma unroll
    for (int k = 0; k < FN; k++) all &= nflag[conn0[e * NPC + face_local<NPC>(f, k)]] != 0;
    out[t] = all ? 1 : 0;
}
template <int NPC>
__global__ void k_compact_facets(const int* __restrict__ flag, const int* __restrict__ pos, i64 total, int64_t* __restrict__ out) {
    const int NF = NPC == 4 ? 4 : 6;
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total || !flag[t]) return;
    out[2 * (size_t)pos[t]] = t / NF + 1; out[2 * (size_t)pos[t] + 1] = t % NF + 1;
}

// one thread per facet: quadrature points, dΓ and — when vload != null — the loads Σ_q N_a t(x_q) dΓ_q of its vertices, left in
// vload[(i*4+k)*3+c] with the vertex's dof-node in vnode[i*4+k] (-1: none); k_vertex_loads_to_f adds them up in a fixed order.
// traction: per-qp values t_qp (3 per point) or, if null, the uniform vector (tx,ty,tz).
template <int NPC>
__global__ void k_facets(const int64_t* __restrict__ facets, i64 nf, const int* __restrict__ conn0, const double* __restrict__ xyz, i64 ne,
                         const int* __restrict__ node_q, double* __restrict__ xq_out, double* __restrict__ dg_out,
                         const double* __restrict__ t_qp, double tx, double ty, double tz, double* __restrict__ vload, int* __restrict__ vnode,
                         double* __restrict__ area_part, double* __restrict__ force_part, int* err) {
    const int FN = NPC == 4 ? 3 : 4, NQP = NPC == 4 ? 3 : 4, NF = NPC == 4 ? 4 : 6;
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const int64_t e = facets[2 * i] - 1, fid = facets[2 * i + 1] - 1;
    double area = 0.0, tot[3] = {0, 0, 0};
    if (e < 0 || e >= ne || fid < 0 || fid >= NF) {
        atomicExch(err + 2, 1);
        if (vload) { for (int k = 0; k < 4; k++) vnode[4 * i + k] = -1; }
    } else {
        int g[4]; double P[4][3];
#pragma unroll
        for (int k = 0; k < FN; k++) {
            g[k] = conn0[e * NPC + face_local<NPC>((int)fid, k)];
#pragma unroll
            for (int c = 0; c < 3; c++) P[k][c] = xyz[3 * (size_t)g[k] + c];
        }
        double load[4][3];
#pragma unroll
        for (int k = 0; k < 4; k++) load[k][0] = load[k][1] = load[k][2] = 0.0;
        for (int q = 0; q < NQP; q++) {
            double N[4] = {0, 0, 0, 0}, dg, x[3];
            if (NPC == 4) {
                const double ax = P[1][0] - P[0][0], ay = P[1][1] - P[0][1], az = P[1][2] - P[0][2];
                const double bx = P[2][0] - P[0][0], by = P[2][1] - P[0][1], bz = P[2][2] - P[0][2];
                const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
                dg = sqrt(cx * cx + cy * cy + cz * cz) * (1.0 / 6.0);
#pragma unroll
                for (int k = 0; k < 3; k++) N[k] = (k == q) ? 2.0 / 3.0 : 1.0 / 6.0;
            } else {
                const double gs = 0.57735026918962576451;
                const double s = (q == 1 || q == 2) ? gs : -gs, t = (q >= 2) ? gs : -gs;
                N[0] = 0.25 * (1 - s) * (1 - t); N[1] = 0.25 * (1 + s) * (1 - t); N[2] = 0.25 * (1 + s) * (1 + t); N[3] = 0.25 * (1 - s) * (1 + t);
                const double ds[4] = {-0.25 * (1 - t), 0.25 * (1 - t), 0.25 * (1 + t), -0.25 * (1 + t)};
                const double dt[4] = {-0.25 * (1 - s), -0.25 * (1 + s), 0.25 * (1 + s), 0.25 * (1 - s)};
                double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { a[c] += ds[k] * P[k][c]; b[c] += dt[k] * P[k][c]; }
                const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
                dg = sqrt(cx * cx + cy * cy + cz * cz);
            }
#pragma unroll
            for (int c = 0; c < 3; c++) { x[c] = 0.0; for (int k = 0; k < FN; k++) x[c] += N[k] * P[k][c]; }
            if (xq_out) { for (int c = 0; c < 3; c++) xq_out[((size_t)i * NQP + q) * 3 + c] = x[c]; }
            if (dg_out) dg_out[(size_t)i * NQP + q] = dg;
            area += dg;
            if (vload) {
                double t3[3] = {tx, ty, tz};
                if (t_qp) { for (int c = 0; c < 3; c++) t3[c] = t_qp[((size_t)i * NQP + q) * 3 + c]; }
#pragma unroll
                for (int k = 0; k < FN; k++)
#pragma unroll
                    for (int c = 0; c < 3; c++) load[k][c] += N[k] * t3[c] * dg;
#pragma unroll
                for (int c = 0; c < 3; c++) tot[c] += t3[c] * dg;
            }
        }
        if (vload) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                vnode[4 * i + k] = k < FN ? node_q[g[k < FN ? k : 0]] : -1;
#pragma unroll
                for (int c = 0; c < 3; c++) vload[(4 * (size_t)i + k) * 3 + c] = load[k][c];
            }
        }
    }
    area_part[i] = area;
    if (force_part) { force_part[3 * i] = tot[0]; force_part[3 * i + 1] = tot[1]; force_part[3 * i + 2] = tot[2]; }
}

// Deterministic scatter of the per-vertex loads: vertex entries (facet*4+k) are grouped by dof-node with a counting sort (integer
// atomics only), every node's short list is put in ascending entry order, and ONE thread per node adds its entries to f in that
// order — the sum the reference forms facet by facet in list order, bit-reproducible from run to run (no floating-point atomics).
__global__ void k_vl_count(const int* __restrict__ vnode, i64 nent, int* cnt) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nent && vnode[t] >= 0) atomicAdd(&cnt[vnode[t]], 1);
}
__global__ void k_vl_fill(const int* __restrict__ vnode, i64 nent, const int* __restrict__ ptr, int* cursor, int* __restrict__ list) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nent) return;
    const int q = vnode[t];
    if (q >= 0) list[ptr[q] + atomicAdd(&cursor[q], 1)] = (int)t;
}
__global__ void k_vertex_loads_to_f(const int* __restrict__ ptr, int* __restrict__ list, const double* __restrict__ vload, double* __restrict__ f, int nq) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int lo = ptr[q], hi = ptr[q + 1];
    if (lo == hi) return;
    for (int i = lo + 1; i < hi; i++) { int v = list[i], j = i - 1; while (j >= lo && list[j] > v) { list[j + 1] = list[j]; j--; } list[j + 1] = v; }
    double s[3] = {f[3 * (size_t)q], f[3 * (size_t)q + 1], f[3 * (size_t)q + 2]};
    for (int i = lo; i < hi; i++) {
        const double* l = vload + 3 * (size_t)list[i];
        s[0] += l[0]; s[1] += l[1]; s[2] += l[2];
    }
    f[3 * (size_t)q] = s[0]; f[3 * (size_t)q + 1] = s[1]; f[3 * (size_t)q + 2] = s[2];
}

// fixed-order sum of n values with stride `stride` (one block)
__global__ void k_sum_fixed(const double* __restrict__ in, i64 n, int stride, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) s += in[i * stride];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
}

static int surface_build(toe_ctx* ctx) {
    if (ctx->have_surface) return TOE_OK;
    if (ctx->dist) return toe_fail(ctx, TOE_ERR_STATE, "boundary-node selection works on the whole mesh: use an unpartitioned ctx");
    if (!ctx->have_pattern) return toe_fail(ctx, TOE_ERR_STATE, "select_nodes: call setup_problem first (needs the node-cell incidence lists)");
    DevBuf<int> qflag; CU(qflag.alloc(ctx->nq));
    CU(cudaMemsetAsync(qflag.p, 0, (size_t)ctx->nq * sizeof(int), ctx->stream));
    const i64 nfaces = ctx->ne * (ctx->npc == 4 ? 4 : 6);
    if (ctx->npc == 4) LAUNCH(ctx, k_surface_faces<4>, div_up(nfaces, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p, ctx->ne, qflag.p);
    else               LAUNCH(ctx, k_surface_faces<8>, div_up(nfaces, 128), 128, 0, (const int*)ctx->inc_ptr.p, (const int*)ctx->inc.p, (const int*)ctx->cq.p, ctx->ne, qflag.p);
    CU(ctx->surf_flag.alloc(ctx->nn));
    LAUNCH(ctx, k_surface_node_flags, div_up(ctx->nn, 256), 256, 0, (const int*)ctx->node_q.p, (const int*)qflag.p, ctx->nn, ctx->surf_flag.p);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_surface = true;
    return TOE_OK;
}

// mode 0 / 1 / 2 = all surface nodes / plane / circle; nodes_out may be null (count only); ascending 1-based ids
int select_nodes(toe_ctx* ctx, int mode, const double* point, const double* normal, double radius, double tol, int64_t* nodes_out, int64_t* count_out) {
    TRY(surface_build(ctx));
    double n[3] = {0, 0, 1}, p[3] = {0, 0, 0};
    if (mode >= 1) {
        if (!point || !normal) return toe_fail(ctx, TOE_ERR_ARG, "select_nodes: point and normal are required");
        double len = sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
        if (!(len > 0.0)) return toe_fail(ctx, TOE_ERR_ARG, "select_nodes: zero normal vector");
        for (int k = 0; k < 3; k++) { n[k] = normal[k] / len; p[k] = point[k]; }       // unit_normal = normal / norm(normal), :164
    }
    const i64 nn = ctx->nn;
    DevBuf<int> flag; CU(flag.alloc(nn + 1));
    LAUNCH(ctx, k_select, div_up(nn, 256), 256, 0, (const int*)ctx->surf_flag.p, (const double*)ctx->xyz.p, nn, mode, p[0], p[1], p[2], n[0], n[1], n[2], tol, radius, flag.p);
    DevBuf<int> pos; CU(pos.alloc(nn + 1));
    i64 cnt = 0;
    TRY(scan_exclusive_i32(ctx, flag.p, pos.p, nn, &cnt));
    if (count_out) *count_out = cnt;
    if (nodes_out && cnt > 0) {
        DevBuf<int64_t> ids; CU(ids.alloc(cnt));
        LAUNCH(ctx, k_compact_ids, div_up(nn, 256), 256, 0, (const int*)flag.p, (const int*)pos.p, nn, ids.p);
        CU(cudaMemcpyAsync(nodes_out, ids.p, cnt * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return TOE_OK;
}

int boundary_facets(toe_ctx* ctx, const int64_t* nodes, i64 nnodes, int64_t* facets_out, i64 capacity, int64_t* count_out) {
    if (!ctx->have_mesh) return toe_fail(ctx, TOE_ERR_STATE, "get_boundary_facets: no mesh set");
    if (ctx->dist) return toe_fail(ctx, TOE_ERR_STATE, "get_boundary_facets works on the whole mesh: use an unpartitioned ctx");
    if (nnodes < 0 || (nnodes > 0 && !nodes)) return toe_fail(ctx, TOE_ERR_ARG, "get_boundary_facets: bad node list");
    TRY(ensure_vectors(ctx));
    const i64 nn = ctx->nn, ne = ctx->ne;
    DevBuf<unsigned char> nflag; CU(nflag.alloc(nn));
    CU(cudaMemsetAsync(nflag.p, 0, nn, ctx->stream));
    CU(cudaMemsetAsync(ctx->errflag.p + 2, 0, sizeof(int), ctx->stream));
    if (nnodes > 0) {
        DevBuf<int64_t> d; CU(d.alloc(nnodes));
        CU(cudaMemcpyAsync(d.p, nodes, nnodes * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, k_flag_nodes, div_up(nnodes, 256), 256, 0, (const int64_t*)d.p, nnodes, nn, nflag.p, ctx->errflag.p);
        int e = 0;
        CU(cudaMemcpyAsync(&e, ctx->errflag.p + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (e) return toe_fail(ctx, TOE_ERR_ARG, "get_boundary_facets: node id outside 1..%lld", (long long)nn);
    }
    const i64 total = ne * (ctx->npc == 4 ? 4 : 6);
    DevBuf<int> flag, pos; CU(flag.alloc(total + 1)); CU(pos.alloc(total + 1));
    if (ctx->npc == 4) LAUNCH(ctx, k_facet_flags<4>, div_up(total, 256), 256, 0, (const int*)ctx->conn0.p, ne, (const unsigned char*)nflag.p, flag.p);
    else               LAUNCH(ctx, k_facet_flags<8>, div_up(total, 256), 256, 0, (const int*)ctx->conn0.p, ne, (const unsigned char*)nflag.p, flag.p);
    i64 cnt = 0;
    TRY(scan_exclusive_i32(ctx, flag.p, pos.p, total, &cnt));
    if (count_out) *count_out = cnt;
    if (facets_out && cnt > 0) {
        if (capacity < cnt) return toe_fail(ctx, TOE_ERR_ARG, "get_boundary_facets: %lld facets found, room for %lld", (long long)cnt, (long long)capacity);
        DevBuf<int64_t> out; CU(out.alloc(2 * cnt));
        if (ctx->npc == 4) LAUNCH(ctx, k_compact_facets<4>, div_up(total, 256), 256, 0, (const int*)flag.p, (const int*)pos.p, total, out.p);
        else               LAUNCH(ctx, k_compact_facets<8>, div_up(total, 256), 256, 0, (const int*)flag.p, (const int*)pos.p, total, out.p);
        CU(cudaMemcpyAsync(facets_out, out.p, 2 * cnt * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return TOE_OK;
}

// quadrature data of the facets (xq_out / dg_out may be null), optional load: add_load != 0 adds ∫ N·t dΓ to f with the per-point
// tractions t_qp (3 per quadrature point, facet-major) or, if t_qp is null, the uniform vector t_uniform
int facet_integrals(toe_ctx* ctx, const int64_t* facets, i64 nf, double* xq_out, double* dg_out, int add_load, const double* t_qp, const double* t_uniform,
                    double* area_out, double* total_force_out) {
    if (!ctx->have_mesh) return toe_fail(ctx, TOE_ERR_STATE, "surface traction: no mesh set");
    if (ctx->dist) return toe_fail(ctx, TOE_ERR_STATE, "surface traction works on the whole mesh: use an unpartitioned ctx (or toe_set_rhs)");
    if (add_load && !ctx->have_dofs) return toe_fail(ctx, TOE_ERR_STATE, "apply_surface_traction!: DOFs not built");
    if (nf < 0 || (nf > 0 && !facets)) return toe_fail(ctx, TOE_ERR_ARG, "surface traction: bad facet list");
    if (add_load && !t_qp && !t_uniform) return toe_fail(ctx, TOE_ERR_ARG, "surface traction: no traction given");
    TRY(ensure_vectors(ctx));
    if (area_out) *area_out = 0.0;
    if (total_force_out) total_force_out[0] = total_force_out[1] = total_force_out[2] = 0.0;
    if (nf == 0) return TOE_OK;
    const int nqp = ctx->npc == 4 ? 3 : 4;
    StageTimer T(ctx, &ctx->tm.loads);
    DevBuf<int64_t> d; CU(d.alloc(2 * nf));
    CU(cudaMemcpyAsync(d.p, facets, 2 * nf * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<double> xq, dg, tq, apart, fpart;
    if (xq_out) CU(xq.alloc((size_t)3 * nqp * nf));
    if (dg_out) CU(dg.alloc((size_t)nqp * nf));
    if (add_load && t_qp) { CU(tq.alloc((size_t)3 * nqp * nf)); CU(cudaMemcpyAsync(tq.p, t_qp, (size_t)3 * nqp * nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)); }
    CU(apart.alloc(nf + 4)); CU(fpart.alloc(3 * nf + 4));
    DevBuf<double> vload; DevBuf<int> vnode;
    if (add_load) {
        if (4 * nf > 2147483647LL) return toe_fail(ctx, TOE_ERR_ARG, "surface traction: too many facets (%lld)", (long long)nf);
        CU(vload.alloc((size_t)12 * nf)); CU(vnode.alloc((size_t)4 * nf));
    }
    CU(cudaMemsetAsync(ctx->errflag.p + 2, 0, sizeof(int), ctx->stream));
    const double tu[3] = {t_uniform ? t_uniform[0] : 0.0, t_uniform ? t_uniform[1] : 0.0, t_uniform ? t_uniform[2] : 0.0};
#define FACET_ARGS (const int64_t*)d.p, nf, (const int*)ctx->conn0.p, (const double*)ctx->xyz.p, ctx->ne, (const int*)ctx->node_q.p, xq_out ? xq.p : (double*)nullptr, \
        dg_out ? dg.p : (double*)nullptr, (add_load && t_qp) ? (const double*)tq.p : (const double*)nullptr, tu[0], tu[1], tu[2], add_load ? vload.p : (double*)nullptr, add_load ? vnode.p : (int*)nullptr, \
        apart.p, fpart.p, ctx->errflag.p
    if (ctx->npc == 4) LAUNCH(ctx, k_facets<4>, div_up(nf, 128), 128, 0, FACET_ARGS);
    else               LAUNCH(ctx, k_facets<8>, div_up(nf, 128), 128, 0, FACET_ARGS);
#undef FACET_ARGS
    DevBuf<int> vptr, vcur, vlist;
    if (add_load) {
        const i64 nent = 4 * nf;
        const int nq = ctx->nq;
        CU(vptr.alloc((size_t)nq + 1)); CU(vcur.alloc((size_t)nq + 1)); CU(vlist.alloc((size_t)nent));
        CU(cudaMemsetAsync(vptr.p, 0, ((size_t)nq + 1) * sizeof(int), ctx->stream));
        CU(cudaMemsetAsync(vcur.p, 0, ((size_t)nq + 1) * sizeof(int), ctx->stream));
        LAUNCH(ctx, k_vl_count, div_up(nent, 256), 256, 0, (const int*)vnode.p, nent, vptr.p);
        i64 tot_ent = 0;
        TRY(scan_exclusive_i32(ctx, vptr.p, vptr.p, nq, &tot_ent));
        LAUNCH(ctx, k_vl_fill, div_up(nent, 256), 256, 0, (const int*)vnode.p, nent, (const int*)vptr.p, vcur.p, vlist.p);
        LAUNCH(ctx, k_vertex_loads_to_f, div_up(nq, 128), 128, 0, (const int*)vptr.p, vlist.p, (const double*)vload.p, ctx->f.p, nq);
    }
    LAUNCH(ctx, k_sum_fixed, 1, 256, 0, (const double*)apart.p, nf, 1, apart.p + nf);
    for (int c = 0; c < 3; c++) LAUNCH(ctx, k_sum_fixed, 1, 256, 0, (const double*)(fpart.p + c), nf, 3, fpart.p + 3 * nf + c);
    double h[4] = {0, 0, 0, 0}; int e = 0;
    CU(cudaMemcpyAsync(&h[0], apart.p + nf, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&h[1], fpart.p + 3 * nf, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&e, ctx->errflag.p + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (xq_out) CU(cudaMemcpyAsync(xq_out, xq.p, (size_t)3 * nqp * nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (dg_out) CU(cudaMemcpyAsync(dg_out, dg.p, (size_t)nqp * nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(T.finish());
    if (e) return toe_fail(ctx, TOE_ERR_ARG, "surface traction: facet (cell, local face) out of range");
    if (area_out) *area_out = h[0];
    if (total_force_out) { total_force_out[0] = h[1]; total_force_out[1] = h[2]; total_force_out[2] = h[3]; }
    if (add_load) ctx->have_solution = false;
    return TOE_OK;
}
