// dist.cu — multi-GPU layer: one ctx per GPU / process, element-based domain decomposition.
//
// The reference is single-process, single-threaded; this layer is new.  Every rank receives the same
// global mesh, numbers the DOFs globally (so `u` comes back in the reference's Ferrite order), splits the cells by
// recursive coordinate bisection of their centroids — computed redundantly and deterministically on every GPU, so
// no partition data is exchanged — and keeps its own cells plus the nodes they touch.  K is sub-assembled per part
// (assembly needs no communication); vectors are stored "interface-consistent" (every rank holding a node holds its
// full value).  One operator application = local product + interface sum: pack → ncclSend/ncclRecv with every
// neighbouring part → unpack, contributions added in ascending rank order so that all copies of an interface value
// are bit-identical.  Dot products are owner-masked and closed by ncclAllReduce.  NCCL is dlopen'ed (libnccl.so.2).
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>
#include <cstdlib>

struct NcclApi {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(std::string& err) {
    if (g_nccl.h) return TOE_OK;
#ifdef TOE_EMU
    // tests/cuda_emu: rank threads of one process rendezvous through the stand-in functions (test infrastructure only)
    (void)err;
    g_nccl.GetUniqueId = &ncclGetUniqueId; g_nccl.CommInitRank = &ncclCommInitRank; g_nccl.CommDestroy = &ncclCommDestroy;
    g_nccl.AllReduce = &ncclAllReduce; g_nccl.AllGather = &ncclAllGather; g_nccl.Send = &ncclSend; g_nccl.Recv = &ncclRecv; g_nccl.GroupStart = &ncclGroupStart;
    g_nccl.GroupEnd = &ncclGroupEnd; g_nccl.GetErrorString = &ncclGetErrorString;
    g_nccl.h = (void*)&g_nccl;
    return TOE_OK;
#endif
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return TOE_ERR_COMM; }
#define SYM(f) g_nccl.f = (decltype(g_nccl.f))dlsym(h, "nccl" #f); if (!g_nccl.f) { err = "libnccl lacks nccl" #f; return TOE_ERR_COMM; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(AllGather) SYM(Send) SYM(Recv) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    g_nccl.h = h;
    return TOE_OK;
}

#define NC(call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) \
    return toe_fail(ctx, TOE_ERR_COMM, "NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r)); } while (0)

struct DistState {
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    // global problem
    i64 nn_g = 0, ne_g = 0;
    int nq_g = 0;
    DevBuf<int> node_q_g;        // nn_g: global dof-node id of each mesh node (or -1)
    DevBuf<int> part;            // ne_g: part id of every global cell
    DevBuf<int> eloc2glob;       // ne_local: global cell id of each local cell (ascending)
    DevBuf<int> loc2glob;        // nq_local: global dof-node id of each local dof-node (ascending)
    DevBuf<int> glob2loc;        // nq_g
    DevBuf<unsigned char> owned; // nq_local
    // interface exchange
    std::vector<int> nbr;                 // neighbouring ranks (ascending)
    std::vector<int> nbr_count;           // shared nodes with each neighbour
    std::vector<int> nbr_off;             // offset (in nodes) of each neighbour's segment in the send/recv buffers
    int n_shared_total = 0;               // Σ nbr_count
    int n_if = 0;                         // local interface nodes
    DevBuf<int> send_nodes;               // n_shared_total: local node id per send slot
    DevBuf<double> sendbuf, recvbuf;      // 2 x 3 * n_shared_total: staging alternates between two halves from one exchange to the next
    int xpar = 0;
    DevBuf<int> if_node, if_ptr, if_src;  // unpack CSR: for interface node i, sources in ascending rank order; src = -1 → own value, else recv slot
    DevBuf<double> gvec;                  // global-length scratch for gathers
    // peer-memory exchange (CUDA IPC over NVLink/NVSwitch): one kernel does pack + interface sum + scalar allreduce
    bool p2p_ok = false;
    char* mbox = nullptr;                 // this rank's mailbox: flags | scalars | receive areas (written by the peers)
    size_t mbox_bytes = 0;
    i64 stride3 = 0;                      // doubles per (parity, source rank) receive area
    std::vector<char*> peer_mbox;         // [nranks] device pointers (own entry = mbox)
    DevBuf<char*> peer_mbox_dev;
    DevBuf<int> seg_rank, seg_off, seg_cnt, if_srcx;
    u64 xseq = 0;                         // exchange sequence number (identical on all ranks)
    DevBuf<u64> gbar;                     // grid-barrier counter of the exchange kernel (monotonic)
    u64 xlaunch = 0;                      // launches of the exchange kernel so far (local)
    // all-gather exchange (default transport): ONE collective per operator application carries every rank's packed
    // interface values plus its two partial scalars; each rank picks its neighbours' segments out of the gathered buffer
    bool ag_ok = false;
    i64 ag_stride = 0;                    // doubles per rank in the gathered buffer: 3 * max_r(n_shared_total) + 2 (even → 16-byte slices)
    DevBuf<double> ag_send, ag_recv;      // 2 x stride, 2 x nranks * stride (halves alternate like the p2p staging)
    DevBuf<int> if_src_ag;                // unpack sources in gathered-buffer coordinates (node slots; -1 = own value)
};

static int mailbox_setup(toe_ctx* ctx, DistState* d, const std::vector<int>& if_src_host);
static int allgather_setup(toe_ctx* ctx, DistState* d, const std::vector<int>& if_src_host);
static int exchange_allgather(toe_ctx* ctx, double* y, double* scal, int count);

bool dist_active(toe_ctx* ctx) { return ctx->dist != nullptr; }

// Transport of the per-iteration exchange (read at every set-up):
//   allgather   ONE ncclAllGather carries every rank's packed interface values and its partial scalars: one collective and 3 launches
//               per exchange.  Default below 8 ranks (10M tets: 3.11 s per solve at N=2 against 3.24 s for send/recv + allreduce and
//               3.13 s for the peer-memory kernel; 1.836 s against 1.853 s for the peer-memory kernel at N=4).
//   p2p         fused peer-memory kernel over CUDA IPC mailboxes: neighbour-only traffic, one launch.  Default from 8 ranks on, where
//               the all-gather moves 8 slices to everybody (1 183 ms per 10M-tet step at N=8 against 1 304 ms for the all-gather and
//               1 440 ms for send/recv + allreduce).  If the mailboxes cannot be set up (no peer access), the all-gather is used.
//   sendrecv    grouped ncclSend/ncclRecv per neighbour + ncclAllReduce of the scalars — the literal north-star form.
// TOE_DIST_XCHG=allgather|p2p|sendrecv forces one (TOE_DIST_P2P=1 = p2p).  All three do the same arithmetic in the same order
// (bit-identical iterates).  The transient CG faults of round 1 were NOT a transport problem: they came from the SpMV's pipeline protocol
// (profiles/r2_dist_diagnosis.md); 48 of 48 solves are clean on either NCCL transport, 38 of 38 on the peer-memory kernel.
enum XchgMode { XCHG_ALLGATHER = 0, XCHG_SENDRECV = 1, XCHG_P2P = 2 };
static XchgMode xchg_mode(int nranks) {
    const char* m = getenv("TOE_DIST_XCHG");
    if (m && strcmp(m, "sendrecv") == 0) return XCHG_SENDRECV;
    if (m && strcmp(m, "allgather") == 0) return XCHG_ALLGATHER;
    if ((m && strcmp(m, "p2p") == 0) || (!m && getenv("TOE_DIST_P2P"))) return XCHG_P2P;
    return nranks >= 8 ? XCHG_P2P : XCHG_ALLGATHER;
}

static void mailbox_close_peers(DistState* d) {
    for (int r = 0; r < (int)d->peer_mbox.size(); r++)
        if (r != d->rank && d->peer_mbox[r]) cudaIpcCloseMemHandle(d->peer_mbox[r]);
    d->peer_mbox.clear();
    d->p2p_ok = false;
}
static void mailbox_free_own(DistState* d) {
    if (d->mbox) cudaFree(d->mbox);
    d->mbox = nullptr; d->mbox_bytes = 0;
}
static void mailbox_release(DistState* d) { mailbox_close_peers(d); mailbox_free_own(d); }

void dist_destroy(toe_ctx* ctx) {
    if (!ctx->dist) return;
    cudaStreamSynchronize(ctx->stream);
    mailbox_release(ctx->dist);
    if (ctx->dist->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->dist->comm);
    delete ctx->dist;
    ctx->dist = nullptr; ctx->owned = nullptr; ctx->glob2loc = nullptr; ctx->n_global = 0;
}

int dist_comm_unique_id(char id_out[128], std::string& err) {
    TRY(nccl_load(err));
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return TOE_ERR_COMM; }
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id_out, &id, 128);
    return TOE_OK;
}

int dist_comm_init(toe_ctx* ctx, int nranks, int rank, const char id[128]) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return toe_fail(ctx, TOE_ERR_ARG, "toe_comm_init: bad rank %d of %d", rank, nranks);
    if (nranks & (nranks - 1)) return toe_fail(ctx, TOE_ERR_ARG, "toe_comm_init: the bisection partitioner needs a power-of-two number of ranks, got %d", nranks);
    if (nranks > 32) return toe_fail(ctx, TOE_ERR_ARG, "toe_comm_init: at most 32 ranks");
    std::string err;
    if (nccl_load(err) != TOE_OK) return toe_fail(ctx, TOE_ERR_COMM, "%s", err.c_str());
    dist_destroy(ctx);
    DistState* d = new DistState();
    d->nranks = nranks; d->rank = rank;
    ncclUniqueId uid; memcpy(&uid, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&d->comm, nranks, uid, rank);
    if (r != ncclSuccess) { delete d; return toe_fail(ctx, TOE_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    ctx->dist = d;
    return TOE_OK;
}

int dist_allreduce(toe_ctx* ctx, double* dev_vals, int count) {
    if (!ctx->dist || ctx->dist->nranks == 1) return TOE_OK;
    NC(g_nccl.AllReduce(dev_vals, dev_vals, (size_t)count, ncclDouble, ncclSum, ctx->dist->comm, ctx->stream));
    return TOE_OK;
}

// Rendezvous of the ranks' streams: one 8-byte allreduce, then the host waits for it.  Set-up work differs per rank (host-side map
// building, partition sizes), so the ranks reach the first exchange of a step up to ~100 ms apart; meeting here makes the ranks'
// stage timers start together (two ~20 µs collectives per step).  Correctness does not depend on it (TOE_DIST_NO_ALIGN=1 runs clean;
// the faults it was once added against were the SpMV pipeline's, profiles/r2_dist_diagnosis.md).
int dist_align(toe_ctx* ctx) {
    DistState* d = ctx->dist;
    if (!d || d->nranks == 1) return TOE_OK;
    static const bool off = getenv("TOE_DIST_NO_ALIGN") != nullptr;      // A/B switch for tools/dist_diag.py
    if (off) return TOE_OK;
    TRY(ensure_vectors(ctx));
    double* slot = ctx->partials.p + PARTIALS_ALIGN_SLOT;
    CU(cudaMemsetAsync(slot, 0, sizeof(double), ctx->stream));
    NC(g_nccl.AllReduce(slot, slot, 1, ncclDouble, ncclSum, d->comm, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// recursive coordinate bisection by exact radix selection on (quantised centroid coordinate, cell id) keys
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 enc_double(double x) {      // order-preserving map double -> u64
    u64 b = (u64)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ inline double dec_double(u64 k) {
    u64 b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    double x; memcpy(&x, &b, 8); return x;
}

__global__ void k_bbox(const double* __restrict__ xyz, i64 nn, u64* mn, u64* mx) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 lo[3] = {~0ULL, ~0ULL, ~0ULL}, hi[3] = {0, 0, 0};
    if (g < nn) for (int c = 0; c < 3; c++) { u64 k = enc_double(xyz[3 * g + c]); lo[c] = k; hi[c] = k; }
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
            u64 a = __shfl_xor_sync(0xffffffffu, lo[c], o), b = __shfl_xor_sync(0xffffffffu, hi[c], o);
            lo[c] = a < lo[c] ? a : lo[c]; hi[c] = b > hi[c] ? b : hi[c];
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&mn[c], lo[c]); atomicMax(&mx[c], hi[c]); }
    }
}

struct SplitParams {          // per current part (max 16 parts before the last split of 32)
    int axis[16];
    double lo[16], scale[16];
    u64 prefix[16];           // key prefix selected so far
    u64 thresh[16];
};

__device__ __forceinline__ u64 cell_key(const int* __restrict__ conn0, const double* __restrict__ xyz, int npc, i64 e, int axis, double lo, double scale) {
    double s = 0.0;
    for (int a = 0; a < npc; a++) s += xyz[3 * (size_t)conn0[e * npc + a] + axis];
    s /= (double)npc;
    double q = (s - lo) * scale;
    if (q < 0) q = 0; if (q > 4294967295.0) q = 4294967295.0;
    return ((u64)(unsigned int)q << 32) | (u64)(unsigned int)e;
}

// histogram of the 16-bit digit `pass` (0 = most significant) among cells whose higher digits equal prefix[part]
__global__ void k_split_hist(const int* __restrict__ conn0, const double* __restrict__ xyz, int npc, i64 ne, const int* __restrict__ part,
                             SplitParams sp, int pass, unsigned int* __restrict__ hist) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int p = part[e];
    u64 key = cell_key(conn0, xyz, npc, e, sp.axis[p], sp.lo[p], sp.scale[p]);
    int shift = 16 * (3 - pass);
    if (pass > 0 && (key >> (shift + 16)) != sp.prefix[p]) return;
    unsigned int digit = (unsigned int)(key >> shift) & 0xffffu;
    atomicAdd(&hist[(size_t)p * 65536 + digit], 1u);
}

__global__ void k_split_apply(const int* __restrict__ conn0, const double* __restrict__ xyz, int npc, i64 ne, int* __restrict__ part, SplitParams sp) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int p = part[e];
    u64 key = cell_key(conn0, xyz, npc, e, sp.axis[p], sp.lo[p], sp.scale[p]);
    part[e] = 2 * p + (key >= sp.thresh[p] ? 1 : 0);
}

static int partition_rcb(toe_ctx* ctx, DistState* d) {
    i64 ne = d->ne_g;
    int npc = ctx->npc;
    CU(d->part.alloc(ne));
    CU(cudaMemsetAsync(d->part.p, 0, ne * sizeof(int), ctx->stream));
    if (d->nranks == 1) return TOE_OK;
    DevBuf<u64> bb; CU(bb.alloc(6));
    u64 init[6] = {~0ULL, ~0ULL, ~0ULL, 0, 0, 0};
    CU(cudaMemcpyAsync(bb.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_bbox, div_up(d->nn_g, 256), 256, 0, (const double*)ctx->xyz.p, d->nn_g, bb.p, bb.p + 3);
    u64 hb[6];
    CU(cudaMemcpyAsync(hb, bb.p, sizeof hb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    double glo[3], ghi[3];
    for (int c = 0; c < 3; c++) { glo[c] = dec_double(hb[c]); ghi[c] = dec_double(hb[3 + c]); }
    std::vector<double> ext(3 * 32);
    std::vector<i64> count(32, 0);
    for (int c = 0; c < 3; c++) ext[c] = ghi[c] - glo[c];
    count[0] = ne;
    DevBuf<unsigned int> hist; CU(hist.alloc((size_t)16 * 65536));
    std::vector<unsigned int> hh((size_t)16 * 65536);
    for (int nparts = 1; nparts < d->nranks; nparts *= 2) {
        SplitParams sp;
        memset(&sp, 0, sizeof sp);
        std::vector<i64> target(nparts), rem(nparts);
        for (int p = 0; p < nparts; p++) {
            int ax = 0;
            for (int c = 1; c < 3; c++) if (ext[3 * p + c] > ext[3 * p + ax]) ax = c;
            sp.axis[p] = ax; sp.lo[p] = glo[ax];
            sp.scale[p] = (ghi[ax] > glo[ax]) ? 4294967295.0 / (ghi[ax] - glo[ax]) : 0.0;
            target[p] = count[p] / 2;            // cells that go to the lower child
            rem[p] = target[p];
        }
        for (int pass = 0; pass < 4; pass++) {
            CU(cudaMemsetAsync(hist.p, 0, (size_t)nparts * 65536 * sizeof(unsigned int), ctx->stream));
            LAUNCH(ctx, k_split_hist, div_up(ne, 256), 256, 0, (const int*)ctx->conn0.p, (const double*)ctx->xyz.p, npc, ne, (const int*)d->part.p, sp, pass, hist.p);
            CU(cudaMemcpyAsync(hh.data(), hist.p, (size_t)nparts * 65536 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            for (int p = 0; p < nparts; p++) {
                const unsigned int* h = hh.data() + (size_t)p * 65536;
                i64 cum = 0; int dg = 0;
                for (; dg < 65535; dg++) { if (cum + h[dg] > rem[p]) break; cum += h[dg]; }
                rem[p] -= cum;
                sp.prefix[p] = (sp.prefix[p] << 16) | (u64)dg;
            }
        }
        for (int p = 0; p < nparts; p++) sp.thresh[p] = sp.prefix[p];     // key of the cell with rank target[p]
        LAUNCH(ctx, k_split_apply, div_up(ne, 256), 256, 0, (const int*)ctx->conn0.p, (const double*)ctx->xyz.p, npc, ne, d->part.p, sp);
        std::vector<double> next(3 * 32);
        std::vector<i64> ncount(32, 0);
        for (int p = nparts - 1; p >= 0; p--) {
            for (int c = 0; c < 3; c++) { double v = ext[3 * p + c] * (c == sp.axis[p] ? 0.5 : 1.0); next[3 * (2 * p) + c] = v; next[3 * (2 * p + 1) + c] = v; }
            ncount[2 * p] = target[p]; ncount[2 * p + 1] = count[p] - target[p];
        }
        ext = next; count = ncount;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// local sub-mesh extraction
// ---------------------------------------------------------------------------------------------------------
__global__ void k_flag_eq(const int* __restrict__ a, int v, int* __restrict__ flag, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = a[i] == v ? 1 : 0;
}
__global__ void k_compact_cells(const int* __restrict__ part, int rank, const int* __restrict__ pos, const int* __restrict__ cq_g, int npc,
                                int* __restrict__ eloc2glob, int* __restrict__ cq_tmp, int* __restrict__ touched, i64 ne) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne || part[e] != rank) return;
    int j = pos[e];
    eloc2glob[j] = (int)e;
    for (int a = 0; a < npc; a++) { int q = cq_g[e * npc + a]; cq_tmp[(size_t)j * npc + a] = q; touched[q] = 1; }
}
__global__ void k_rank_mask(const int* __restrict__ part, const int* __restrict__ cq_g, int npc, unsigned int* __restrict__ mask, i64 ne) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne * npc) return;
    atomicOr(&mask[cq_g[i]], 1u << part[i / npc]);
}
__global__ void k_glob2loc(const int* __restrict__ touched, const int* __restrict__ pos, int* __restrict__ glob2loc, int* __restrict__ loc2glob,
                           const double* __restrict__ xq_g, double* __restrict__ xq_l, const unsigned int* __restrict__ mask, int rank,
                           unsigned char* __restrict__ owned, unsigned int* __restrict__ mask_l, int nq_g) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_g) return;
    if (!touched[q]) { glob2loc[q] = -1; return; }
    int l = pos[q];
    glob2loc[q] = l; loc2glob[l] = q;
    xq_l[3 * (size_t)l] = xq_g[3 * (size_t)q]; xq_l[3 * (size_t)l + 1] = xq_g[3 * (size_t)q + 1]; xq_l[3 * (size_t)l + 2] = xq_g[3 * (size_t)q + 2];
    unsigned int m = mask[q];
    mask_l[l] = m;
    owned[l] = ((m & (0u - m)) == (1u << rank)) ? 1 : 0;      // lowest touching rank owns the node
}
__global__ void k_remap(const int* __restrict__ in, const int* __restrict__ map, int* __restrict__ out, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] >= 0 ? map[in[i]] : -1;
}

int dist_set_mesh(toe_ctx* ctx, i64 nn, const double* xyz, i64 ne, int npc, const int64_t* conn) {
    DistState* d = ctx->dist;
    if (!d) return toe_fail(ctx, TOE_ERR_STATE, "toe_set_mesh_distributed: call toe_comm_init first");
    ctx->owned = nullptr; ctx->glob2loc = nullptr; ctx->n_global = 0;
    TRY(mesh_upload(ctx, nn, xyz, ne, npc, conn));
    TRY(mesh_build_dofs(ctx));                 // global first-touch numbering: identical to the 1-GPU numbering
    d->nn_g = nn; d->ne_g = ne; d->nq_g = ctx->nq;
    TRY(partition_rcb(ctx, d));
    // local cells (ascending global id)
    DevBuf<int> flag; CU(flag.alloc(std::max<i64>(ne, d->nq_g) + 1));
    LAUNCH(ctx, k_flag_eq, div_up(ne, 256), 256, 0, (const int*)d->part.p, d->rank, flag.p, ne);
    i64 ne_l = 0;
    TRY(scan_exclusive_i32(ctx, flag.p, flag.p, ne, &ne_l));
    if (ne_l == 0) return toe_fail(ctx, TOE_ERR_MESH, "rank %d received no cells", d->rank);
    CU(d->eloc2glob.alloc(ne_l));
    DevBuf<int> cq_tmp, touched; CU(cq_tmp.alloc(ne_l * npc)); CU(touched.alloc(d->nq_g + 1));
    CU(cudaMemsetAsync(touched.p, 0, (d->nq_g + 1) * sizeof(int), ctx->stream));
    LAUNCH(ctx, k_compact_cells, div_up(ne, 256), 256, 0, (const int*)d->part.p, d->rank, (const int*)flag.p, (const int*)ctx->cq.p, npc,
           d->eloc2glob.p, cq_tmp.p, touched.p, ne);
    // rank masks of every global dof-node
    DevBuf<unsigned int> mask; CU(mask.alloc(d->nq_g));
    CU(cudaMemsetAsync(mask.p, 0, d->nq_g * sizeof(unsigned int), ctx->stream));
    LAUNCH(ctx, k_rank_mask, div_up(ne * npc, 256), 256, 0, (const int*)d->part.p, (const int*)ctx->cq.p, npc, mask.p, ne);
    // local dof-nodes (ascending global id)
    DevBuf<int> pos; CU(pos.alloc(d->nq_g + 1));
    i64 nq_l = 0;
    TRY(scan_exclusive_i32(ctx, touched.p, pos.p, d->nq_g, &nq_l));
    CU(d->glob2loc.alloc(d->nq_g)); CU(d->loc2glob.alloc(nq_l)); CU(d->owned.alloc(nq_l));
    DevBuf<double> xq_l; CU(xq_l.alloc(3 * nq_l));
    DevBuf<unsigned int> mask_l; CU(mask_l.alloc(nq_l));
    LAUNCH(ctx, k_glob2loc, div_up(d->nq_g, 256), 256, 0, (const int*)touched.p, (const int*)pos.p, d->glob2loc.p, d->loc2glob.p,
           (const double*)ctx->xq.p, xq_l.p, (const unsigned int*)mask.p, d->rank, d->owned.p, mask_l.p, d->nq_g);
    // swap the ctx over to the local sub-mesh
    CU(d->node_q_g.alloc(nn));
    CU(cudaMemcpyAsync(d->node_q_g.p, ctx->node_q.p, nn * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(ctx, k_remap, div_up(nn, 256), 256, 0, (const int*)d->node_q_g.p, (const int*)d->glob2loc.p, ctx->node_q.p, nn);   // node -> LOCAL dof-node
    CU(ctx->cq.alloc(ne_l * npc));
    LAUNCH(ctx, k_remap, div_up(ne_l * npc, 256), 256, 0, (const int*)cq_tmp.p, (const int*)d->glob2loc.p, ctx->cq.p, ne_l * npc);
    CU(ctx->xq.alloc(3 * nq_l));
    CU(cudaMemcpyAsync(ctx->xq.p, xq_l.p, 3 * nq_l * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    // interface maps (host-built from the local rank masks)
    std::vector<unsigned int> hm(nq_l);
    CU(cudaMemcpyAsync(hm.data(), mask_l.p, nq_l * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ne = ne_l; ctx->nq = (int)nq_l;
    ctx->owned = d->owned.p; ctx->glob2loc = d->glob2loc.p; ctx->n_global = 3 * (i64)d->nq_g;
    const unsigned int me = 1u << d->rank;
    d->nbr.clear(); d->nbr_count.clear(); d->nbr_off.clear();
    std::vector<int> cnt(d->nranks, 0);
    int n_if = 0;
    for (i64 l = 0; l < nq_l; l++) {
        unsigned int m = hm[l] & ~me;
        if (!m) continue;
        n_if++;
        for (int r = 0; r < d->nranks; r++) if (m & (1u << r)) cnt[r]++;
    }
    std::vector<int> slot_of_rank(d->nranks, -1);
    int off = 0;
    for (int r = 0; r < d->nranks; r++) if (cnt[r]) { slot_of_rank[r] = (int)d->nbr.size(); d->nbr.push_back(r); d->nbr_count.push_back(cnt[r]); d->nbr_off.push_back(off); off += cnt[r]; }
    d->n_shared_total = off; d->n_if = n_if;
    std::vector<int> send_nodes(off), fill(d->nbr.size(), 0), if_node(n_if), if_ptr(n_if + 1), if_src;
    if_src.reserve((size_t)off + n_if);
    int ii = 0;
    for (i64 l = 0; l < nq_l; l++) {
        unsigned int m = hm[l];
        if (!(m & ~me)) continue;
        if_node[ii] = (int)l; if_ptr[ii] = (int)if_src.size();
        for (int r = 0; r < d->nranks; r++) {
            if (!(m & (1u << r))) continue;
            if (r == d->rank) { if_src.push_back(-1); continue; }
            int s = slot_of_rank[r];
            int p = d->nbr_off[s] + fill[s]++;          // ascending local id = ascending global id on both sides
            send_nodes[p] = (int)l;
            if_src.push_back(p);
        }
        ii++;
    }
    if_ptr[n_if] = (int)if_src.size();
    CU(d->send_nodes.alloc(off)); CU(d->sendbuf.alloc(6 * (size_t)off)); CU(d->recvbuf.alloc(6 * (size_t)off));
    CU(d->if_node.alloc(n_if)); CU(d->if_ptr.alloc(n_if + 1)); CU(d->if_src.alloc(if_src.size()));
    if (off) CU(cudaMemcpy(d->send_nodes.p, send_nodes.data(), off * sizeof(int), cudaMemcpyHostToDevice));
    if (n_if) {
        CU(cudaMemcpy(d->if_node.p, if_node.data(), n_if * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d->if_src.p, if_src.data(), if_src.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(d->if_ptr.p, if_ptr.data(), (n_if + 1) * sizeof(int), cudaMemcpyHostToDevice));
    // the uploads above ran on the NULL stream from pageable memory (the call returns once the data is staged, the DMA may
    // still be in flight) and ctx->stream is non-blocking: order them before anything the ctx stream does with the maps
    CU(cudaDeviceSynchronize());
    ctx->have_dofs = true; ctx->have_pattern = ctx->have_contrib = ctx->have_K = false; ctx->have_tiles = false;
    TRY(ensure_vectors(ctx));                       // the exchange kernel reads the PCG `done` flag
    CU(cudaMemsetAsync(ctx->cgs.p, 0, sizeof(CGScalars), ctx->stream));
    TRY(mailbox_setup(ctx, d, if_src));
    TRY(allgather_setup(ctx, d, if_src));
    return dist_align(ctx);
}

// ---------------------------------------------------------------------------------------------------------
// interface sum
// ---------------------------------------------------------------------------------------------------------
__global__ void k_pack(const int* __restrict__ send_nodes, const double* __restrict__ y, double* __restrict__ buf, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    int s = i / 3, c = i - 3 * s;
    buf[i] = y[3 * (size_t)send_nodes[s] + c];
}
__global__ void k_unpack_sum(const int* __restrict__ if_node, const int* __restrict__ if_ptr, const int* __restrict__ if_src,
                             const double* __restrict__ recv, double* __restrict__ y, int n_if) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n_if) return;
    int k = i / 3, c = i - 3 * k;
    size_t dof = 3 * (size_t)if_node[k] + c;
    double own = y[dof], s = 0.0;
    for (int j = if_ptr[k]; j < if_ptr[k + 1]; j++) { int src = if_src[j]; s += src < 0 ? own : recv[3 * (size_t)src + c]; }   // ascending rank order
    y[dof] = s;
}

int dist_post_spmv(toe_ctx* ctx, double* y) {
    DistState* d = ctx->dist;
    if (!d || d->nranks == 1 || !y) return TOE_OK;
    if (d->ag_ok) return exchange_allgather(ctx, y, nullptr, 0);      // a collective: also ranks without an interface take part
    if (d->n_shared_total == 0) return TOE_OK;
    int n = d->n_shared_total;
    // the staging buffers alternate between two halves: an exchange never reuses the buffers of the previous one, whatever
    // the transport's completion semantics for the peer's side of a send/receive are
    d->xpar ^= 1;
    double* sb = d->sendbuf.p + (size_t)d->xpar * 3 * n;
    double* rb = d->recvbuf.p + (size_t)d->xpar * 3 * n;
    LAUNCH(ctx, k_pack, div_up(3 * (i64)n, 256), 256, 0, (const int*)d->send_nodes.p, (const double*)y, sb, n);
    NC(g_nccl.GroupStart());
    for (size_t k = 0; k < d->nbr.size(); k++) {
        NC(g_nccl.Send(sb + 3 * (size_t)d->nbr_off[k], 3 * (size_t)d->nbr_count[k], ncclDouble, d->nbr[k], d->comm, ctx->stream));
        NC(g_nccl.Recv(rb + 3 * (size_t)d->nbr_off[k], 3 * (size_t)d->nbr_count[k], ncclDouble, d->nbr[k], d->comm, ctx->stream));
    }
    NC(g_nccl.GroupEnd());
    LAUNCH(ctx, k_unpack_sum, div_up(3 * (i64)d->n_if, 256), 256, 0, (const int*)d->if_node.p, (const int*)d->if_ptr.p, (const int*)d->if_src.p,
           (const double*)rb, y, d->n_if);
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Fused exchange over NVLink peer memory (no NCCL on the iteration path).  Every rank owns a mailbox that its peers
// map through CUDA IPC:   [ flags u64[32] | scalars double[2][32][2] | receive areas double[2][nranks][stride3] ]
// One single-CTA kernel per operator application:
//   1. stores this rank's interface values of y straight into each neighbour's receive area (peer stores over NVLink),
//      and its two partial dot products into every rank's scalar slots;
//   2. __threadfence_system(), then publishes the sequence number in every peer's flag word;
//   3. spins (with a time-out) until every peer's flag has reached the sequence number;
//   4. sums the scalars in rank order (bit-identical on all ranks) and the interface contributions in ascending rank
//      order (own value included) — the same arithmetic as the NCCL path, so results do not depend on the transport.
// Buffers alternate with the parity of the sequence number: a rank can run at most one exchange ahead of a peer.
// ---------------------------------------------------------------------------------------------------------
static const size_t MB_OFF_FLAG = 0, MB_OFF_SCAL = 256, MB_OFF_RECV = 256 + 2 * 32 * 2 * sizeof(double);

#ifndef TOE_EMU
static const int XCHG_CTAS = 32, XCHG_THREADS = 256;      // all co-resident (the stream is otherwise idle while it runs)
#else
static const int XCHG_CTAS = 4, XCHG_THREADS = 64;        // tests/cuda_emu: same protocol, fewer fibers
#endif

// software grid barrier on a monotonically increasing counter (every CTA adds 1 per barrier).  Bounded: if the other CTAs do
// not arrive within ≈3 s (they are not co-resident, or one of them is stuck) the error flag is raised and the kernel moves
// on — a wrong exchange is reported by dist_check_exchange, a hung GPU could not be.
__device__ __forceinline__ void xchg_grid_barrier(u64* ctr, u64 target, int* err_flag, int* done_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1ULL);
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile u64*>(ctr) < target) {
#ifdef TOE_EMU
            emu::yield();
#endif
            if (clock64() - t0 > 6000000000LL) { atomicExch(err_flag, 1); if (done_flag) atomicExch(done_flag, 1); break; }
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(XCHG_THREADS) k_xchg(char* const* __restrict__ peers, int nranks, int me, u64 seq, i64 stride3,
                                                      const int* __restrict__ seg_rank, const int* __restrict__ seg_off, const int* __restrict__ seg_cnt, int nseg,
                                                      int n_send, const int* __restrict__ send_nodes, const int* __restrict__ if_node,
                                                      const int* __restrict__ if_ptr, const int* __restrict__ if_srcx, int n_if, double* __restrict__ y,
                                                      double* scal, int nscal, int* done_flag, int* err_flag, u64* gbar, u64 launch_index) {
    const u64 base = launch_index * 2ULL * gridDim.x;
    if (done_flag && *done_flag) {                         // keep the barrier counter in step with the launch index
        if (threadIdx.x == 0) atomicAdd(gbar, 2ULL);
        return;
    }
    const int tid = threadIdx.x, gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const int par = (int)(seq & 1);
    char* mine = peers[me];
    // 1. halo push: this rank's interface values straight into each neighbour's receive area (peer stores over NVLink)
    for (int i = gtid; i < 3 * n_send; i += gsz) {
        int p = i / 3, c = i - 3 * p;
        int k = 0;
        while (k + 1 < nseg && p >= seg_off[k + 1]) k++;
        double* dst = reinterpret_cast<double*>(peers[seg_rank[k]] + MB_OFF_RECV) + ((size_t)par * nranks + me) * stride3;
        dst[3 * (size_t)(p - seg_off[k]) + c] = y[3 * (size_t)send_nodes[p] + c];
    }
    if (blockIdx.x == 0 && tid < nranks && nscal > 0) {
        double* dst = reinterpret_cast<double*>(peers[tid] + MB_OFF_SCAL) + ((size_t)par * 32 + me) * 2;
        dst[0] = scal[0]; dst[1] = nscal > 1 ? scal[1] : 0.0;
    }
    __threadfence_system();
    xchg_grid_barrier(gbar, base + gridDim.x, err_flag, done_flag);
    // 2. publish the sequence number in every peer's flag word, 3. wait for every peer's
    if (blockIdx.x == 0) {
        if (tid < nranks && tid != me) {
            // release: the grid barrier made every CTA's peer stores (each followed by a system fence) visible to this
            // thread; this fence + st.release.sys orders them before the flag for any observer of the flag
            __threadfence_system();
            u64* out = reinterpret_cast<u64*>(peers[tid] + MB_OFF_FLAG) + me;
#ifndef TOE_EMU
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(out), "l"(seq) : "memory");
#else
            *reinterpret_cast<volatile u64*>(out) = seq;
#endif
            const u64* in = reinterpret_cast<const u64*>(mine + MB_OFF_FLAG) + tid;
            long long t0 = clock64();
            u64 v;
            while (true) {
#ifndef TOE_EMU
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(in) : "memory");
#else
                v = *reinterpret_cast<const volatile u64*>(in); emu::yield();
#endif
                if (v >= seq) break;
                if (clock64() - t0 > 6000000000LL) { atomicExch(err_flag, 1); if (done_flag) atomicExch(done_flag, 1); break; }   // ≈3 s: a peer never arrived → stop the solve
            }
        }
        __syncthreads();
        __threadfence_system();
    }
    xchg_grid_barrier(gbar, base + 2ULL * gridDim.x, err_flag, done_flag);
    // 4. scalars in rank order (bit-identical on all ranks)
    if (blockIdx.x == 0 && tid == 0 && nscal > 0) {
        const double* sc = reinterpret_cast<const double*>(mine + MB_OFF_SCAL) + (size_t)par * 32 * 2;
        double s0 = 0.0, s1 = 0.0;
        for (int r = 0; r < nranks; r++) { s0 += __ldcv(sc + 2 * r); s1 += __ldcv(sc + 2 * r + 1); }
        scal[0] = s0; if (nscal > 1) scal[1] = s1;
    }
    // interface sum, ascending rank order
    const double* recv = reinterpret_cast<const double*>(mine + MB_OFF_RECV) + (size_t)par * nranks * stride3;
    for (int i = gtid; i < 3 * n_if; i += gsz) {
        int k = i / 3, c = i - 3 * k;
        size_t dof = 3 * (size_t)if_node[k] + c;
        double own = y[dof], s = 0.0;
        for (int j = if_ptr[k]; j < if_ptr[k + 1]; j++) { int src = if_srcx[j]; s += src < 0 ? own : __ldcv(recv + 3 * (size_t)src + c); }
        y[dof] = s;
    }
}

// (re)creates the mailboxes and exchanges their IPC handles; collective.  Falls back to the NCCL path on any failure.
static int mailbox_setup(toe_ctx* ctx, DistState* d, const std::vector<int>& if_src_host) {
    d->p2p_ok = false;
    // opt-in (TOE_DIST_XCHG=p2p): 9 % faster than the all-gather at N=8, 1 % slower at N=4; clean in 38 solves at N=2/4/8
    if (d->nranks == 1 || xchg_mode(d->nranks) != XCHG_P2P) return TOE_OK;
    // agree on the receive-area stride
    int my_max = 1;
    for (int c : d->nbr_count) my_max = std::max(my_max, c);
    DevBuf<int> gm; CU(gm.alloc(2));
    CU(cudaMemcpyAsync(gm.p, &my_max, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NC(g_nccl.AllReduce(gm.p, gm.p, 1, ncclInt, ncclMax, d->comm, ctx->stream));
    int gmax = 0;
    CU(cudaMemcpyAsync(&gmax, gm.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    i64 stride3 = 3 * (i64)gmax;
    size_t need = MB_OFF_RECV + (size_t)2 * d->nranks * stride3 * sizeof(double);
    int ok = 1;
    if (need > d->mbox_bytes || d->peer_mbox.empty()) {
        // everybody drops its mappings of the old mailboxes, a collective orders that, then everybody frees its own
        mailbox_close_peers(d);
        NC(g_nccl.AllReduce(gm.p, gm.p, 1, ncclInt, ncclMax, d->comm, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        mailbox_free_own(d);
        size_t bytes = need + need / 2;
        cudaIpcMemHandle_t mine_h;
        memset(&mine_h, 0, sizeof mine_h);
        if (cudaMalloc((void**)&d->mbox, bytes) != cudaSuccess) { ok = 0; d->mbox = nullptr; }
        if (ok) { d->mbox_bytes = bytes; cudaMemset(d->mbox, 0, bytes); if (cudaIpcGetMemHandle(&mine_h, d->mbox) != cudaSuccess) ok = 0; }
        cudaGetLastError();
        // all-gather of the 64-byte handles through a byte-sum allreduce (own slot filled, the rest zero)
        const size_t hs = sizeof(cudaIpcMemHandle_t);
        std::vector<unsigned char> hbuf((size_t)d->nranks * hs + 8, 0);
        if (ok) memcpy(hbuf.data() + (size_t)d->rank * hs, &mine_h, hs);
        hbuf[(size_t)d->nranks * hs] = ok ? 0 : 1;          // failure votes
        DevBuf<unsigned char> db; CU(db.alloc(hbuf.size()));
        CU(cudaMemcpyAsync(db.p, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice, ctx->stream));
        NC(g_nccl.AllReduce(db.p, db.p, hbuf.size(), ncclUint8, ncclSum, d->comm, ctx->stream));
        CU(cudaMemcpyAsync(hbuf.data(), db.p, hbuf.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (hbuf[(size_t)d->nranks * hs] != 0) ok = 0;
        d->peer_mbox.assign(d->nranks, nullptr);
        if (ok) {
            for (int r = 0; r < d->nranks && ok; r++) {
                if (r == d->rank) { d->peer_mbox[r] = d->mbox; continue; }
                cudaIpcMemHandle_t h; memcpy(&h, hbuf.data() + (size_t)r * hs, hs);
                void* p = nullptr;
                if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
                d->peer_mbox[r] = (char*)p;
            }
        }
        // second vote: did every rank manage to map every peer?
        int vote = ok ? 0 : 1;
        CU(cudaMemcpyAsync(gm.p, &vote, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        NC(g_nccl.AllReduce(gm.p, gm.p, 1, ncclInt, ncclMax, d->comm, ctx->stream));
        CU(cudaMemcpyAsync(&vote, gm.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (vote) { mailbox_release(d); return TOE_OK; }
        CU(d->peer_mbox_dev.alloc(d->nranks));
        CU(cudaMemcpy(d->peer_mbox_dev.p, d->peer_mbox.data(), d->nranks * sizeof(char*), cudaMemcpyHostToDevice));
        d->xseq = 0;
    }
    d->stride3 = stride3;
    if (!d->gbar.p) { CU(d->gbar.alloc(1)); CU(cudaMemset(d->gbar.p, 0, sizeof(u64))); d->xlaunch = 0; }
    // send segments and receive indices in mailbox coordinates
    int nseg = (int)d->nbr.size();
    CU(d->seg_rank.alloc(nseg)); CU(d->seg_off.alloc(nseg)); CU(d->seg_cnt.alloc(nseg));
    if (nseg) {
        CU(cudaMemcpy(d->seg_rank.p, d->nbr.data(), nseg * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d->seg_off.p, d->nbr_off.data(), nseg * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d->seg_cnt.p, d->nbr_count.data(), nseg * sizeof(int), cudaMemcpyHostToDevice));
    }
    std::vector<int> srcx(if_src_host.size());
    for (size_t j = 0; j < if_src_host.size(); j++) {
        int p = if_src_host[j];
        if (p < 0) { srcx[j] = -1; continue; }
        int k = 0;
        while (k + 1 < nseg && p >= d->nbr_off[k + 1]) k++;
        srcx[j] = d->nbr[k] * (int)(stride3 / 3) + (p - d->nbr_off[k]);
    }
    CU(d->if_srcx.alloc(srcx.size()));
    if (!srcx.empty()) CU(cudaMemcpy(d->if_srcx.p, srcx.data(), srcx.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());                    // NULL-stream uploads / memsets above vs the non-blocking ctx stream
    d->p2p_ok = true;
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// All-gather exchange (the default transport).  Interfaces are small (157 KB per neighbour at 10M tets / N=2), so
// instead of a group of sends / receives plus an allreduce every rank contributes ONE slice — its packed interface values in
// its own send layout, then its two partial scalars — to ONE ncclAllGather; the unpack kernel reads the neighbours' segments
// straight out of the gathered buffer (offsets exchanged once per set-up) and sums the scalars in rank order.  Same arithmetic
// and summation order as the other transports → bit-identical iterates.  3 launches per exchange instead of 4, and no
// point-to-point traffic at all.
// ---------------------------------------------------------------------------------------------------------
static const int AG_SCALARS = 4;       // scalar slots at the end of every rank's slice ({γ, δ, ν, spare})
__global__ void k_pack_ag(const int* __restrict__ send_nodes, const double* __restrict__ y, const double* scal, int nscal,
                          double* __restrict__ slice, int n, i64 stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n) { int s = i / 3, c = i - 3 * s; slice[i] = y[3 * (size_t)send_nodes[s] + c]; }
    if (i < AG_SCALARS) slice[stride - AG_SCALARS + i] = (i < nscal) ? scal[i] : 0.0;
}
__global__ void k_unpack_ag(const int* __restrict__ if_node, const int* __restrict__ if_ptr, const int* __restrict__ if_src_ag,
                            const double* __restrict__ all, double* __restrict__ y, int n_if, double* scal, int nscal, int nranks, i64 stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nscal) {                                             // scalars in rank order: identical bits on every rank
        double s = 0.0;
        for (int r = 0; r < nranks; r++) s += all[(size_t)r * stride + stride - AG_SCALARS + i];
        scal[i] = s;
    }
    if (i >= 3 * n_if) return;
    int k = i / 3, c = i - 3 * k;
    size_t dof = 3 * (size_t)if_node[k] + c;
    double own = y[dof], s = 0.0;
    for (int j = if_ptr[k]; j < if_ptr[k + 1]; j++) { int src = if_src_ag[j]; s += src < 0 ? own : all[3 * (size_t)src + c]; }   // ascending rank order
    y[dof] = s;
}

// collective; called at every set-up.  Publishes, for every (rank r, peer q), where r's segment for q starts in r's send layout.
static int allgather_setup(toe_ctx* ctx, DistState* d, const std::vector<int>& if_src_host) {
    d->ag_ok = false;
    if (d->nranks == 1 || d->p2p_ok || xchg_mode(d->nranks) == XCHG_SENDRECV) return TOE_OK;      // all-gather: chosen, or the peer-memory set-up did not succeed
    const int R = d->nranks;
    std::vector<int> tab((size_t)R * R + 1, 0);                  // own row: 1 + offset of the segment for peer q (0 = not a neighbour); last: max n_shared_total
    for (size_t k = 0; k < d->nbr.size(); k++) tab[(size_t)d->rank * R + d->nbr[k]] = 1 + d->nbr_off[k];
    DevBuf<int> dt; CU(dt.alloc(tab.size()));
    CU(cudaMemcpyAsync(dt.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NC(g_nccl.AllReduce(dt.p, dt.p, (size_t)R * R, ncclInt, ncclSum, d->comm, ctx->stream));        // rows are disjoint: the sum is a gather
    DevBuf<int> dm; CU(dm.alloc(2));
    CU(cudaMemcpyAsync(dm.p, &d->n_shared_total, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NC(g_nccl.AllReduce(dm.p, dm.p, 1, ncclInt, ncclMax, d->comm, ctx->stream));
    int gmax = 0;
    CU(cudaMemcpyAsync(tab.data(), dt.p, (size_t)R * R * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&gmax, dm.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // per-rank slice: 3 * max_r(n_shared_total) interface doubles + AG_SCALARS scalars, rounded up to a multiple of 6 doubles so that slices
    // are 16-byte aligned AND start on a node slot (a source is then addressed by ONE int: node slot in the gathered buffer)
    d->ag_stride = (3 * (i64)gmax + AG_SCALARS + 5) / 6 * 6;
    const i64 slots_per_rank = d->ag_stride / 3;
    if ((i64)R * slots_per_rank > 2147483647LL / 4) return TOE_OK;        // would not fit the int indices: stay on send/recv
    // source of my recv slot p (segment k of MY layout, position j) = slice of rank nbr[k], its segment for me, position j
    std::vector<int> src(if_src_host.size());
    const int nseg = (int)d->nbr.size();
    for (size_t i = 0; i < if_src_host.size(); i++) {
        const int p = if_src_host[i];
        if (p < 0) { src[i] = -1; continue; }
        int k = 0;
        while (k + 1 < nseg && p >= d->nbr_off[k + 1]) k++;
        const int r = d->nbr[k];
        const int off_r = tab[(size_t)r * R + d->rank] - 1;
        if (off_r < 0) return toe_fail(ctx, TOE_ERR_COMM, "all-gather exchange: rank %d does not list rank %d as a neighbour (interface maps disagree)", r, d->rank);
        src[i] = (int)(r * slots_per_rank) + off_r + (p - d->nbr_off[k]);
    }
    CU(d->if_src_ag.alloc(src.size()));
    if (!src.empty()) CU(cudaMemcpyAsync(d->if_src_ag.p, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(d->ag_send.alloc(2 * (size_t)d->ag_stride)); CU(d->ag_recv.alloc(2 * (size_t)R * d->ag_stride));
    CU(cudaMemsetAsync(d->ag_send.p, 0, d->ag_send.bytes(), ctx->stream));       // slice padding travels: keep it defined
    CU(cudaStreamSynchronize(ctx->stream));                                       // `src` (pageable) goes out of scope
    d->ag_ok = true;
    return TOE_OK;
}

static int exchange_allgather(toe_ctx* ctx, double* y, double* scal, int count) {
    DistState* d = ctx->dist;
    const int n = d->n_shared_total;
    d->xpar ^= 1;
    double* slice = d->ag_send.p + (size_t)d->xpar * d->ag_stride;
    double* all = d->ag_recv.p + (size_t)d->xpar * d->nranks * d->ag_stride;
    const i64 work = std::max<i64>(3 * (i64)n, AG_SCALARS);
    LAUNCH(ctx, k_pack_ag, div_up(work, 256), 256, 0, (const int*)d->send_nodes.p, (const double*)y, (const double*)scal, count, slice, n, d->ag_stride);
    NC(g_nccl.AllGather(slice, all, (size_t)d->ag_stride, ncclDouble, d->comm, ctx->stream));
    const i64 work2 = std::max<i64>(3 * (i64)d->n_if, AG_SCALARS);
    LAUNCH(ctx, k_unpack_ag, div_up(work2, 256), 256, 0, (const int*)d->if_node.p, (const int*)d->if_ptr.p, (const int*)d->if_src_ag.p,
           (const double*)all, y, d->n_if, scal, count, d->nranks, d->ag_stride);
    return TOE_OK;
}

// interface sum of y (grouped ncclSend/ncclRecv) followed by the allreduce of `count` scalars
int dist_exchange_allreduce(toe_ctx* ctx, double* y, double* scal, int count) {
    DistState* d = ctx->dist;
    if (!d || d->nranks == 1) return TOE_OK;
    if (d->p2p_ok && count <= 2) {
        d->xseq++;
#ifdef TOE_EMU
        emu::next_launch_coresident();
#endif
        LAUNCH(ctx, k_xchg, XCHG_CTAS, XCHG_THREADS, 0, (char* const*)d->peer_mbox_dev.p, d->nranks, d->rank, d->xseq, d->stride3,
               (const int*)d->seg_rank.p, (const int*)d->seg_off.p, (const int*)d->seg_cnt.p, (int)d->nbr.size(), d->n_shared_total,
               (const int*)d->send_nodes.p, (const int*)d->if_node.p, (const int*)d->if_ptr.p, (const int*)d->if_srcx.p, d->n_if, y, scal, count,
               &ctx->cgs.p->done, ctx->errflag.p + 2, d->gbar.p, d->xlaunch++);
        return TOE_OK;
    }
    if (d->ag_ok && count <= AG_SCALARS) return exchange_allgather(ctx, y, scal, count);
    int n = d->n_shared_total;
    d->xpar ^= 1;
    double* sb = d->sendbuf.p + (size_t)d->xpar * 3 * n;
    double* rb = d->recvbuf.p + (size_t)d->xpar * 3 * n;
    if (n) LAUNCH(ctx, k_pack, div_up(3 * (i64)n, 256), 256, 0, (const int*)d->send_nodes.p, (const double*)y, sb, n);
    NC(g_nccl.GroupStart());
    for (size_t k = 0; k < d->nbr.size(); k++) {
        NC(g_nccl.Send(sb + 3 * (size_t)d->nbr_off[k], 3 * (size_t)d->nbr_count[k], ncclDouble, d->nbr[k], d->comm, ctx->stream));
        NC(g_nccl.Recv(rb + 3 * (size_t)d->nbr_off[k], 3 * (size_t)d->nbr_count[k], ncclDouble, d->nbr[k], d->comm, ctx->stream));
    }
    NC(g_nccl.GroupEnd());
    // kept outside the p2p group (conservative: per-rank p2p sets differ — edge parts have one neighbour, inner parts more)
    NC(g_nccl.AllReduce(scal, scal, (size_t)count, ncclDouble, ncclSum, d->comm, ctx->stream));
    if (n) LAUNCH(ctx, k_unpack_sum, div_up(3 * (i64)d->n_if, 256), 256, 0, (const int*)d->if_node.p, (const int*)d->if_ptr.p, (const int*)d->if_src.p,
                  (const double*)rb, y, d->n_if);
    return TOE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// global <-> local vectors (Ferrite dof order at the ABI)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_scatter_owned(const int* __restrict__ loc2glob, const unsigned char* __restrict__ owned, const double* __restrict__ local,
                                double* __restrict__ global, int nq_l) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * nq_l) return;
    int l = i / 3, c = i - 3 * l;
    if (owned[l]) global[3 * (size_t)loc2glob[l] + c] = local[i];
}
__global__ void k_gather_local(const int* __restrict__ loc2glob, const double* __restrict__ global, double* __restrict__ local, int nq_l) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * nq_l) return;
    int l = i / 3, c = i - 3 * l;
    local[i] = global[3 * (size_t)loc2glob[l] + c];
}
__global__ void k_scatter_cells(const int* __restrict__ eloc2glob, const double* __restrict__ local, double* __restrict__ global, i64 ne_l, int width) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne_l * width) return;
    i64 j = i / width; int w = (int)(i - j * width);
    global[(size_t)eloc2glob[j] * width + w] = local[i];
}
__global__ void k_gather_cells(const int* __restrict__ eloc2glob, const double* __restrict__ global, double* __restrict__ local, i64 ne_l) {
    i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ne_l) local[j] = global[eloc2glob[j]];
}

int dist_gather_vector(toe_ctx* ctx, const double* local_dev, double* global_host) {
    DistState* d = ctx->dist;
    size_t ng = 3 * (size_t)d->nq_g;
    CU(d->gvec.alloc(ng));
    CU(cudaMemsetAsync(d->gvec.p, 0, ng * sizeof(double), ctx->stream));
    LAUNCH(ctx, k_scatter_owned, div_up(3 * (i64)ctx->nq, 256), 256, 0, (const int*)d->loc2glob.p, (const unsigned char*)d->owned.p, local_dev, d->gvec.p, ctx->nq);
    if (d->nranks > 1) NC(g_nccl.AllReduce(d->gvec.p, d->gvec.p, ng, ncclDouble, ncclSum, d->comm, ctx->stream));
    CU(cudaMemcpyAsync(global_host, d->gvec.p, ng * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

int dist_scatter_vector(toe_ctx* ctx, const double* global_host, double* local_dev) {
    DistState* d = ctx->dist;
    size_t ng = 3 * (size_t)d->nq_g;
    CU(d->gvec.alloc(ng));
    CU(cudaMemcpyAsync(d->gvec.p, global_host, ng * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_gather_local, div_up(3 * (i64)ctx->nq, 256), 256, 0, (const int*)d->loc2glob.p, (const double*)d->gvec.p, local_dev, ctx->nq);
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

int dist_sum_per_element_w(toe_ctx* ctx, const double* local_dev, double* global_host, int width);
int dist_sum_per_element(toe_ctx* ctx, const double* local_dev, double* global_host) { return dist_sum_per_element_w(ctx, local_dev, global_host, 1); }

int dist_sum_per_element_w(toe_ctx* ctx, const double* local_dev, double* global_host, int width) {
    DistState* d = ctx->dist;
    size_t ng = (size_t)d->ne_g * width;
    CU(d->gvec.alloc(ng));
    CU(cudaMemsetAsync(d->gvec.p, 0, ng * sizeof(double), ctx->stream));
    LAUNCH(ctx, k_scatter_cells, div_up(ctx->ne * width, 256), 256, 0, (const int*)d->eloc2glob.p, local_dev, d->gvec.p, ctx->ne, width);
    if (d->nranks > 1) NC(g_nccl.AllReduce(d->gvec.p, d->gvec.p, ng, ncclDouble, ncclSum, d->comm, ctx->stream));
    CU(cudaMemcpyAsync(global_host, d->gvec.p, ng * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

// per-cell input given in global cell order (densities, Lamé fields) -> local device array
int dist_localize_cells(toe_ctx* ctx, const double* global_host, double* local_dev) {
    DistState* d = ctx->dist;
    size_t ng = (size_t)d->ne_g;
    CU(d->gvec.alloc(ng));
    CU(cudaMemcpyAsync(d->gvec.p, global_host, ng * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_gather_cells, div_up(ctx->ne, 256), 256, 0, (const int*)d->eloc2glob.p, (const double*)d->gvec.p, local_dev, ctx->ne);
    CU(cudaStreamSynchronize(ctx->stream));
    return TOE_OK;
}

int dist_node_dofs(toe_ctx* ctx, const int** node_q_g) { *node_q_g = ctx->dist->node_q_g.p; return TOE_OK; }

int dist_get_partition(toe_ctx* ctx, int32_t* part) {
    DistState* d = ctx->dist;
    if (!d || !d->part.p) return toe_fail(ctx, TOE_ERR_STATE, "no partition: call toe_set_mesh_distributed first");
    CU(cudaMemcpy(part, d->part.p, d->ne_g * sizeof(int), cudaMemcpyDeviceToHost));
    return TOE_OK;
}

int dist_local_sizes(toe_ctx* ctx, int64_t* a, int64_t* b, int64_t* c, int64_t* e) {
    if (a) *a = ctx->ne; if (b) *b = 3 * (int64_t)ctx->nq; if (c) *c = 9 * ctx->nnzb;
    if (e) *e = ctx->dist ? 3 * (int64_t)ctx->dist->n_if : 0;
    return TOE_OK;
}

i64 dist_global_ne(toe_ctx* ctx) { return ctx->dist ? ctx->dist->ne_g : ctx->ne; }
i64 dist_global_ndofs(toe_ctx* ctx) { return ctx->dist ? 3 * (i64)ctx->dist->nq_g : 3 * (i64)ctx->nq; }

// global arg-max over ranks of a per-rank (value, local cell) pair; ties go to the smallest global cell id
int dist_argmax(toe_ctx* ctx, double* max_inout, i64* cell_inout) {
    DistState* d = ctx->dist;
    int gl = 0;
    CU(cudaMemcpy(&gl, d->eloc2glob.p + *cell_inout, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<double> h(2 * d->nranks, 0.0);
    h[2 * d->rank] = *max_inout; h[2 * d->rank + 1] = (double)gl;
    DevBuf<double> buf; CU(buf.alloc(h.size()));
    CU(cudaMemcpyAsync(buf.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (d->nranks > 1) NC(g_nccl.AllReduce(buf.p, buf.p, h.size(), ncclDouble, ncclSum, d->comm, ctx->stream));
    CU(cudaMemcpyAsync(h.data(), buf.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    double best = -1.0; i64 arg = 0;
    for (int r = 0; r < d->nranks; r++) {
        double v = h[2 * r]; i64 a = (i64)h[2 * r + 1];
        if (v > best || (v == best && a < arg)) { best = v; arg = a; }
    }
    *max_inout = best; *cell_inout = arg;
    return TOE_OK;
}

// did a peer-memory exchange give up waiting for a peer?  (checked once per solve)
int dist_check_exchange(toe_ctx* ctx) {
    if (!ctx->dist || !ctx->dist->p2p_ok) return TOE_OK;
    int e = 0;
    CU(cudaMemcpyAsync(&e, ctx->errflag.p + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (e) {
        CU(cudaMemsetAsync(ctx->errflag.p + 2, 0, sizeof(int), ctx->stream));
        return toe_fail(ctx, TOE_ERR_COMM, "peer-memory exchange timed out waiting for another rank (rank %d of %d)", ctx->dist->rank, ctx->dist->nranks);
    }
    return TOE_OK;
}

int dist_info(toe_ctx* ctx, int* nranks, int* rank, int* transport) {
    DistState* d = ctx->dist;
    if (nranks) *nranks = d ? d->nranks : 1;
    if (rank) *rank = d ? d->rank : 0;
    if (transport) *transport = (!d || d->nranks == 1) ? 0 : (d->p2p_ok ? 2 : (d->ag_ok ? 3 : 1));
    return TOE_OK;
}
