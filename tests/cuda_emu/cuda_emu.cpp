// cuda_emu.cpp — runtime of the host-side CUDA emulation used by the tests (see cuda_emu.h: TEST INFRASTRUCTURE ONLY).
#include "cuda_emu.h"
#include "include/nccl.h"

#include <sys/mman.h>
#include <time.h>
#include <stdio.h>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <chrono>
#include <algorithm>

#if !defined(__x86_64__)
#error "the fiber switch below is written for x86-64 (System V ABI)"
#endif

thread_local emu_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

// ---------------------------------------------------------------------------------------------------------------------
// fibers: callee-saved registers + stack pointer; everything else lives on the fiber's own stack
// ---------------------------------------------------------------------------------------------------------------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace {

struct PendingCopy { void* dst; const void* src; unsigned bytes; };
struct Fiber { void* sp = nullptr; char* stack = nullptr; bool done = true; std::vector<PendingCopy> pending; };

struct Warp { unsigned long long slot[2][32]; unsigned arrived; unsigned long long gen; unsigned live; };

struct BlockState {
    unsigned nthreads = 0, live = 0;
    unsigned bar_arrived = 0; unsigned long long bar_gen = 0;
    unsigned nb_arrived[16]; unsigned long long nb_gen[16];
    Warp warps[34];
};

struct GraphNode { unsigned grid, block; size_t smem; std::function<void()> body; };
}  // namespace
struct emu_graph { std::vector<GraphNode> nodes; };
struct emu_stream { int id; };
struct emu_event { struct timespec t; };

namespace {
struct Rank {                      // per OS thread (= per emulated GPU / ctx driver thread)
    std::vector<Fiber> fibers;
    void* sched_sp = nullptr;
    int cur = -1;
    BlockState bs0;
    BlockState* bs = &bs0;                 // state of the block the running fiber belongs to
    std::vector<BlockState> co_bs;         // co-resident launches: one state per block
    bool next_coresident = false;
    const std::function<void()>* body = nullptr;
    char* dyn = nullptr; size_t dyn_cap = 0;
    long long clk = 0;
    unsigned long long progress = 0;
    emu_graph* capture = nullptr;
    cudaError_t last_error = cudaSuccess;
    size_t stack_bytes = 0;
    bool deadlock = false;
    // EMU_BULK_DELAY: bulk async copies complete late and out of order (see emu::bulk_g2s)
    struct PendingCopy { void* dst; const void* src; unsigned bytes; void* bar; unsigned long long due; };
    std::vector<PendingCopy> pending;
    unsigned long long pass = 0, fill_rng = 0x2545F4914F6CDD1DULL;
    void* last_fill_bar = nullptr; unsigned long long last_fill_pass = ~0ULL, last_fill_delay = 0;
};
thread_local Rank* g_rank = nullptr;

Rank& rank_state() {
    if (!g_rank) {
        g_rank = new Rank();
        const char* kb = getenv("EMU_STACK_KB");
        g_rank->stack_bytes = (size_t)(kb ? atoi(kb) : 256) * 1024;
    }
    return *g_rank;
}

void fiber_entry() {
    Rank& r = *g_rank;
    (*r.body)();
    Fiber& f = r.fibers[r.cur];
    for (auto& c : f.pending) memcpy(c.dst, c.src, c.bytes);     // the hardware completes outstanding copies eventually
    f.pending.clear();
    f.done = true;
    r.progress++;
    emu_switch(&f.sp, r.sched_sp);
    abort();                       // never resumed
}

void fiber_prepare(Rank& r, int t) {
    if ((int)r.fibers.size() <= t) r.fibers.resize(t + 1);
    Fiber& f = r.fibers[t];
    if (!f.stack) {
        void* p = mmap(nullptr, r.stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) { perror("cuda_emu: mmap fiber stack"); abort(); }
        f.stack = (char*)p;
    }
    uintptr_t top = ((uintptr_t)f.stack + r.stack_bytes) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                       // fake return address of fiber_entry (keeps rsp ≡ 8 mod 16 at its entry)
    *--sp = (void*)&fiber_entry;           // popped by `ret` in emu_switch
    for (int k = 0; k < 6; k++) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = (void*)sp;
    f.done = false;
    f.pending.clear();
}

void block_state_init(BlockState& bs, unsigned T) {
    bs.nthreads = T; bs.live = T; bs.bar_arrived = 0; bs.bar_gen = 0;
    for (int k = 0; k < 16; k++) { bs.nb_arrived[k] = 0; bs.nb_gen[k] = 0; }
    unsigned nw = (T + 31) / 32;
    for (unsigned w = 0; w < nw; w++) { bs.warps[w].arrived = 0; bs.warps[w].gen = 0; bs.warps[w].live = (w + 1) * 32 <= T ? 32 : T - w * 32; }
}

// all blocks of the grid resident at once (kernels with a software grid barrier); such kernels must not use __shared__
void run_grid_coresident(Rank& r, unsigned grid, unsigned T) {
    r.co_bs.resize(grid);
    for (unsigned b = 0; b < grid; b++) block_state_init(r.co_bs[b], T);
    unsigned total = grid * T;
    for (unsigned i = 0; i < total; i++) fiber_prepare(r, (int)i);
    unsigned remaining = total;
    unsigned long long idle_passes = 0;
    while (remaining > 0) {
        unsigned long long p0 = r.progress;
        for (unsigned i = 0; i < total; i++) {
            Fiber& f = r.fibers[i];
            if (f.done) continue;
            unsigned b = i / T, t = i - b * T;
            r.cur = (int)i; blockIdx.x = b; threadIdx.x = t; r.bs = &r.co_bs[b];
            emu_switch(&r.sched_sp, f.sp);
            if (f.done) { remaining--; r.co_bs[b].live--; r.co_bs[b].warps[t / 32].live--; }
        }
        if (r.progress == p0) {
            if (++idle_passes > 200000000ULL) {
                fprintf(stderr, "cuda_emu: DEADLOCK in a co-resident %u x %u kernel (%u threads never finished)\n", grid, T, remaining);
                r.last_error = 719;
                for (unsigned i = 0; i < total; i++) r.fibers[i].done = true;
                break;
            }
        } else idle_passes = 0;
    }
    r.bs = &r.bs0;
}

// bulk async copies whose (emulated) completion time has come: data lands, the transaction bytes are taken off the barrier
void deliver_due_copies(Rank& r, bool all) {
    for (size_t k = 0; k < r.pending.size();) {
        Rank::PendingCopy& c = r.pending[k];
        if (!all && c.due > r.pass) { k++; continue; }
        memcpy(c.dst, c.src, c.bytes);
        unsigned long long* b = (unsigned long long*)c.bar;
        unsigned ph = (unsigned)(*b & 1); int cnt = (int)((*b >> 1) & 0xfff), pend = (int)((*b >> 13) & 0xfff); long long tx = (long long)(int)(*b >> 32);
        tx -= c.bytes;
        *b = (unsigned long long)(ph & 1) | ((unsigned long long)(cnt & 0xfff) << 1) | ((unsigned long long)(pend & 0xfff) << 13) | ((unsigned long long)(unsigned)(int)tx << 32);
        if (pend == 0 && tx == 0) *b = (unsigned long long)((ph ^ 1u) & 1) | ((unsigned long long)(cnt & 0xfff) << 1) | ((unsigned long long)(cnt & 0xfff) << 13);
        r.progress++;
        r.pending[k] = r.pending.back(); r.pending.pop_back();
    }
}

void run_block(Rank& r, unsigned b, unsigned T, size_t smem) {
    blockIdx.x = b;
    BlockState& bs = r.bs0;
    r.bs = &bs;
    block_state_init(bs, T);
    if (smem) memset(r.dyn, 0xCD, smem);                 // uninitialised dynamic shared memory is garbage, not zeros
    for (unsigned t = 0; t < T; t++) fiber_prepare(r, (int)t);
    unsigned remaining = T;
    unsigned long long idle_passes = 0;
    // EMU_SHUFFLE=<seed>: threads of a block are resumed in a fresh pseudo-random order on every pass instead of 0,1,2,… — code
    // that silently depends on the order in which threads happen to run (a missing barrier) then gives different answers
    static const char* shuffle_env = getenv("EMU_SHUFFLE");
    static thread_local std::vector<unsigned> order;
    static thread_local unsigned long long rng = 0;
    if (shuffle_env) { if (!rng) rng = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)atoll(shuffle_env); order.resize(T); for (unsigned t = 0; t < T; t++) order[t] = t; }
    while (remaining > 0) {
        unsigned long long p0 = r.progress;
        r.pass++;
        deliver_due_copies(r, false);
        if (shuffle_env)
            for (unsigned t = T; t > 1; t--) { rng = rng * 6364136223846793005ULL + 1442695040888963407ULL; unsigned j = (unsigned)((rng >> 33) % t); std::swap(order[t - 1], order[j]); }
        for (unsigned tt = 0; tt < T; tt++) {
            const unsigned t = shuffle_env ? order[tt] : tt;
            Fiber& f = r.fibers[t];
            if (f.done) continue;
            r.cur = (int)t; threadIdx.x = t;
            emu_switch(&r.sched_sp, f.sp);
            if (f.done) { remaining--; bs.live--; bs.warps[t / 32].live--; }
        }
        if (r.progress == p0) {
            if (++idle_passes > 2000000ULL) {
                fprintf(stderr, "cuda_emu: DEADLOCK in block %u of a %u-thread kernel (%u threads never finished)\n", b, T, remaining);
                r.deadlock = true; r.last_error = 719;
                for (unsigned t = 0; t < T; t++) r.fibers[t].done = true;    // abandon the block
                return;
            }
        } else idle_passes = 0;
    }
    if (!r.pending.empty()) {                              // a block must have waited for every copy it issued
        fprintf(stderr, "cuda_emu: block %u ended with %zu bulk copies in flight\n", b, r.pending.size());
        r.last_error = 719;
        r.pending.clear();
    }
}

static int block_order_from_env() {
    const char* e = getenv("EMU_BLOCKS");
    return !e ? 0 : (!strcmp(e, "reverse") ? 1 : 2);
}
std::atomic<int> g_block_order{block_order_from_env()};

void run_grid(Rank& r, unsigned grid, unsigned block, size_t smem, const std::function<void()>& body) {
    if (block == 0 || block > 1024) { r.last_error = cudaErrorInvalidValue; return; }
    if (smem > 227 * 1024) { r.last_error = cudaErrorInvalidValue; return; }
    if (smem > r.dyn_cap) { free(r.dyn); r.dyn = nullptr; if (posix_memalign((void**)&r.dyn, 1024, smem + 1024)) abort(); r.dyn_cap = smem; }
    gridDim.x = grid; blockDim.x = block;
    r.body = &body;
    if (r.next_coresident) { r.next_coresident = false; run_grid_coresident(r, grid, block); }
    else {
        // EMU_BLOCKS=reverse | shuffle: the blocks of a grid run last-to-first / in a fresh pseudo-random order.  CUDA promises no block
        // order, so every result must stay the same (bit for bit where the code claims determinism); a kernel in which some block
        // silently relies on another block of the SAME launch having run (or not yet run) gives different answers here.
        const int mode = g_block_order.load();                 // 0 forward, 1 reverse, 2 shuffle (EMU_BLOCKS or emu_set_block_order)
        if (mode == 0) { for (unsigned b = 0; b < grid && !r.deadlock; b++) run_block(r, b, block, smem); }
        else if (mode == 1) { for (unsigned b = grid; b-- > 0 && !r.deadlock;) run_block(r, b, block, smem); }
        else {
            static thread_local unsigned long long brng = 0x2545F4914F6CDD1DULL;
            std::vector<unsigned> ord(grid);
            for (unsigned b = 0; b < grid; b++) ord[b] = b;
            for (unsigned b = grid; b > 1; b--) { brng = brng * 6364136223846793005ULL + 1442695040888963407ULL; std::swap(ord[b - 1], ord[(unsigned)((brng >> 33) % b)]); }
            for (unsigned k = 0; k < grid && !r.deadlock; k++) run_block(r, ord[k], block, smem);
        }
    }
    r.body = nullptr; r.cur = -1;
    r.deadlock = false;
}
}  // namespace

namespace {
std::mutex g_attr_mutex;
std::map<const void*, int> g_max_dyn_smem;
}
extern "C" void emu_set_block_order(int mode) { g_block_order.store(mode); }
extern "C" void emu_set_max_dyn_smem(const void* fn, int bytes) { std::lock_guard<std::mutex> lk(g_attr_mutex); g_max_dyn_smem[fn] = bytes; }
extern "C" void emu_misaligned(const void* p, unsigned bytes) {
    fprintf(stderr, "cuda_emu: MISALIGNED %u-byte load at %p\n", bytes, p);
    if (g_rank) g_rank->last_error = 716;
}

namespace emu {
void launch(unsigned grid, unsigned block, size_t smem, std::function<void()> body, const void* fn) {
    Rank& r = rank_state();
    if (smem > 48 * 1024 && fn) {             // CUDA refuses more than 48 KB of dynamic shared memory unless the kernel opted in
        std::lock_guard<std::mutex> lk(g_attr_mutex);
        auto it = g_max_dyn_smem.find(fn);
        if (it == g_max_dyn_smem.end() || (size_t)it->second < smem) {
            fprintf(stderr, "cuda_emu: launch with %zu bytes of dynamic shared memory without cudaFuncSetAttribute(MaxDynamicSharedMemorySize)\n", smem);
            r.last_error = cudaErrorInvalidValue;
            return;
        }
    }
    if (grid == 0) { r.last_error = cudaErrorInvalidValue; return; }    // CUDA: invalid configuration
    if (r.capture) { r.capture->nodes.push_back(GraphNode{grid, block, smem, std::move(body)}); return; }
    run_grid(r, grid, block, smem, body);
}
void* dyn_smem() { return g_rank->dyn; }
void cp_async(void* dst, const void* src, unsigned bytes) {
    if ((bytes != 4 && bytes != 8 && bytes != 16) || ((size_t)dst % bytes) || ((size_t)src % bytes)) {
        fprintf(stderr, "cuda_emu: illegal cp.async (%u bytes, dst %p, src %p)\n", bytes, dst, src);
        g_rank->last_error = 716;
    }
    Rank& r = *g_rank;
    r.fibers[r.cur].pending.push_back(PendingCopy{dst, src, bytes});
}
void cp_async_wait_all() {
    Rank& r = *g_rank;
    Fiber& f = r.fibers[r.cur];
    for (auto& c : f.pending) memcpy(c.dst, c.src, c.bytes);
    f.pending.clear();
}
void next_launch_coresident() { rank_state().next_coresident = true; }
static thread_local unsigned g_jitter_state = 12345u;
void yield() {
    Rank& r = *g_rank; Fiber& f = r.fibers[r.cur];
    static const bool jitter = getenv("EMU_JITTER") != nullptr;          // perturbs the interleaving of rank threads
    if (jitter && r.bs != &r.bs0) {            // only inside co-resident (exchange) kernels
        g_jitter_state = g_jitter_state * 1664525u + 1013904223u;
        if ((g_jitter_state >> 20) % 257 == 0) { struct timespec ts = {0, (long)((g_jitter_state >> 8) % 200000)}; nanosleep(&ts, nullptr); }
    }
    emu_switch(&f.sp, r.sched_sp);
}
void sync_threads() {
    Rank& r = *g_rank; BlockState& bs = *r.bs;
    unsigned long long g = bs.bar_gen;
    bs.bar_arrived++;
    for (;;) {
        if (bs.bar_gen != g) return;
        if (bs.bar_arrived >= bs.live) { bs.bar_arrived = 0; bs.bar_gen++; r.progress++; return; }
        yield();
    }
}
void named_barrier(int id, int count) {
    Rank& r = *g_rank; BlockState& bs = *r.bs;
    unsigned long long g = bs.nb_gen[id];
    bs.nb_arrived[id]++;
    for (;;) {
        if (bs.nb_gen[id] != g) return;
        if ((int)bs.nb_arrived[id] >= count) { bs.nb_arrived[id] = 0; bs.nb_gen[id]++; r.progress++; return; }
        yield();
    }
}
unsigned long long shfl(unsigned long long v, int src_lane) {
    Rank& r = *g_rank;
    unsigned t = threadIdx.x;
    Warp& w = r.bs->warps[t / 32];
    unsigned long long g = w.gen;
    int buf = (int)(g & 1);
    w.slot[buf][t & 31] = v;
    w.arrived++;
    for (;;) {
        if (w.gen != g) break;
        if (w.arrived >= w.live) { w.arrived = 0; w.gen++; r.progress++; break; }
        yield();
    }
    return w.slot[buf][src_lane & 31];
}
long long clock() { return g_rank->clk += 1000; }

// mbarrier word: bit 0 phase | bits 1..12 expected arrivals | bits 13..24 pending arrivals | bits 32..63 pending tx bytes (signed)
static inline void mb_unpack(unsigned long long w, unsigned& phase, int& count, int& pending, long long& tx) {
    phase = (unsigned)(w & 1); count = (int)((w >> 1) & 0xfff); pending = (int)((w >> 13) & 0xfff); tx = (long long)(int)(w >> 32);
}
static inline unsigned long long mb_pack(unsigned phase, int count, int pending, long long tx) {
    return (unsigned long long)(phase & 1) | ((unsigned long long)(count & 0xfff) << 1) | ((unsigned long long)(pending & 0xfff) << 13) |
           ((unsigned long long)(unsigned)(int)tx << 32);
}
static inline void mb_complete_if_ready(unsigned long long* b) {
    unsigned ph; int c, p; long long tx; mb_unpack(*b, ph, c, p, tx);
    if (p == 0 && tx == 0) { *b = mb_pack(ph ^ 1u, c, c, 0); g_rank->progress++; }
}
void mbar_init(void* bar, unsigned count) { *(unsigned long long*)bar = mb_pack(0, (int)count, (int)count, 0); }
void mbar_expect_tx(void* bar, unsigned bytes) {                 // arrive + expect_tx
    unsigned long long* b = (unsigned long long*)bar;
    unsigned ph; int c, p; long long tx; mb_unpack(*b, ph, c, p, tx);
    *b = mb_pack(ph, c, p - 1, tx + bytes);
    mb_complete_if_ready(b);
}
void mbar_arrive(void* bar) {
    unsigned long long* b = (unsigned long long*)bar;
    unsigned ph; int c, p; long long tx; mb_unpack(*b, ph, c, p, tx);
    *b = mb_pack(ph, c, p - 1, tx);
    mb_complete_if_ready(b);
}
void mbar_wait(void* bar, unsigned parity) {                      // returns once the phase with this parity has completed
    volatile unsigned long long* b = (volatile unsigned long long*)bar;
    while ((unsigned)(*b & 1) == (parity & 1u)) yield();
}
void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    if ((bytes & 15) || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) {
        fprintf(stderr, "cuda_emu: cp.async.bulk with unaligned operands (dst %p src %p bytes %u)\n", dst, src, bytes);
        g_rank->last_error = 716;
    }
    // EMU_BULK_DELAY=<passes>: the copies of one fill (same barrier, issued in the same scheduler pass) complete together but LATE —
    // after 0 or <passes> scheduler passes, pseudo-randomly per fill — so the fills of consecutive stages land out of order, as they
    // may on hardware.  A pipeline protocol whose waits can be satisfied by the wrong phase then reads a stage before its data is there.
    const char* delay_s = getenv("EMU_BULK_DELAY");
    const long long delay_env = delay_s ? atoll(delay_s) : 0;
    Rank& r = *g_rank;
    if (delay_env > 0 && r.bs == &r.bs0) {
        if (r.last_fill_bar != bar || r.last_fill_pass != r.pass) {
            r.fill_rng = r.fill_rng * 6364136223846793005ULL + 1442695040888963407ULL;
            r.last_fill_delay = ((r.fill_rng >> 40) % 3 == 0) ? (unsigned long long)delay_env : 0ULL;
            r.last_fill_bar = bar; r.last_fill_pass = r.pass;
        }
        r.pending.push_back({dst, src, bytes, bar, r.pass + 1 + r.last_fill_delay});
        return;
    }
    memcpy(dst, src, bytes);
    unsigned long long* b = (unsigned long long*)bar;
    unsigned ph; int c, p; long long tx; mb_unpack(*b, ph, c, p, tx);
    *b = mb_pack(ph, c, p, tx - bytes);
    mb_complete_if_ready(b);
}
}  // namespace emu

// ---------------------------------------------------------------------------------------------------------------------
// memory: 0xFF-filled allocations with canary zones
// ---------------------------------------------------------------------------------------------------------------------
namespace {
const size_t GUARD = 256;
std::mutex g_mem_mutex;
std::map<void*, size_t> g_allocs;
size_t g_live_bytes = 0, g_peak_bytes = 0;

bool check_guards(void* user, size_t bytes) {
    unsigned char* base = (unsigned char*)user - GUARD;
    for (size_t i = 0; i < GUARD; i++) if (base[i] != 0xA5) return false;
    unsigned char* tail = (unsigned char*)user + bytes;
    for (size_t i = 0; i < GUARD; i++) if (tail[i] != 0xA5) return false;
    return true;
}
}  // namespace

extern "C" {

size_t emu_peak_bytes(void) { std::lock_guard<std::mutex> lk(g_mem_mutex); return g_peak_bytes; }
size_t emu_live_allocations(void) { std::lock_guard<std::mutex> lk(g_mem_mutex); return g_allocs.size(); }
// number of allocations whose canary zones were overwritten (out-of-bounds writes)
int emu_check_all_guards(void) {
    std::lock_guard<std::mutex> lk(g_mem_mutex);
    int bad = 0;
    for (auto& kv : g_allocs) if (!check_guards(kv.first, kv.second)) bad++;
    return bad;
}

cudaError_t cudaMalloc(void** p, size_t bytes) {
    if (!p) return cudaErrorInvalidValue;
    size_t rounded = (bytes + 255) & ~(size_t)255;
    unsigned char* base = nullptr;
    if (posix_memalign((void**)&base, 256, rounded + 2 * GUARD)) { *p = nullptr; return cudaErrorMemoryAllocation; }
    memset(base, 0xA5, GUARD);
    memset(base + GUARD, 0xFF, rounded);
    memset(base + GUARD + bytes, 0xA5, GUARD);          // canary right after the requested size (the rounding slack stays 0xFF)
    *p = base + GUARD;
    std::lock_guard<std::mutex> lk(g_mem_mutex);
    g_allocs[*p] = bytes;
    g_live_bytes += bytes; if (g_live_bytes > g_peak_bytes) g_peak_bytes = g_live_bytes;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    size_t bytes;
    {
        std::lock_guard<std::mutex> lk(g_mem_mutex);
        auto it = g_allocs.find(p);
        if (it == g_allocs.end()) { fprintf(stderr, "cuda_emu: cudaFree of unknown pointer %p\n", p); return cudaErrorInvalidValue; }
        bytes = it->second; g_allocs.erase(it); g_live_bytes -= bytes;
    }
    if (!check_guards(p, bytes)) {
        fprintf(stderr, "cuda_emu: OUT-OF-BOUNDS WRITE detected around allocation %p (%zu bytes)\n", p, bytes);
        if (g_rank) g_rank->last_error = 700;
        if (getenv("EMU_ABORT_ON_OOB")) abort();
    }
    size_t rounded = (bytes + 255) & ~(size_t)255;
    memset((unsigned char*)p - GUARD, 0xEE, rounded + 2 * GUARD);      // poison: use-after-free reads give garbage
    free((unsigned char*)p - GUARD);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t bytes) { return posix_memalign(p, 64, bytes ? bytes : 64) ? cudaErrorMemoryAllocation : cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }

static cudaError_t capture_guard(const char* what) {
    if (g_rank && g_rank->capture) { fprintf(stderr, "cuda_emu: %s during stream capture is not supported\n", what); g_rank->last_error = cudaErrorStreamCaptureUnsupported; return cudaErrorStreamCaptureUnsupported; }
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { cudaError_t e = capture_guard("cudaMemcpy"); if (e) return e; if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(dst, src, bytes, k); }
cudaError_t cudaMemset(void* dst, int v, size_t bytes) { cudaError_t e = capture_guard("cudaMemset"); if (e) return e; if (bytes) memset(dst, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t) { return cudaMemset(dst, v, bytes); }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { rank_state(); *s = new emu_stream{1}; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return capture_guard("cudaStreamSynchronize"); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return capture_guard("cudaDeviceSynchronize"); }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event(); clock_gettime(CLOCK_MONOTONIC, &(*e)->t); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { clock_gettime(CLOCK_MONOTONIC, &e->t); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)((b->t.tv_sec - a->t.tv_sec) * 1e3 + (b->t.tv_nsec - a->t.tv_nsec) * 1e-6);
    if (*ms <= 0.f) *ms = 1e-6f;
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return rank_state().last_error; }
cudaError_t cudaGetLastError(void) { Rank& r = rank_state(); cudaError_t e = r.last_error; r.last_error = cudaSuccess; return e; }
const char* cudaGetErrorName(cudaError_t e) {
    switch (e) { case 0: return "cudaSuccess"; case 1: return "cudaErrorInvalidValue"; case 2: return "cudaErrorMemoryAllocation";
                 case 700: return "emuOutOfBoundsWrite"; case 716: return "emuMisalignedBulkCopy"; case 719: return "emuDeadlock";
                 case 801: return "cudaErrorNotSupported"; case 900: return "cudaErrorStreamCaptureUnsupported"; default: return "emuError"; }
}
const char* cudaGetErrorString(cudaError_t e) { return cudaGetErrorName(e); }
cudaError_t cudaSetDevice(int d) { return (d >= 0 && d < 8) ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "cuda_emu host emulation (not a GPU)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
    Rank& r = rank_state();
    if (r.capture) return cudaErrorInvalidValue;
    r.capture = new emu_graph();
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) {
    Rank& r = rank_state();
    if (!r.capture) { *g = nullptr; return cudaErrorInvalidValue; }
    *g = r.capture; r.capture = nullptr;
    if (r.last_error == cudaErrorStreamCaptureUnsupported) { delete *g; *g = nullptr; return cudaErrorStreamCaptureUnsupported; }
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* ge, cudaGraph_t g, unsigned long long) { *ge = new emu_graph(*g); return cudaSuccess; }
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t) {
    Rank& r = rank_state();
    for (auto& n : g->nodes) run_grid(r, n.grid, n.block, n.smem, n.body);
    return cudaSuccess;
}
// rank threads share one address space: the handle simply carries the pointer
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); h->reserved[16] = 1; return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { if (h.reserved[16] != 1) return cudaErrorInvalidValue; memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// NCCL stand-in: ranks are threads of this process; collectives are blocking rendezvous (streams are synchronous here)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct World {
    int n = 0, joined = 0, alive = 0;
    std::mutex m;
    std::condition_variable cv;
    // barrier
    int bar_arrived = 0; unsigned long long bar_gen = 0;
    std::vector<const void*> send;
    std::map<std::pair<int, int>, std::deque<std::vector<char>>> p2p;    // (src, dst) -> messages in order
};
struct PendingOp { bool is_send; const void* sbuf; void* rbuf; size_t bytes; int peer; struct emu_nccl_comm* comm; };
std::mutex g_world_mutex;
std::map<std::string, std::shared_ptr<World>> g_worlds;
thread_local int g_group_depth = 0;
thread_local std::vector<PendingOp> g_group_ops;
const auto NCCL_TIMEOUT = std::chrono::seconds(120);
}  // namespace
struct emu_nccl_comm { std::shared_ptr<World> w; int rank; std::string key; };

namespace {
size_t dt_size(ncclDataType_t dt) {
    switch (dt) { case ncclInt8: case ncclUint8: return 1; case ncclInt32: case ncclUint32: case ncclFloat32: return 4; default: return 8; }
}
bool world_barrier(World& w, std::unique_lock<std::mutex>& lk) {
    unsigned long long g = w.bar_gen;
    if (++w.bar_arrived == w.n) { w.bar_arrived = 0; w.bar_gen++; w.cv.notify_all(); return true; }
    return w.cv.wait_for(lk, NCCL_TIMEOUT, [&] { return w.bar_gen != g; });
}
template <class T> void reduce_into(T* acc, const T* x, size_t n, ncclRedOp_t op, bool first) {
    for (size_t i = 0; i < n; i++) {
        if (first) acc[i] = x[i];
        else if (op == ncclSum) acc[i] = (T)(acc[i] + x[i]);
        else if (op == ncclMax) acc[i] = x[i] > acc[i] ? x[i] : acc[i];
        else if (op == ncclMin) acc[i] = x[i] < acc[i] ? x[i] : acc[i];
        else acc[i] = (T)(acc[i] * x[i]);
    }
}
ncclResult_t run_p2p(std::vector<PendingOp>& ops) {
    for (auto& o : ops) if (o.is_send) {
        World& w = *o.comm->w;
        std::vector<char> msg((const char*)o.sbuf, (const char*)o.sbuf + o.bytes);
        std::lock_guard<std::mutex> lk(w.m);
        w.p2p[{o.comm->rank, o.peer}].push_back(std::move(msg));
        w.cv.notify_all();
    }
    for (auto& o : ops) if (!o.is_send) {
        World& w = *o.comm->w;
        std::unique_lock<std::mutex> lk(w.m);
        auto key = std::make_pair(o.peer, o.comm->rank);
        if (!w.cv.wait_for(lk, NCCL_TIMEOUT, [&] { return !w.p2p[key].empty(); })) { fprintf(stderr, "cuda_emu nccl: recv from %d timed out\n", o.peer); return ncclSystemError; }
        std::vector<char> msg = std::move(w.p2p[key].front());
        w.p2p[key].pop_front();
        if (msg.size() != o.bytes) { fprintf(stderr, "cuda_emu nccl: recv size mismatch (%zu vs %zu)\n", msg.size(), o.bytes); return ncclInvalidArgument; }
        memcpy(o.rbuf, msg.data(), o.bytes);
    }
    return ncclSuccess;
}
}  // namespace

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
    static std::mutex m; static unsigned long long ctr = 0;
    std::lock_guard<std::mutex> lk(m);
    memset(id, 0, sizeof *id);
    snprintf(id->internal, sizeof id->internal, "emu-nccl-%llu-%ld", ++ctr, (long)time(nullptr));
    return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
    std::string key(id.internal, sizeof id.internal);
    std::shared_ptr<World> w;
    {
        std::lock_guard<std::mutex> lk(g_world_mutex);
        auto& slot = g_worlds[key];
        if (!slot) { slot = std::make_shared<World>(); slot->n = nranks; slot->send.assign(nranks, nullptr); }
        w = slot;
    }
    if (w->n != nranks || rank < 0 || rank >= nranks) return ncclInvalidArgument;
    std::unique_lock<std::mutex> lk(w->m);
    w->joined++; w->alive++;
    w->cv.notify_all();
    if (!w->cv.wait_for(lk, NCCL_TIMEOUT, [&] { return w->joined >= w->n; })) return ncclSystemError;
    *comm = new emu_nccl_comm{w, rank, key};
    return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t comm) {
    if (!comm) return ncclSuccess;
    bool last;
    { std::lock_guard<std::mutex> lk(comm->w->m); last = (--comm->w->alive == 0); }
    if (last) { std::lock_guard<std::mutex> lk(g_world_mutex); g_worlds.erase(comm->key); }
    delete comm;
    return ncclSuccess;
}
ncclResult_t ncclAllReduce(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm, cudaStream_t) {
    World& w = *comm->w;
    size_t bytes = count * dt_size(dt);
    std::vector<char> acc(bytes ? bytes : 1);
    {
        std::unique_lock<std::mutex> lk(w.m);
        w.send[comm->rank] = sendbuff;
        if (!world_barrier(w, lk)) return ncclSystemError;                 // every rank's pointer is published
        std::vector<const void*> src = w.send;
        lk.unlock();
        for (int r = 0; r < w.n; r++) {                                    // rank order: identical result on every rank
            switch (dt) {
                case ncclDouble: reduce_into((double*)acc.data(), (const double*)src[r], count, op, r == 0); break;
                case ncclFloat32: reduce_into((float*)acc.data(), (const float*)src[r], count, op, r == 0); break;
                case ncclInt32: reduce_into((int*)acc.data(), (const int*)src[r], count, op, r == 0); break;
                case ncclUint32: reduce_into((unsigned*)acc.data(), (const unsigned*)src[r], count, op, r == 0); break;
                case ncclInt64: reduce_into((long long*)acc.data(), (const long long*)src[r], count, op, r == 0); break;
                case ncclUint64: reduce_into((unsigned long long*)acc.data(), (const unsigned long long*)src[r], count, op, r == 0); break;
                case ncclInt8: reduce_into((signed char*)acc.data(), (const signed char*)src[r], count, op, r == 0); break;
                default: reduce_into((unsigned char*)acc.data(), (const unsigned char*)src[r], count, op, r == 0); break;
            }
        }
        lk.lock();
        if (!world_barrier(w, lk)) return ncclSystemError;                 // everybody has read every send buffer
    }
    memcpy(recvbuff, acc.data(), bytes);
    return ncclSuccess;
}
ncclResult_t ncclAllGather(const void* sendbuff, void* recvbuff, size_t sendcount, ncclDataType_t dt, ncclComm_t comm, cudaStream_t) {
    World& w = *comm->w;
    const size_t bytes = sendcount * dt_size(dt);
    std::vector<char> all(bytes * w.n ? bytes * w.n : 1);
    {
        std::unique_lock<std::mutex> lk(w.m);
        w.send[comm->rank] = sendbuff;
        if (!world_barrier(w, lk)) return ncclSystemError;                 // every rank's pointer is published
        std::vector<const void*> src = w.send;
        lk.unlock();
        for (int r = 0; r < w.n; r++) memcpy(all.data() + (size_t)r * bytes, src[r], bytes);
        lk.lock();
        if (!world_barrier(w, lk)) return ncclSystemError;                 // everybody has read every send buffer (in-place safe)
    }
    memcpy(recvbuff, all.data(), bytes * w.n);
    return ncclSuccess;
}
ncclResult_t ncclGroupStart(void) { g_group_depth++; return ncclSuccess; }
ncclResult_t ncclGroupEnd(void) {
    if (g_group_depth <= 0) return ncclInvalidUsage;
    if (--g_group_depth > 0) return ncclSuccess;
    std::vector<PendingOp> ops; ops.swap(g_group_ops);
    return run_p2p(ops);
}
ncclResult_t ncclSend(const void* sendbuff, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t) {
    g_group_ops.push_back(PendingOp{true, sendbuff, nullptr, count * dt_size(dt), peer, comm});
    if (g_group_depth == 0) { std::vector<PendingOp> ops; ops.swap(g_group_ops); return run_p2p(ops); }
    return ncclSuccess;
}
ncclResult_t ncclRecv(void* recvbuff, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t) {
    g_group_ops.push_back(PendingOp{false, nullptr, recvbuff, count * dt_size(dt), peer, comm});
    if (g_group_depth == 0) { std::vector<PendingOp> ops; ops.swap(g_group_ops); return run_p2p(ops); }
    return ncclSuccess;
}
const char* ncclGetErrorString(ncclResult_t r) {
    switch (r) { case ncclSuccess: return "no error"; case ncclSystemError: return "emu: rendezvous timed out"; case ncclInvalidArgument: return "emu: invalid argument";
                 case ncclInvalidUsage: return "emu: invalid usage"; default: return "emu: error"; }
}
}  // extern "C"
