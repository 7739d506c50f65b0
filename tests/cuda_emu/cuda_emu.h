// cuda_emu.h — TEST INFRASTRUCTURE, not a product path.
//
// A host-side emulation of the small CUDA subset libtopopt_b200's kernels use, so that the *logic* of the very same .cu
// sources (indexing, initialisation, barriers, mbarrier pipeline, reductions, CG recurrences, partition maps) can be
// exercised on a machine without a GPU:  g++ -x c++ -DTOE_EMU -include cuda_emu.h  csrc/*.cu  →  libtopopt_emu.so.
//
//   * every CUDA thread of a block is a fiber (own stack, hand-written context switch); blocks run one after another
//   * __syncthreads / named barriers / warp shuffles / mbarrier waits are cooperative yields
//   * cudaMalloc returns memory filled with 0xFF (NaN doubles, -1 ints): a read-before-write shows up in the results
//     instead of being hidden by the zero pages a fresh CUDA allocation usually hands out
//   * streams execute immediately, events read the wall clock, CUDA graphs replay recorded launches (arguments baked in)
//   * NCCL is replaced by an in-process rendezvous between rank *threads* (one ctx per thread)
//
// Only tests/ load the resulting library (tests/emu_support.py); the package never does: the product path still fails
// loudly without a B200.  Nothing measured or shipped comes from here.
#pragma once
#ifndef TOE_EMU
#error "cuda_emu.h is only for the TOE_EMU test build"
#endif

#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <functional>
#include <tuple>
#include <utility>

// ---- qualifiers ---------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) alignas(n)
#define __constant__ static

// ---- built-in variables -----------------------------------------------------------------------------------------
struct emu_dim3 { unsigned x, y, z; };
extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

// ---- vector types -------------------------------------------------------------------------------------------------
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- runtime API (subset) --------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801,
       cudaErrorStreamCaptureUnsupported = 900 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
typedef struct emu_graph* cudaGraph_t;
typedef struct emu_graph* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { char name[256]; int major, minor; int multiProcessorCount; };

extern "C" {
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
// stream-ordered allocation: the emulated stream executes in program order, so these are the plain calls
typedef struct emu_mempool* cudaMemPool_t;
enum { cudaMemPoolAttrReleaseThreshold = 4 };
static inline cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void*) { return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t bytes);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemset(void* dst, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaGetLastError(void);
cudaError_t cudaPeekAtLastError(void);
const char* cudaGetErrorName(cudaError_t e);
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode mode);
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* ge, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g);
cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t s);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
}
extern "C" void emu_set_max_dyn_smem(const void* fn, int bytes);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F f, int attr, int value) {
    if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize) emu_set_max_dyn_smem(reinterpret_cast<const void*>(f), value);
    return cudaSuccess;
}

// ---- kernel launch ---------------------------------------------------------------------------------------------------
namespace emu {
void launch(unsigned grid, unsigned block, size_t smem, std::function<void()> body, const void* fn = nullptr);   // runs (or records, during capture) the grid
void* dyn_smem();                          // dynamic shared memory of the running block
void yield();                              // cooperative reschedule (used by spin loops)
void cp_async(void* dst, const void* src, unsigned bytes);   // cp.async modelled at its LATEST legal completion: performed at the issuing thread's wait
void cp_async_wait_all();
void next_launch_coresident();             // the next launch runs all its blocks at once (software grid barrier inside)
void sync_threads();
void named_barrier(int id, int count);
unsigned long long shfl(unsigned long long v, int src_lane);   // warp exchange: returns the value `src_lane` deposited
long long clock();
}

// ---- device intrinsics ----------------------------------------------------------------------------------------------
static inline void __syncthreads() { emu::sync_threads(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu::shfl(0ULL, (int)(threadIdx.x & 31)); }     // a warp-wide rendezvous (full-mask use only)
static inline long long clock64() { return emu::clock(); }
extern "C" void emu_misaligned(const void* p, unsigned bytes);
template <class T> static inline T __ldg(const T* p) {
    if (sizeof(T) >= 8 && ((size_t)p % (sizeof(T) > 16 ? 16 : sizeof(T)))) emu_misaligned(p, (unsigned)sizeof(T));   // vector loads fault on a GPU when misaligned
    return *p;
}
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcv(const T* p) { return *(const volatile T*)p; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

template <class T> static inline T emu_shfl_any(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
    unsigned long long b = 0; memcpy(&b, &v, sizeof(T));
    b = emu::shfl(b, src);
    T r; memcpy(&r, &b, sizeof(T)); return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int mask) { return emu_shfl_any(v, (int)((threadIdx.x & 31) ^ (unsigned)mask)); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    int lane = (int)(threadIdx.x & 31);
    int src = lane - (int)delta;
    return emu_shfl_any(v, src < 0 ? lane : src);       // lanes without a source keep their own value
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_any(v, src & 31); }

// atomics: one OS thread runs all fibers of a ctx, so plain read-modify-write is atomic by construction
template <class T> static inline T emu_atomic_add(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { return emu_atomic_add(p, v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return emu_atomic_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return emu_atomic_add(p, v); }
static inline double atomicAdd(double* p, double v) { return emu_atomic_add(p, v); }
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }

// CUDA's global-namespace min/max
template <class T> static inline T min(T a, T b) { return b < a ? b : a; }
template <class T> static inline T max(T a, T b) { return a < b ? b : a; }

// ---- mbarrier / bulk-copy emulation (solver.cu's pipeline helpers call these under TOE_EMU) -----------------------------
namespace emu {
// state packed into the kernel's own 8-byte barrier word: phase | expected arrivals | pending arrivals | pending tx bytes
void mbar_init(void* bar, unsigned count);
void mbar_expect_tx(void* bar, unsigned bytes);
void mbar_arrive(void* bar);
void mbar_wait(void* bar, unsigned parity);
void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar);
}
