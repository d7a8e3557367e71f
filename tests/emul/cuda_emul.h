// cuda_emul.h -- TEST INFRASTRUCTURE ONLY.  A tiny single-process SIMT emulator that lets the *same* kernel
// sources (nafcodec_b200/csrc/*.cu) be compiled with g++ and executed on the CPU, one CTA at a time, each CUDA
// thread a ucontext fiber, with real __syncthreads/__syncwarp/shuffle semantics (and deadlock detection).
//
// It exists so the logic of the kernels can be checked in the CPU-only test tier (`-m "not gpu"`) and debugged
// without a GPU round trip.  It is built into tests/emul/_build/libnafgpu_emul.so, loaded ONLY by tests/, and is
// never reachable from the product package: nafcodec_b200 loads nafcodec_b200/csrc/libnafgpu.so (nvcc, sm_100a)
// and fails loudly when that library or a CUDA device is missing.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct ulonglong2 { unsigned long long x, y; } __attribute__((aligned(16)));
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }

namespace emul {
struct Idx { unsigned x, y, z; };
extern Idx g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
void sync_cta();
void sync_warp();
uint64_t warp_exchange(uint64_t v, int src_lane);      // value of src_lane (all live lanes of the warp must call)
uint32_t warp_ballot(int pred);
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
extern int g_vote_slot[3];
int vote_enter();
}  // namespace emul

#define threadIdx emul::g_threadIdx
#define blockIdx emul::g_blockIdx
#define blockDim emul::g_blockDim
#define gridDim emul::g_gridDim
#define warpSize 32

static inline void __syncthreads() { emul::sync_cta(); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emul::sync_warp(); }
int __syncthreads_or(int pred);
int __syncthreads_and(int pred);
int __syncthreads_count(int pred);
static inline void __threadfence() {}
static inline void __threadfence_block() {}

static inline int emul_lane() { return (int)((threadIdx.x + threadIdx.y * blockDim.x) & 31); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); raw = emul::warp_exchange(raw, src & 31); T r; memcpy(&r, &raw, sizeof(T)); return r;
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int = 32) {
    int l = emul_lane(); int src = l - (int)d; T r = __shfl_sync(m, v, src < 0 ? l : src); return src < 0 ? v : r;
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int = 32) {
    int l = emul_lane(); int src = l + (int)d; T r = __shfl_sync(m, v, src > 31 ? l : src); return src > 31 ? v : r;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int = 32) { return __shfl_sync(m, v, emul_lane() ^ x); }
static inline unsigned __ballot_sync(unsigned, int p) { return emul::warp_ballot(p); }
template <class T> static inline unsigned __match_any_sync(unsigned m, T v) {       // lanes holding the same value (all 32 lanes must call)
    unsigned r = 0;
    for (int l = 0; l < 32; l++) { T x = __shfl_sync(m, v, l); if (x == v) r |= 1u << l; }
    return r;
}
static inline int __any_sync(unsigned, int p) { return emul::warp_ballot(p) != 0; }
static inline int __all_sync(unsigned m, int p) { return emul::warp_ballot(!p) == 0; (void)m; }
static inline unsigned __activemask() { return emul::warp_ballot(1); }

static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xF; unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }

template <class T> struct emul_id { typedef T type; };
#define EMUL_V typename emul_id<T>::type
template <class T> static inline T atomicAdd(T* p, EMUL_V v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T* p, EMUL_V v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicXor(T* p, EMUL_V v) { T o = *p; *p = o ^ v; return o; }
template <class T> static inline T atomicAnd(T* p, EMUL_V v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicMin(T* p, EMUL_V v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, EMUL_V v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, EMUL_V v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, EMUL_V c, EMUL_V v) { T o = *p; if (o == c) *p = v; return o; }

// ---- runtime subset ---------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct emul_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDefault = 0 };
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = 0; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
cudaError_t cudaEventCreate(cudaEvent_t* e);
enum { cudaEventDisableTiming = 2, cudaEventBlockingSync = 1 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
