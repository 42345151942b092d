// cuda_sim.h - a tiny host emulator of the CUDA execution model.  TEST INFRASTRUCTURE ONLY.
//
// The kernels under sapienza_video_contrastive_b200/csrc are written for sm_100a and are only ever
// shipped as such.  The authoring container has no GPU, so to check kernel *logic* (indexing, barriers,
// shuffles, workspace layout, the C-ABI host code) before spending GPU minutes, tests/cusim/build_sim.py
// compiles the same .cu sources with g++ and this header force-included: every CUDA thread of a block
// becomes a ucontext fiber, __syncthreads / __shfl*_sync / __syncwarp become cooperative yields, and
// blocks run one after another.  The resulting libcrw_b200_sim.so is loaded ONLY by
// tests/test_sim_kernels.py (with CPU pointers).  The product package never loads it and has no CPU path.
// Kernels that use TMA / tcgen05 inline PTX are compiled out under CRW_SIM.
#pragma once
#ifndef CRW_SIM
#error "cuda_sim.h is only for -DCRW_SIM host builds"
#endif

#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __align__(16) float4 { float x, y, z, w; };
struct __align__(8) float2 { float x, y; };
struct __align__(16) int4 { int x, y, z, w; };
struct __align__(16) uint4 { unsigned x, y, z, w; };
struct __align__(8) uint2 { unsigned x, y; };
struct __align__(8) int2 { int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "sim"; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxThreadsPerMultiProcessor = 39, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = a == cudaDevAttrMultiProcessorCount ? 148 : a == cudaDevAttrMaxThreadsPerMultiProcessor ? 2048 : 232448;
    return cudaSuccess;
}

namespace cusim {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    uint3 tid;
    int lin = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    int nthreads = 0, alive = 0;
    int bar_count = 0;
    unsigned long bar_gen = 0;
    // per-warp state
    std::vector<int> warp_count;
    std::vector<unsigned long> warp_gen;
    std::vector<uint64_t> xchg;           // one 64-bit slot per thread
    unsigned char* smem = nullptr;
    int cur = 0;
    ucontext_t sched;
    const std::function<void()>* body = nullptr;
};

extern Block* g_blk;
extern size_t g_stack_bytes;
void yield();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
unsigned char* dyn_smem();
void block_barrier();
void warp_barrier(unsigned mask);
int lane_id();
int warp_id();
uint64_t* warp_slots();
unsigned active_mask();

}  // namespace cusim

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
static const int warpSize = 32;

static inline void __syncthreads() { cusim::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cusim::warp_barrier(mask); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned __activemask() { return cusim::active_mask(); }

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    uint64_t* s = cusim::warp_slots();
    int lane = cusim::lane_id();
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    s[lane] = raw;
    cusim::warp_barrier(mask);
    int base = lane & ~(width - 1);
    int from = base + (src & (width - 1));
    uint64_t got = s[from];
    cusim::warp_barrier(mask);
    T out; memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32) {
    int lane = cusim::lane_id();
    int from = lane ^ lm;
    if ((from & ~(width - 1)) != (lane & ~(width - 1))) from = lane;
    return __shfl_sync(mask, v, from & 31, 32);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int lane = cusim::lane_id();
    int from = lane + (int)d;
    if ((from & ~(width - 1)) != (lane & ~(width - 1))) from = lane;
    return __shfl_sync(mask, v, from & 31, 32);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int lane = cusim::lane_id();
    int from = lane - (int)d;
    if (from < (lane & ~(width - 1))) from = lane;
    return __shfl_sync(mask, v, from & 31, 32);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint64_t* s = cusim::warp_slots();
    int lane = cusim::lane_id();
    s[lane] = pred ? 1 : 0;
    cusim::warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if (((mask >> l) & 1u) && s[l]) r |= 1u << l;
    cusim::warp_barrier(mask);
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == (mask & cusim::active_mask()); }

// atomics: fibers are cooperatively scheduled on one OS thread, so plain RMW is atomic
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o | v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; *p = o > v ? o : v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; *p = o > v ? o : v; return o; }
static inline int atomicMin(int* p, int v) { int o = *p; *p = o < v ? o : v; return o; }
static inline int atomicCAS(int* p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }
static inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }

// math / intrinsics
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31u; return sh ? (lo >> sh) | (hi << (32u - sh)) : lo; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::min;
using std::max;
static inline float fminf_(float a, float b) { return a < b ? a : b; }
