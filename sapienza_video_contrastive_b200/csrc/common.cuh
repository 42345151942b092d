// common.cuh - shared helpers for the crw_b200 kernels (sm_100a).
#pragma once

#include <stddef.h>
#include <stdint.h>

#ifdef CRW_SIM
// host emulation build used only by tests/cusim (cuda_sim.h is force-included there)
#define CRW_LAUNCH(kern, grid, block, smem, stream, ...) \
    cusim::launch(grid, block, smem, [&]() { kern(__VA_ARGS__); })
#define CRW_DYN_SMEM(name) unsigned char* name = cusim::dyn_smem()
#else
#include <cuda_runtime.h>
#define CRW_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define CRW_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

#include "../../include/crw_b200.h"

namespace crw {

constexpr float kEpsLog = 1e-20f;    // model.py:12
constexpr float kEpsZs = 1e-5f;      // utils/__init__.py:418
constexpr float kNegDrop = -1e20f;   // model.py:81
constexpr float kEpsNorm = 1e-12f;   // F.normalize default eps, model.py:118
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char* fmt, ...);
int check_launch(const char* what);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

__device__ __forceinline__ float ld_cg(const float* p) {     // L2-coherent load (data written by other CTAs)
#ifdef CRW_SIM
    return *p;
#else
    return __ldcg(p);
#endif
}

__device__ __forceinline__ uint64_t ld_cg64(const uint64_t* p) {
#ifdef CRW_SIM
    return *p;
#else
    return (uint64_t)__ldcg(reinterpret_cast<const unsigned long long*>(p));
#endif
}

// ---- Philox4x32-10, bit-compatible with curand / torch's CUDA generator -----------------------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// Uniform in [0,1) for flat element `e` of a tensor filled by torch.rand / rand_like on CUDA:
// torch launches `threads` = grid*256 threads; thread idx = e % threads draws 4 values per step from
// Philox(seed; counter = offset/4 + step, subsequence = idx) and element e takes component (e/threads)%4 of
// step (e/threads)/4.  (ATen/native/cuda/DistributionTemplates.h; curand_uniform4 then maps u32 -> (0,1]
// and torch folds 1.0 back to 0.)
__host__ __device__ __forceinline__ float torch_uniform(uint64_t seed, uint64_t offset, uint32_t threads, uint64_t e) {
    uint64_t idx = e, k = 0;
    if (e >= threads) {                    // (the common case e < threads avoids the 64-bit division)
        if (e < 0xffffffffull) { idx = (uint32_t)e % threads; k = (uint32_t)e / threads; }
        else { idx = e % threads; k = e / threads; }
    }
    const uint64_t ctr = (offset >> 2) + (k >> 2);
    const Philox4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)idx, (uint32_t)(idx >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t c = (uint32_t)(k & 3);
    const uint32_t bits = c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w;
#if defined(__CUDA_ARCH__)
    float u = __fmaf_rn((float)bits, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
#else
    float u = fmaf((float)bits, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
#endif
    return u == 1.0f ? 0.0f : u;
}

// torch's launch geometry for a rand over `numel` elements (block 256, unroll 4).
__host__ __forceinline__ uint32_t torch_rand_threads(int64_t numel, int sm_count, int max_threads_per_sm) {
    int64_t grid = (numel + 255) / 256;
    const int64_t cap = (int64_t)sm_count * (max_threads_per_sm / 256);
    if (grid > cap) grid = cap;
    return (uint32_t)(grid * 256);
}

// packed fp32 FMA (Blackwell FFMA2): two independent IEEE fused multiply-adds per instruction, scalar x vector or vector x vector
__device__ __forceinline__ float2 ffma2(float a, float2 x, float2 acc) {
#ifdef CRW_SIM
    return make_float2(fmaf(a, x.x, acc.x), fmaf(a, x.y, acc.y));
#else
    return __ffma2_rn(make_float2(a, a), x, acc);
#endif
}
__device__ __forceinline__ float2 ffma2v(float2 a, float2 x, float2 acc) {
#ifdef CRW_SIM
    return make_float2(fmaf(a.x, x.x, acc.x), fmaf(a.y, x.y, acc.y));
#else
    return __ffma2_rn(a, x, acc);
#endif
}

// global -> shared TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: thread 0 announces the total
// byte count, issues the copies, and the whole CTA waits on the barrier.  Sizes are multiples of 16, both sides aligned.
__device__ __forceinline__ void bulk_bar_init(uint64_t* bar, int tid) {
#ifndef CRW_SIM
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#else
    (void)bar; (void)tid;
#endif
    __syncthreads();
}
__device__ __forceinline__ void bulk_expect(uint64_t* bar, unsigned total_bytes, int tid) {
#ifndef CRW_SIM
    if (tid == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(total_bytes) : "memory");
#else
    (void)bar; (void)total_bytes; (void)tid;
#endif
}
__device__ __forceinline__ void bulk_copy(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar, int tid) {
#ifdef CRW_SIM
    for (unsigned o = tid * 16; o < bytes; o += blockDim.x * 16)
        *reinterpret_cast<float4*>((char*)dst_smem + o) = *reinterpret_cast<const float4*>((const char*)src_gmem + o);
    (void)bar;
#else
    if (tid == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem), "r"(bytes),
                        "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#endif
}
__device__ __forceinline__ void bulk_wait(uint64_t* bar, unsigned phase) {
#ifdef CRW_SIM
    (void)bar; (void)phase;
    __syncthreads();
#else
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tCRW_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra CRW_DONE_%=;\n\t"
                 "bra CRW_WAIT_%=;\n\tCRW_DONE_%=:\n\t}" :: "r"(b), "r"(phase) : "memory");
#endif
}

}  // namespace crw
