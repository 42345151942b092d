// pool.cu - patch node pooling (reference: code/model.py:116, `maps.sum(-1).sum(-1) / (H*W)`), fwd + bwd.
//
// Pure HBM streaming: forward reads rows*hw floats once and writes rows floats; backward the reverse.
// A warp owns 32 consecutive rows per step (8 KB for hw = 64).  Every lane issues L = hw/4 independent
// 128-bit loads (all in flight before the first use), the L lanes of a row then run a butterfly
// reduce-scatter (L-1 shuffles per lane instead of L*log2 L), which leaves row r's sum in a distinct lane,
// so the 32 results leave as one coalesced 128-byte store.  The summation tree is fixed -> deterministic.
#include "common.cuh"

namespace crw {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
#ifdef CRW_SIM
    return *p;
#else
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}

__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
#ifdef CRW_SIM
    *p = v;
#else
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}

// L = lanes per row = hw / 4 (power of two, 4..32)
template <int L, int BLOCK = 256>
__global__ void __launch_bounds__(BLOCK) pool_fwd_kernel(const float4* __restrict__ maps, float* __restrict__ pooled,
                                                       int64_t rows, float inv_hw_div) {
    constexpr int RPL = 32 / L;                     // rows covered by one warp-wide load
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ngroups = (rows + 31) / 32;
    const int sub = lane / L;                       // which of the RPL rows of a load this lane is in
    for (int64_t g = warp; g < ngroups; g += nwarps) {
        const int64_t row0 = g * 32;
        float v[L];
        if (row0 + 32 <= rows) {
            float4 x[L];
#pragma unroll
            for (int i = 0; i < L; ++i) x[i] = ldg_stream(maps + row0 * L + (int64_t)i * 32 + lane);
#pragma unroll
            for (int i = 0; i < L; ++i) v[i] = (x[i].x + x[i].y) + (x[i].z + x[i].w);
        } else {
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const int64_t r = row0 + i * RPL + sub;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < rows) x = ldg_stream(maps + row0 * L + (int64_t)i * 32 + lane);
                v[i] = (x.x + x.y) + (x.z + x.w);
            }
        }
        // butterfly reduce-scatter inside each group of L lanes: value i ends up summed in lane (i mod L)
#pragma unroll
        for (int s = L / 2; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int k = 0; k < s; ++k) {
                const float send = up ? v[k] : v[k + s];
                const float keep = up ? v[k + s] : v[k];
                v[k] = keep + __shfl_xor_sync(kFull, send, s);
            }
        }
        const int64_t r = row0 + (lane & (L - 1)) * RPL + sub;
        if (r < rows) pooled[r] = v[0] / inv_hw_div;
    }
}

template <int L, int BLOCK = 256>
__global__ void __launch_bounds__(BLOCK) pool_bwd_kernel(const float* __restrict__ gpooled, float4* __restrict__ gmaps,
                                                       int64_t rows, float hw_div) {
    constexpr int RPL = 32 / L;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ngroups = (rows + 31) / 32;
    const int sub = lane / L;
    for (int64_t g = warp; g < ngroups; g += nwarps) {
        const int64_t row0 = g * 32;
        float gv = 0.f;
        if (row0 + lane < rows) gv = __ldg(gpooled + row0 + lane) / hw_div;
        const bool full = row0 + 32 <= rows;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const int rl = i * RPL + sub;
            const float b = __shfl_sync(kFull, gv, rl);
            if (full || row0 + rl < rows) stg_stream(gmaps + row0 * L + (int64_t)i * 32 + lane, make_float4(b, b, b, b));
        }
    }
}

// any hw: one warp per row, lanes stride over the row
__global__ void __launch_bounds__(256) pool_fwd_generic(const float* __restrict__ maps, float* __restrict__ pooled,
                                                        int64_t rows, int hw) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        float s = 0.f;
        for (int j = lane; j < hw; j += 32) s += __ldg(maps + r * hw + j);
        s = warp_sum(s);
        if (lane == 0) pooled[r] = s / (float)hw;
    }
}

__global__ void __launch_bounds__(256) pool_bwd_generic(const float* __restrict__ gpooled, float* __restrict__ gmaps,
                                                        int64_t rows, int hw) {
    const int64_t n = rows * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        gmaps[i] = __ldg(gpooled + i / hw) / (float)hw;
}

static int pool_grid(int64_t rows) {
    const int64_t groups = (rows + 31) / 32;
    const int64_t blocks = (groups + 7) / 8;                 // 8 warps per block
    const int64_t cap = 148 * 8;                             // 8 resident 256-thread blocks per SM (6 per SM measured 5 us slower per kernel)
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace crw

using namespace crw;

extern "C" int crw_pool_patch_fwd(const float* maps, float* pooled, int64_t rows, int hw, crw_stream_t stream) {
    if (rows < 0 || hw <= 0) { set_error("pool_patch_fwd: bad shape rows=%lld hw=%d", (long long)rows, hw); return CRW_ERR_SHAPE; }
    if (rows == 0) return CRW_OK;
    const int grid = pool_grid(rows);
    const bool vec = (((uintptr_t)maps) & 15) == 0;
    if (vec && hw == 64) { auto k = pool_fwd_kernel<16>; CRW_LAUNCH(k, grid, 256, 0, stream, (const float4*)maps, pooled, rows, (float)hw); }
    else if (vec && hw == 32) { auto k = pool_fwd_kernel<8>; CRW_LAUNCH(k, grid, 256, 0, stream, (const float4*)maps, pooled, rows, (float)hw); }
    else if (vec && hw == 16) { auto k = pool_fwd_kernel<4>; CRW_LAUNCH(k, grid, 256, 0, stream, (const float4*)maps, pooled, rows, (float)hw); }
    else { CRW_LAUNCH(pool_fwd_generic, grid, 256, 0, stream, maps, pooled, rows, hw); }
    return check_launch("pool_patch_fwd");
}

extern "C" int crw_pool_patch_bwd(const float* gpooled, float* gmaps, int64_t rows, int hw, crw_stream_t stream) {
    if (rows < 0 || hw <= 0) { set_error("pool_patch_bwd: bad shape rows=%lld hw=%d", (long long)rows, hw); return CRW_ERR_SHAPE; }
    if (rows == 0) return CRW_OK;
    const int grid = pool_grid(rows);
    const bool vec = (((uintptr_t)gmaps) & 15) == 0;
    if (vec && hw == 64) { auto k = pool_bwd_kernel<16>; CRW_LAUNCH(k, grid, 256, 0, stream, gpooled, (float4*)gmaps, rows, (float)hw); }
    else if (vec && hw == 32) { auto k = pool_bwd_kernel<8>; CRW_LAUNCH(k, grid, 256, 0, stream, gpooled, (float4*)gmaps, rows, (float)hw); }
    else if (vec && hw == 16) { auto k = pool_bwd_kernel<4>; CRW_LAUNCH(k, grid, 256, 0, stream, gpooled, (float4*)gmaps, rows, (float)hw); }
    else { CRW_LAUNCH(pool_bwd_generic, grid, 256, 0, stream, gpooled, gmaps, rows, hw); }
    return check_launch("pool_patch_bwd");
}
// SM-limited launches (crw_pool_patch_*_sm): `sms` CTAs of 1024 threads, each asking for more than half of an SM's shared
// memory so that exactly one lands on an SM - the other SMs stay completely free for kernels of another stream (a walk CTA
// needs a whole register file).  1024 threads x 16 outstanding 128-bit loads keep 256 KB in flight per SM.
constexpr int kPoolBigBlock = 1024;
constexpr int kPoolReserveSmem = 120 * 1024;

template <typename K, typename... Args>
static void launch_limited(K k, int sms, crw_stream_t stream, Args... args) {
#ifndef CRW_SIM
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoolReserveSmem);
#endif
    CRW_LAUNCH(k, sms, kPoolBigBlock, kPoolReserveSmem, stream, args...);
}

extern "C" int crw_pool_patch_fwd_sm(const float* maps, float* pooled, int64_t rows, int hw, int sms, crw_stream_t stream) {
    if (sms <= 0) return crw_pool_patch_fwd(maps, pooled, rows, hw, stream);
    if (rows < 0 || hw <= 0) { set_error("pool_patch_fwd: bad shape rows=%lld hw=%d", (long long)rows, hw); return CRW_ERR_SHAPE; }
    if (rows == 0) return CRW_OK;
    if ((((uintptr_t)maps) & 15) != 0 || (hw != 64 && hw != 32 && hw != 16)) return crw_pool_patch_fwd(maps, pooled, rows, hw, stream);
    if (hw == 64) launch_limited(pool_fwd_kernel<16, kPoolBigBlock>, sms, stream, (const float4*)maps, pooled, rows, (float)hw);
    else if (hw == 32) launch_limited(pool_fwd_kernel<8, kPoolBigBlock>, sms, stream, (const float4*)maps, pooled, rows, (float)hw);
    else launch_limited(pool_fwd_kernel<4, kPoolBigBlock>, sms, stream, (const float4*)maps, pooled, rows, (float)hw);
    return check_launch("pool_patch_fwd_sm");
}

extern "C" int crw_pool_patch_bwd_sm(const float* gpooled, float* gmaps, int64_t rows, int hw, int sms, crw_stream_t stream) {
    return crw_pool_patch_bwd_scaled(gpooled, gmaps, rows, hw, 1.0f, sms, stream);
}

// gmaps = gpooled * scale / hw: the micro-batch pipeline folds the 1 / n_parts of "mean over all clips" into this pass
// (evaluated as gpooled / (hw / scale): bit-identical to scaling first whenever scale is a power of two)
extern "C" int crw_pool_patch_bwd_scaled(const float* gpooled, float* gmaps, int64_t rows, int hw, float scale, int sms,
                                         crw_stream_t stream) {
    if (rows < 0 || hw <= 0 || !(scale > 0.f)) { set_error("pool_patch_bwd: bad arguments rows=%lld hw=%d scale=%g", (long long)rows, hw, (double)scale); return CRW_ERR_SHAPE; }
    if (rows == 0) return CRW_OK;
    const float div = (float)hw / scale;
    const bool vec = (((uintptr_t)gmaps) & 15) == 0 && (hw == 64 || hw == 32 || hw == 16);
    if (!vec) {
        if (scale != 1.0f) { set_error("pool_patch_bwd_scaled: scale needs 16-byte aligned maps and hw in {16, 32, 64}"); return CRW_ERR_UNSUPPORTED; }
        return crw_pool_patch_bwd(gpooled, gmaps, rows, hw, stream);
    }
    if (sms <= 0) {
        const int grid = pool_grid(rows);
        if (hw == 64) { auto k = pool_bwd_kernel<16>; CRW_LAUNCH(k, grid, 256, 0, stream, gpooled, (float4*)gmaps, rows, div); }
        else if (hw == 32) { auto k = pool_bwd_kernel<8>; CRW_LAUNCH(k, grid, 256, 0, stream, gpooled, (float4*)gmaps, rows, div); }
        else { auto k = pool_bwd_kernel<4>; CRW_LAUNCH(k, grid, 256, 0, stream, gpooled, (float4*)gmaps, rows, div); }
        return check_launch("pool_patch_bwd_scaled");
    }
    if (hw == 64) launch_limited(pool_bwd_kernel<16, kPoolBigBlock>, sms, stream, gpooled, (float4*)gmaps, rows, div);
    else if (hw == 32) launch_limited(pool_bwd_kernel<8, kPoolBigBlock>, sms, stream, gpooled, (float4*)gmaps, rows, div);
    else launch_limited(pool_bwd_kernel<4, kPoolBigBlock>, sms, stream, gpooled, (float4*)gmaps, rows, div);
    return check_launch("pool_patch_bwd_sm");
}
