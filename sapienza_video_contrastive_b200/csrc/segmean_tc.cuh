// segmean_tc.cuh - the superpixel pooling forward as ONE tensor-core kernel (included by segmean.cu, after SegWs).
//
// out[b, s, t, :] = sum over cells of count[s][cell] * maps[b, :, t, cell] / (size[s] + eps)          (model.py:311-325)
// is a product of the (labels x cells) window-count matrix - sparse, but small: <= 256 x 1024 - with the (cells x channels)
// feature map of the frame.  The SIMT scatter-reduce (segmean_csr + segmean_accum) pays per non-zero count (one reduction
// chain per (label, cell) entry: 74 us plain, 214 us with dilated masks at the BASELINE configs[2] shape, 0.27 of HBM at best);
// here a CTA streams its frame's map through TMA exactly once and the count matrix is never stored anywhere but in the
// shared-memory operand tile of the current 32 cells:
//   warp 0      TMA producer: box (32 cells, 1 frame, NCH channels) of the (cells, T, B*C) view of the maps, 128B-swizzled
//               = the K-major operand tile [channel][cell] as tcgen05 wants it (the M side: 128 channels per MMA)
//   warps 2-17  builders: (a) scatter the (label, count) lists of the slab's cells (written by segmean_count / segdil_count)
//               into the count tile [label][cell] (the N side) - counts are small integers, exact in TF32; only the label rows
//               [lo, hi) the slab touches are zeroed, written and multiplied: label ids follow the scan order of their segments
//               (SLIC with enforce_connectivity numbers them so), a row of cells meets a few dozen of them, and the MMA's N
//               shrinks accordingly (any label order stays correct, the range is just wider);
//               (b) write small = tf32(x - big) of the landed map tile into a twin tile; big = x truncated to TF32 is what the
//               tensor core reads from the fp32 words as they are (x = big + small to 2^-22 relative)
//   warp 1      MMA issuer: per 8-cell k-step and 128-channel tile  D[:, lo:hi) += big.counts, += small.counts  (kind::tf32,
//               fp32 accumulators in TMEM, zeroed up front: lane = channel, column = label)
//   epilogue    (the builder warps) tcgen05.ld, divide by the label's size, one 128-byte row segment of out (B, SP, T, C) per
//               warp store
// Shared memory: a ring of 3-6 landed map tiles (NCH x 128 B each; the TMA producer runs that far ahead), two small tiles and two
// count tiles (the builders work one slab ahead of the tensor core).  Results differ from the SIMT path
// only in summation order (|err| <= 2^-22 |x| per term + fp32 accumulation); both paths are deterministic.
#pragma once
#ifndef CRW_SIM

namespace crw {

constexpr int SG_K = 32, SG_BUILDERS = 512, SG_EG = SG_BUILDERS / 32, SG_PRE = 32 / SG_EG, SG_THREADS = 64 + SG_BUILDERS;

constexpr int SG_MAX_RAW = 8;                          // TMA ring depth (as many as fit beside the other tiles)
constexpr size_t kSgMisc = 4096, kSgSmemMax = 227 * 1024;

template <int NCH> struct SgCfg {
    static constexpr unsigned kB = NCH * 128u;         // one map tile: NCH channel rows x 128 B
};
struct SgLayout { int raw_stages; unsigned cnt_bytes; size_t smem; };
template <int NCH> static SgLayout sg_layout(int SP) {
    SgLayout L;
    L.cnt_bytes = (unsigned)((SP + 15) & ~15) * 128u;                  // count tile: the label rows that exist
    const size_t fixed = 1024 + kSgMisc + 2 * (size_t)SgCfg<NCH>::kB + 2 * (size_t)L.cnt_bytes;
    int r = (int)((kSgSmemMax - fixed) / SgCfg<NCH>::kB);
    L.raw_stages = r > SG_MAX_RAW ? SG_MAX_RAW : r;
    L.smem = fixed + (size_t)L.raw_stages * SgCfg<NCH>::kB;
    return L;
}

__device__ unsigned g_sg_err;                          // mbar_wait's error word (the wait also traps)

__device__ __forceinline__ void sg_mma_tf32(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float sg_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float sg_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <int NCH>
__global__ void __launch_bounds__(SG_THREADS, 1) segmean_mma_kernel(const __grid_constant__ CUtensorMap fmap, SegWs ws, int C, int T, int cells,
                                                                    int SP, int R, unsigned cnt_bytes, int dbg, float* __restrict__ out) {
    using Cfg = SgCfg<NCH>;
    constexpr int kTiles = NCH / 128;                  // 128-channel M-tiles
    constexpr unsigned kCols = 256 * kTiles;           // TMEM columns: 256 labels per M-tile
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);     // stays in the shared state space: LDS / STS, not generic LD / ST
    // raw[R] (the landed fp32 map tiles = the "big" operand) | small[2] | counts[2] | barriers and tables
    unsigned char* small_t = smem + (size_t)R * Cfg::kB;
    unsigned char* cnt_t = small_t + 2 * Cfg::kB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(cnt_t + 2 * (size_t)cnt_bytes);
    uint64_t* full = bars;                             // [R] TMA landed
    uint64_t* rawfree = bars + SG_MAX_RAW;             // [R] the MMAs that read raw[r] retired
    uint64_t* built = bars + 2 * SG_MAX_RAW;           // [2] builders done with slab parity c
    uint64_t* mdone = built + 2;                       // [2] the MMAs of slab parity c retired (counts[c] and small[c] are free)
    uint64_t* done = mdone + 2;
    unsigned* tmem_base_smem = reinterpret_cast<unsigned*>(done + 1);
    int* rng = reinterpret_cast<int*>(bars) + 64;      // [parity]{first label row, number of rows} for the MMA warp
    int* racc = rng + 16;                              // three rotating {min label, max label} pairs of the builders
    float* dtab = reinterpret_cast<float*>(racc + 16);              // [256] size + eps
    float* rtab = dtab + 256;                          // [256] its reciprocal
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bt = blockIdx.y, b = bt / T, t = bt - b * T, n0 = blockIdx.x * NCH;
    const int nslabs = cells / SG_K;
    // every CTA walks the slabs from its own starting point (the sum over cells has no order): CTAs in lockstep on the same
    // slab would all ask for addresses with identical low bits, 32 KB apart
    const int rot = (int)((blockIdx.y * 7u + blockIdx.x * 3u) % (unsigned)nslabs);
    auto slab_of = [&](int it) { const int j = it + rot; return j >= nslabs ? j - nslabs : j; };

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&fmap)) : "memory");
        for (int s = 0; s < R; ++s) { mbar_init(full + s, 1); mbar_init(rawfree + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(built + s, SG_BUILDERS); mbar_init(mdone + s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_base_smem)), "r"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer: runs up to R slabs ahead, in bursts of G = R / 2 =====================
        // (a box row is 128 B of a 4 KB channel row: neighbouring slabs requested together fall into the same DRAM pages)
        const int G = R >= 2 ? R / 2 : 1;
        bool ok = true;
        for (int g0 = 0; ok && g0 < nslabs; g0 += G) {
            const int g1 = g0 + G < nslabs ? g0 + G : nslabs;
            for (int it = g0; ok && it < g1; ++it) ok = mbar_wait(rawfree + it % R, ((unsigned)(it / R) & 1u) ^ 1u, &g_sg_err);
            if (!ok) break;
            if (lane == 0) {
                for (int it = g0; it < g1; ++it) {
                    const int r = it % R;
                    mbar_expect_tx(full + r, Cfg::kB);
                    tma_load_3d(smem + (size_t)r * Cfg::kB, &fmap, slab_of(it) * SG_K, t, b * C + n0, full + r);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const unsigned tb = __shfl_sync(kFull, tmem_base, 0);
        const unsigned sbase = __shfl_sync(kFull, smem_u32(smem), 0);
        const unsigned ssmall0 = sbase + (unsigned)R * Cfg::kB, scnt = ssmall0 + 2 * Cfg::kB;
        // kind::tf32: D fp32, A / B tf32, both K-major, M = 128; N (the label rows of the slab) is filled in per slab
        const unsigned idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(128 >> 4) << 24);
        bool ok = true;
        for (int it = 0; ok && it < nslabs; ++it) {
            const int c = it & 1, r = it % R;
            ok = mbar_wait(built + c, (unsigned)(it >> 1) & 1u, &g_sg_err);
            if (!ok) break;
            tc_fence_after();
            const int lo = rng[2 * c], n = rng[2 * c + 1];
            if (elect_one()) {
                if (n > 0 && !(dbg & 1)) {
                    const unsigned idesc = idesc0 | ((unsigned)(n >> 3) << 17);
                    const unsigned mb = sbase + (unsigned)r * Cfg::kB, ssmall = ssmall0 + (unsigned)c * Cfg::kB, cn = scnt + (unsigned)c * cnt_bytes + (unsigned)lo * 128u;
#pragma unroll
                    for (int kk = 0; kk < SG_K / 8; ++kk) {
                        const uint64_t dc = umma_desc_sw128(cn + 32u * kk);
#pragma unroll
                        for (int m = 0; m < kTiles; ++m) {
                            const unsigned d = tb + (unsigned)(m * 256 + lo);
                            sg_mma_tf32(d, umma_desc_sw128(mb + 16384u * m + 32u * kk), dc, idesc, 1u);
                            sg_mma_tf32(d, umma_desc_sw128(ssmall + 16384u * m + 32u * kk), dc, idesc, 1u);
                        }
                    }
                }
                tc_commit(rawfree + r);
                tc_commit(mdone + c);
                if (it == nslabs - 1) tc_commit(done);
            }
            __syncwarp();
        }
    } else {
        // ===================== builders (warps 2-17), then the epilogue =====================
        const int bid = threadIdx.x - 64;                                 // 0..SG_BUILDERS-1
        const int k = bid & 31, eg = bid >> 5;                            // this thread's cell of the slab, its entry slots eg, eg + SG_EG, ...
        const unsigned char* nent = ws.nent + (int64_t)bt * cells;
        const unsigned* ent = ws.ent + (int64_t)bt * ws.cap * cells;
        const int cap = ws.cap;
        // epilogue geometry (also used to zero the accumulators now): four warps per TMEM lane quarter share the M-tiles and labels
        const int quarter = warp & 3, idx = (warp - 2) >> 2;              // idx 0..3
        const int tile = kTiles == 2 ? (idx & 1) : 0;
        const int lw = kTiles == 2 ? 128 : 64, lbeg = (kTiles == 2 ? (idx >> 1) : idx) * lw;
        const unsigned lane_addr = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(tile * 256);
        {
            unsigned z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0u;
            for (int s0 = lbeg; s0 < lbeg + lw; s0 += 32) tc_st32(lane_addr + (unsigned)s0, z);
            tc_wait_st();
            tc_fence_before();
        }
        if (bid < 256) {
            float d = 1.f, r = 1.f;
            if (bid < SP) {
                d = (float)__ldg(ws.size + (int64_t)bt * SP + bid) + kEpsLog;     // d >= 1, or 1e-20 for an empty label whose sums are 0
                r = 1.0f / d;
                r = fmaf(fmaf(-d, r, 1.0f), r, r);
            }
            dtab[bid] = d;
            rtab[bid] = r;
        }
        // lists of cell k of slab j: nent_p[32 j], entries eg and eg + SG_EG at ent_p[32 j] and ent_p[e2 + 32 j]
        const unsigned char* nent_p = nent + k;
        const unsigned* ent_p = ent + (int64_t)eg * cells + k;
        const int64_t e2 = (int64_t)SG_EG * cells;
        const bool has0 = eg < cap, has1 = eg + SG_EG < cap;
        static_assert(SG_PRE == 2, "two prefetched entries per thread");
        auto load_lists = [&](int it_, int& ne_, unsigned& a_, unsigned& b_) {
            const int slab = slab_of(it_);
            ne_ = nent_p[slab * SG_K];
            a_ = has0 ? __ldg(ent_p + slab * SG_K) : 0u;
            b_ = has1 ? __ldg(ent_p + e2 + slab * SG_K) : 0u;
        };
        // the label rows a slab touches: warp reduction, then shared-memory atomics into one of three rotating {min, max} pairs
        // (published before one barrier, read after it, reset an iteration later); a cell with more entries than are prefetched
        // widens the range to every label
        auto publish_range = [&](int buf, int ne_, unsigned a_, unsigned b_) {
            int lo = 1 << 20, hi = -1;
            if (eg < ne_) { lo = (int)(a_ >> 8); hi = lo; }
            if (eg + SG_EG < ne_) { const int s2 = (int)(b_ >> 8); lo = min(lo, s2); hi = max(hi, s2); }
            if (ne_ > 32) { lo = 0; hi = SP - 1; }
            lo = __reduce_min_sync(kFull, lo);
            hi = __reduce_max_sync(kFull, hi);
            if (lane == 0 && hi >= 0) { atomicMin(racc + 2 * buf, lo); atomicMax(racc + 2 * buf + 1, hi); }
        };
        int ne = 0, ne1 = 0, ne2 = 0;
        unsigned va = 0u, vb = 0u, va1 = 0u, vb1 = 0u, va2 = 0u, vb2 = 0u;
        load_lists(0, ne, va, vb);
        if (nslabs > 1) load_lists(1, ne1, va1, vb1);
        if (bid < 3) { racc[2 * bid] = 1 << 20; racc[2 * bid + 1] = -1; }
        asm volatile("bar.sync 1, %0;" :: "n"(SG_BUILDERS) : "memory");
        publish_range(0, ne, va, vb);
        asm volatile("bar.sync 1, %0;" :: "n"(SG_BUILDERS) : "memory");
        bool ok = true;
        int r = 0, b0 = 0, b1 = 1, b2 = 2;                                // ring slot of the slab; range buffers of slabs it, it + 1, (free)
        unsigned rph = 0u;
        unsigned char* raw = smem;
        for (int it = 0; ok && it < nslabs; ++it) {
            const int c = it & 1;
            if (it + 2 < nslabs) load_lists(it + 2, ne2, va2, vb2);       // travels for two slabs
            if (bid == 0) { racc[2 * b2] = 1 << 20; racc[2 * b2 + 1] = -1; }
            if (it + 1 < nslabs) publish_range(b1, ne1, va1, vb1);        // read after the NEXT barrier
            const int lo = racc[2 * b0], hi = racc[2 * b0 + 1];
            const int lo16 = hi >= 0 ? (lo & ~15) : 0, n16 = hi >= 0 ? (((hi + 16) & ~15) - lo16) : 0;
            if (it >= 2) {                                                // counts[c] and small[c] were read by the MMAs of slab it - 2
                ok = mbar_wait(mdone + c, (unsigned)((it - 2) >> 1) & 1u, &g_sg_err);
                if (!ok) break;
            }
            unsigned char* cnt = cnt_t + (size_t)c * cnt_bytes;
            {
                float4* c4 = reinterpret_cast<float4*>(cnt) + lo16 * 8;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = bid; i < n16 * 8; i += SG_BUILDERS) c4[i] = z;
            }
            asm volatile("bar.sync 1, %0;" :: "n"(SG_BUILDERS) : "memory");
            {
                // element (label s, cell k) of the K-major 128B-swizzled tile: 8-row groups of 1024 B, 16-byte chunk k / 4 XOR-ed
                // with the row's position in its group
                auto put = [&](unsigned x) {
                    const unsigned s = x >> 8;
                    const unsigned off = (s >> 3) * 1024u + (s & 7u) * 128u + ((((unsigned)k >> 2) ^ (s & 7u)) << 4) + ((unsigned)k & 3u) * 4u;
                    *reinterpret_cast<float*>(cnt + off) = (float)(x & 255u);
                };
                if (eg < ne && !(dbg & 4)) put(va);
                if (eg + SG_EG < ne && !(dbg & 4)) put(vb);
                if (ne > 32) {
                    const int cell = slab_of(it) * SG_K + k;
                    for (int e = eg + 32; e < ne; e += SG_EG) put(__ldg(ent + (int64_t)e * cells + cell));
                }
            }
            ok = mbar_wait(full + r, rph, &g_sg_err);
            if (!ok) break;
            if (!(dbg & 2)) {
                // the landed tile is used as it is: kind::tf32 reads the upper 19 bits of an fp32 word (big = the truncated value);
                // only small = tf32(x - big) is written, into a tile with the identical layout
                const float4* big = reinterpret_cast<const float4*>(raw) + bid;
                float4* small = reinterpret_cast<float4*>(small_t + (size_t)c * Cfg::kB) + bid;
#pragma unroll
                for (int u = 0; u < (int)(Cfg::kB / 16) / SG_BUILDERS; ++u) {
                    const float4 x = big[SG_BUILDERS * u];
                    float4 sm;
                    sm.x = sg_tf32(x.x - sg_trunc(x.x)); sm.y = sg_tf32(x.y - sg_trunc(x.y));
                    sm.z = sg_tf32(x.z - sg_trunc(x.z)); sm.w = sg_tf32(x.w - sg_trunc(x.w));
                    small[SG_BUILDERS * u] = sm;
                }
            }
            if (bid == 0) { rng[2 * c] = lo16; rng[2 * c + 1] = n16; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core's async proxy
            mbar_arrive(built + c);
            ne = ne1; ne1 = ne2; va = va1; vb = vb1; va1 = va2; vb1 = vb2;
            { const int tb_ = b0; b0 = b1; b1 = b2; b2 = tb_; }
            if (++r == R) { r = 0; rph ^= 1u; raw = smem; } else raw += Cfg::kB;
        }
        // epilogue: lane = channel, columns = labels; per label the warp stores its 32 channels as one 128-byte segment
        ok = ok && mbar_wait(done, 0, &g_sg_err);
        tc_fence_after();
        if (ok) {
            float* obase = out + ((int64_t)b * SP * T + t) * C + n0 + tile * 128 + quarter * 32 + lane;
            for (int s0 = lbeg; s0 < lbeg + lw && s0 < SP; s0 += 32) {
                unsigned m[32];
                tc_ld32(lane_addr + (unsigned)s0, m);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int s = s0 + j;
                    if (s < SP) obase[(int64_t)s * T * C] = seg_div(__uint_as_float(m[j]), dtab[s], rtab[s]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kCols) : "memory");
    }
}

// shapes the tensor-core forward takes: whole 32-cell slabs, whole channel blocks, at most two 128-label M-tiles, every cell's
// list complete (cap entries), TMA-addressable maps and 16-byte aligned output rows
static bool seg_mma_eligible(const float* maps, const float* out, int C, int cells, int SP) {
    return cells % SG_K == 0 && cells >= 2 * SG_K && C % 128 == 0 && SP <= 256 && (reinterpret_cast<uintptr_t>(maps) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

template <int NCH>
static int seg_mma_launch(const float* maps, const SegWs& ws, int B, int C, int T, int cells, int SP, float* out, crw_stream_t stream) {
    EncodeTiledFn enc = get_encode();
    CUtensorMap fmap;
    cuuint64_t dims[3] = {(cuuint64_t)cells, (cuuint64_t)T, (cuuint64_t)B * C};
    cuuint64_t strides[2] = {(cuuint64_t)cells * 4, (cuuint64_t)T * cells * 4};
    cuuint32_t box[3] = {(cuuint32_t)SG_K, 1, (cuuint32_t)NCH};
    cuuint32_t estr[3] = {1, 1, 1};
    if (!enc || enc(&fmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(maps), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        set_error("segmean_fwd: cuTensorMapEncodeTiled failed");
        return CRW_ERR_CUDA;
    }
    auto k = segmean_mma_kernel<NCH>;
    const SgLayout L = sg_layout<NCH>(SP);
    static thread_local int attr_device = -1;          // the opt-in is per device and sticky: ask once per thread and device
    int device = 0;
    cudaGetDevice(&device);
    if (attr_device != device) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSgSmemMax) != cudaSuccess) {
            set_error("segmean_fwd: %s", cudaGetErrorString(cudaGetLastError()));
            return CRW_ERR_CUDA;
        }
        attr_device = device;
    }
    static const int dbg = getenv("CRW_SEG_DBG") ? atoi(getenv("CRW_SEG_DBG")) : 0;     // profiling only: 1 no MMAs, 2 no split, 4 no scatter
    k<<<dim3(C / NCH, B * T), SG_THREADS, L.smem, (cudaStream_t)stream>>>(fmap, ws, C, T, cells, SP, L.raw_stages, L.cnt_bytes, dbg, out);
    return check_launch("segmean_mma");
}

}  // namespace crw
#endif  // !CRW_SIM
