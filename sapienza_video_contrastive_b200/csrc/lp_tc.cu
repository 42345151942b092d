// lp_tc.cu - label-propagation affinity + streaming top-k on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
// Reference semantics: code/utils/test_utils.py:148-179 (affinity, mask, /T, topk, softmax) with the radius mask of
// code/utils/__init__.py:377-391 + code/test.py:118-122 evaluated as an integer test.  Same results as lp_simt.cu.
//
// fp32-faithful scores on fp16 tensor cores: every feature is split x = hi + lo * 2^-11 (hi = fp16(x),
// lo = fp16((x - hi) * 2^11)); a score is  <q,k> = qhi.khi + (qhi.klo + qlo.khi) * 2^-11  with fp32 accumulation in
// TMEM (two accumulators, three tcgen05.mma per 16-channel step); the dropped lo.lo term is ~2^-22 relative.
// Dyadic inputs (lo = 0) are reproduced exactly.
//
// One CTA = one target frame x one 16x8 block of 128 queries (the M = 128 rows / TMEM lanes of the MMA):
//   TMEM     the query tile itself (hi and lo planes, 2 x C/2 columns) lives in tensor memory for the CTA's lifetime
//            (tcgen05.mma with the A operand in TMEM): it is re-used by every key tile, so it is never re-read from
//            shared memory, which would otherwise bound small-N MMAs.  The remaining 256 columns hold two accumulator
//            stages (main + correction, 64 keys each).
//   warp 0   TMA producer: the query tile once (staged in shared memory, copied to TMEM by the epilogue warps), then a
//            3-stage ring of 64-key tiles (hi+lo planes, 128B-swizzled).  A tile is two 32-key runs: restricted slots
//            only visit the rows of the block's (16+2R) x (8+2R) window, one run per window row; long-memory slots
//            sweep the frame linearly.
//   warp 1   MMA issuer (one elected thread): per key tile 3 x C/16 tcgen05.mma (M128 N64 K16, kind::f16, A from TMEM,
//            B from shared memory) into the main / correction accumulators of one of two ping-pong stages.
//   warps 2-9  epilogue: thread = (query = TMEM lane, one 32-key run of the tile).  tcgen05.ld pulls main + correction
//            values into registers; validity + the radius test are folded into a branch-free running max, and only
//            when some lane's max beats its k-th best does the warp extract candidates into the register-resident
//            sorted top-k list.  The two half-lists of a query are merged once at the end through shared memory; the
//            affinity matrix never leaves the SM.  Final softmax over the k winners.
#include "common.cuh"

#include "tc_common.cuh"
#include "lp_tc.cuh"

namespace crw {

constexpr int TC_M = 128;        // queries per CTA (16 rows x 8 cols)
constexpr int TC_QH = 16, TC_QW = 8;
constexpr int TC_NS = 32;        // keys per run (one TMA box)
constexpr int TC_NT = 64;        // keys per tile = MMA N (two runs)
constexpr int TC_STAGES = 3;     // key ring depth
constexpr int TC_EPI_WARPS = 8;  // two epilogue warps per TMEM lane quarter, each owning half of a sub-tile's columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;

#ifndef CRW_SIM

// ---- split kernel: fp32 channel-last features -> fp16 hi / lo planes ----------------------------------------------
__global__ void __launch_bounds__(256) lp_split_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const float f[4] = {v.x, v.y, v.z, v.w};
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2half_rn(f[j]);
            l[j] = __float2half_rn((f[j] - __half2float(h[j])) * 2048.0f);
        }
        uint2 ho, lo_;
        ho.x = (unsigned)__half_as_ushort(h[0]) | ((unsigned)__half_as_ushort(h[1]) << 16);
        ho.y = (unsigned)__half_as_ushort(h[2]) | ((unsigned)__half_as_ushort(h[3]) << 16);
        lo_.x = (unsigned)__half_as_ushort(l[0]) | ((unsigned)__half_as_ushort(l[1]) << 16);
        lo_.y = (unsigned)__half_as_ushort(l[2]) | ((unsigned)__half_as_ushort(l[3]) << 16);
        hi[i] = ho;
        lo[i] = lo_;
    }
}

// Enumeration of a CTA's key sub-tiles, shared by producer, issuer and epilogue: slots in order; a long-memory slot
// sweeps the frame in runs of 32 keys; a restricted slot visits one 32-key run per row of the block's window.
struct SubIter {
    int slot, step, nsteps, y0, x0;
    const LpTcArgs* a;
    int qy0, qx0;
    __device__ void init(const LpTcArgs* a_, int qy0_, int qx0_) {
        a = a_; qy0 = qy0_; qx0 = qx0_; slot = 0; step = 0; setup();
    }
    __device__ void setup() {
        while (slot < a->S) {
            const bool restricted = a->restricted && slot >= a->n_long;
            if (restricted) {
                y0 = max(qy0 - a->R, 0);
                const int y1 = min(qy0 + TC_QH - 1 + a->R, a->h - 1);
                nsteps = y1 - y0 + 1;
                x0 = qx0 - a->R;
            } else {
                nsteps = (a->h * a->w + TC_NS - 1) / TC_NS;
            }
            if (nsteps > 0) return;
            ++slot;
        }
    }
    __device__ bool done() const { return slot >= a->S; }
    __device__ int total_runs() const {                               // number of runs this CTA will see
        SubIter c = *this;
        int n = 0;
        while (!c.done()) { n += c.nsteps - c.step; ++c.slot; c.step = 0; c.setup(); }
        return n;
    }
    // key run of the current sub-tile: first key position in the frame (may be negative / wrap: validity is tested per key)
    __device__ int kidx0() const {
        const bool restricted = a->restricted && slot >= a->n_long;
        return restricted ? (y0 + step) * a->w + x0 : step * TC_NS;
    }
    __device__ int row_y() const { return y0 + step; }
    __device__ void next() {
        if (++step >= nsteps) { ++slot; step = 0; setup(); }
    }
};

template <int K>
struct RegTopK {
    float v[K];
    int idx[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < K; ++i) { v[i] = -INFINITY; idx[i] = 0x7fffffff; }
    }
    __device__ __forceinline__ void push(float x, int id) {      // candidates arrive with ascending id
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const bool better = x > v[i];
            const float tv = better ? v[i] : x;
            const int ti = better ? idx[i] : id;
            v[i] = better ? x : v[i];
            idx[i] = better ? id : idx[i];
            x = tv;
            id = ti;
        }
    }
};

template <int K>
__global__ void __launch_bounds__(TC_THREADS, 1)
lp_topk_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                  const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo, LpTcArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int C = a.C, KC = C / 64, hw = a.h * a.w;
    const unsigned q_chunk = (unsigned)TC_M * 128;                // staged query tile: one 64-channel chunk = 128 rows x 128 B
    const unsigned q_plane = (unsigned)KC * q_chunk;
    const unsigned k_chunk = (unsigned)TC_NT * 128;               // key tile: one chunk = 64 rows x 128 B
    const unsigned k_plane = (unsigned)KC * k_chunk;
    const unsigned k_stage = 2 * k_plane;
    unsigned char* k_smem = smem;                                 // [stage][hi|lo][chunk][64 rows][128 B]
    unsigned char* q_smem = smem;                                 // staging of the query tile (aliases the first stages)
    uint64_t* bars = reinterpret_cast<uint64_t*>(k_smem + TC_STAGES * k_stage);
    uint64_t* q_full = bars;
    uint64_t* q_ready = bars + 1;
    uint64_t* k_full = bars + 2;
    uint64_t* k_empty = k_full + TC_STAGES;
    uint64_t* t_full = k_empty + TC_STAGES;
    uint64_t* t_empty = t_full + 2;
    unsigned* tmem_base_smem = reinterpret_cast<unsigned*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y;
    const int tiles_x = (a.w + TC_QW - 1) / TC_QW;
    const int qy0 = (blockIdx.x / tiles_x) * TC_QH, qx0 = (blockIdx.x % tiles_x) * TC_QW;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_ready, 32 * TC_EPI_WARPS);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_base_smem;
    const unsigned tm_qhi = tmem_base, tm_qlo = tmem_base + (unsigned)(C / 2), tm_acc = tmem_base + 256u;

    if (warp == 0) {
        // ===================== TMA producer (whole warp: one lane per box, so a tile's 16 loads issue in parallel) ==========
        {
            const int64_t qrow0 = a.query_frames[n] * (int64_t)hw;
            if (lane == 0) mbar_expect_tx(q_full, 2 * q_plane);
            __syncwarp();
            for (int op = lane; op < TC_QH * KC * 2; op += 32) {
                const int yy = op / (KC * 2), rem = op - yy * (KC * 2), c = rem >> 1, pl = rem & 1;
                const int row = (int)(qrow0 + (int64_t)(qy0 + yy) * a.w + qx0);
                tma_load_2d(q_smem + (pl ? q_plane : 0u) + c * q_chunk + yy * 1024, pl ? &map_q_lo : &map_q_hi, c * 64, row, q_full);
            }
            bool ok = mbar_wait(q_ready, 0, a.err);              // the staged query tile has moved to TMEM
            SubIter it;
            it.init(&a, qy0, qx0);
            for (unsigned t = 0; ok && !it.done(); ++t) {
                const unsigned st = t % TC_STAGES, ph = (t / TC_STAGES) & 1u;
                ok = mbar_wait(k_empty + st, ph ^ 1u, a.err);
                if (!ok) break;
                unsigned char* dst = k_smem + st * k_stage;
                int rows[2];
                int nrun = 0;
                for (; nrun < 2 && !it.done(); ++nrun, it.next())
                    rows[nrun] = (int)(a.key_frames[(int64_t)n * a.S + it.slot] * (int64_t)hw + it.kidx0());
                if (lane == 0) mbar_expect_tx(k_full + st, (unsigned)nrun * 2u * (unsigned)KC * (TC_NS * 128u));
                __syncwarp();
                if (lane < nrun * KC * 2) {
                    const int r = lane / (KC * 2), rem = lane - r * (KC * 2), c = rem >> 1, pl = rem & 1;
                    tma_load_2d(dst + (pl ? k_plane : 0u) + c * k_chunk + r * (TC_NS * 128), pl ? &map_k_lo : &map_k_hi, c * 64,
                                rows[r], k_full + st);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs this loop with warp-uniform values (so the descriptors / TMEM addresses live in uniform
        // registers and every tcgen05.mma is a single UTCHMMA, no divergence waterfall); one elected lane issues.
        {
            const unsigned idesc = umma_idesc_f16(TC_M, TC_NT);
            const unsigned tb = __shfl_sync(kFull, tmem_base, 0);
            const unsigned u_qhi = tb, u_qlo = tb + (unsigned)(C / 2), u_acc = tb + 256u;
            bool ok = mbar_wait(q_ready, 0, a.err);
            tc_fence_after();
            SubIter it;
            it.init(&a, qy0, qx0);
            const unsigned ntiles = (unsigned)((it.total_runs() + 1) / 2);
            const unsigned k_base = __shfl_sync(kFull, smem_u32(k_smem), 0);
            for (unsigned t = 0; ok && t < ntiles; ++t) {
                const unsigned as = t & 1u, aph = (t >> 1) & 1u;
                const unsigned st = t % TC_STAGES, ph = (t / TC_STAGES) & 1u;
                ok = mbar_wait(t_empty + as, aph ^ 1u, a.err);
                if (!ok) break;
                ok = mbar_wait(k_full + st, ph, a.err);
                if (!ok) break;
                tc_fence_after();
                const unsigned ka = k_base + st * k_stage;
                const unsigned d_main = u_acc + as * 128u, d_corr = d_main + 64u;
                if (elect_one()) {
                    unsigned lk_hi = desc_lo(ka), lk_lo = desc_lo(ka + k_plane);
                    unsigned a_hi = u_qhi, a_lo = u_qlo;
                    for (int cch = 0; cch < KC; ++cch) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t k_hi = desc_make(lk_hi + 2u * kk), k_lo = desc_make(lk_lo + 2u * kk);
                            const unsigned accum = (cch | kk) ? 1u : 0u;
                            tc_mma_f16_ts(d_main, a_hi + 8u * kk, k_hi, idesc, accum);
                            tc_mma_f16_ts(d_corr, a_hi + 8u * kk, k_lo, idesc, accum);
                            tc_mma_f16_ts(d_corr, a_lo + 8u * kk, k_hi, idesc, 1u);
                        }
                        lk_hi += k_chunk >> 4; lk_lo += k_chunk >> 4;
                        a_hi += 32u; a_lo += 32u;
                    }
                    tc_commit(k_empty + st);              // frees the key stage once these MMAs have read it
                    tc_commit(t_full + as);               // accumulator stage complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: thread = (query = TMEM lane, run of the tile) =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = ew >> 2;                                      // which 32-key run of a tile / which query plane
        const int q = quarter * 32 + lane;                             // query row inside the tile
        const int qy = qy0 + (q >> 3), qx = qx0 + (q & 7);
        const bool qvalid = qy < a.h && qx < a.w;
        const unsigned lane_sel = (unsigned)(quarter * 32) << 16;
        bool ok = mbar_wait(q_full, 0, a.err);
        // move this query row's plane (half 0: hi, half 1: lo) from the swizzled staging tile to TMEM: 8 fp16 = 4 columns per 16 B
        {
            const unsigned char* src = q_smem + (half ? q_plane : 0u) + (unsigned)q * 128u;
            const unsigned dstc = (half ? tm_qlo : tm_qhi) + lane_sel;
            for (int c = 0; c < KC; ++c) {
                unsigned r[32];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint4 v = *reinterpret_cast<const uint4*>(src + c * q_chunk + ((ch ^ (q & 7)) << 4));
                    r[ch * 4 + 0] = v.x; r[ch * 4 + 1] = v.y; r[ch * 4 + 2] = v.z; r[ch * 4 + 3] = v.w;
                }
                tc_st32(dstc + (unsigned)c * 32u, r);
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(q_ready);
        }
        RegTopK<K> top;
        top.init();
        SubIter it;
        it.init(&a, qy0, qx0);
        if (half && !it.done()) it.next();                             // this warp's run of tile 0
        const unsigned ntiles = (unsigned)((it.total_runs() + 1) / 2);
        for (unsigned t = 0; ok && t < ntiles; ++t) {
            const bool mine_exists = !it.done();                       // the last tile may hold a single run
            const unsigned as = t & 1u, aph = (t >> 1) & 1u;
            ok = mbar_wait(t_full + as, aph, a.err);
            if (!ok) break;
            tc_fence_after();
            if (mine_exists) {
                const bool restricted = a.restricted && it.slot >= a.n_long;
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    unsigned m[16], c[16];
                    const unsigned col = tm_acc + as * 128u + (unsigned)half * 32u + (unsigned)part * 16u + lane_sel;
                    tc_ld16(col, m);
                    tc_ld16(col + 64u, c);
                    tc_wait_ld();
                    const int kidx0 = it.kidx0() + part * 16;
                    const int base_id = it.slot * hw + kidx0;
                    float sv[16];
                    float mx = -INFINITY;
                    if (restricted) {
                        const int dy = it.row_y() - qy;
                        const int rem = a.r2i - dy * dy;               // admissible iff dx^2 <= rem
                        const int kx0 = it.x0 + part * 16, dx0 = kx0 - qx;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int kx = kx0 + j, dx = dx0 + j;
                            const float sc = fmaf(__uint_as_float(c[j]), 4.8828125e-4f, __uint_as_float(m[j]));
                            const bool adm = (unsigned)kx < (unsigned)a.w && dx * dx <= rem;
                            sv[j] = adm ? sc : -INFINITY;
                            mx = fmaxf(mx, sv[j]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float sc = fmaf(__uint_as_float(c[j]), 4.8828125e-4f, __uint_as_float(m[j]));
                            sv[j] = (kidx0 + j < hw) ? sc : -INFINITY;
                            mx = fmaxf(mx, sv[j]);
                        }
                    }
                    // rare path: extract candidates in descending value (equal values: ascending column) while one beats the k-th best
                    while (__any_sync(kFull, mx > top.v[K - 1])) {
                        if (mx > top.v[K - 1]) {
                            int jm = 0;
#pragma unroll
                            for (int j = 15; j >= 0; --j) jm = (sv[j] == mx) ? j : jm;
                            top.push(mx, base_id + jm);
                            float nm = -INFINITY;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                sv[j] = (j == jm) ? -INFINITY : sv[j];
                                nm = fmaxf(nm, sv[j]);
                            }
                            mx = nm;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(t_empty + as);
            // advance to this warp's run of the next tile
            if (!it.done()) it.next();
            if (!it.done()) it.next();
        }
        // merge the two half-lists of each query (sorted by value desc; equal values: lower index first) and write out.
        // All MMAs are complete (the last t_full was observed), so the key ring's shared memory is free.
        float* mv = reinterpret_cast<float*>(smem);                    // [128 queries][2 halves][K]
        int* mi = reinterpret_cast<int*>(smem + (size_t)TC_M * 2 * K * sizeof(float));
        asm volatile("bar.sync 1, %0;" :: "r"(32 * TC_EPI_WARPS) : "memory");
#pragma unroll
        for (int r = 0; r < K; ++r) { mv[(q * 2 + half) * K + r] = top.v[r]; mi[(q * 2 + half) * K + r] = top.idx[r]; }
        asm volatile("bar.sync 1, %0;" :: "r"(32 * TC_EPI_WARPS) : "memory");
        if (ok && qvalid && half == 0) {
            const int64_t obase = (int64_t)n * a.k * hw + qy * a.w + qx;
            const float* v0 = mv + (q * 2 + 0) * K;
            const float* v1 = mv + (q * 2 + 1) * K;
            const int* i0 = mi + (q * 2 + 0) * K;
            const int* i1 = mi + (q * 2 + 1) * K;
            int h0 = 0, h1 = 0;
            float vals[K];
            int ids[K];
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const float a0 = h0 < K ? v0[h0] : -INFINITY, a1 = h1 < K ? v1[h1] : -INFINITY;
                const int b0 = h0 < K ? i0[h0] : 0x7fffffff, b1 = h1 < K ? i1[h1] : 0x7fffffff;
                const bool first = a0 > a1 || (a0 == a1 && b0 <= b1);
                vals[r] = first ? a0 : a1;
                ids[r] = first ? b0 : b1;
                h0 += first ? 1 : 0;
                h1 += first ? 0 : 1;
            }
            float mxv = 0.f, den = 0.f;
#pragma unroll
            for (int r = 0; r < K; ++r) {
                vals[r] = vals[r] / a.tau;
                if (r == 0) mxv = vals[0];
            }
#pragma unroll
            for (int r = 0; r < K; ++r) { vals[r] = r < a.k ? expf(vals[r] - mxv) : 0.f; den += vals[r]; }
#pragma unroll
            for (int r = 0; r < K; ++r) {
                if (r < a.k) {
                    a.Ws[obase + (int64_t)r * hw] = vals[r] / den;
                    a.Is[obase + (int64_t)r * hw] = ids[r] == 0x7fffffff ? 0 : ids[r];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
static bool make_map(CUtensorMap* m, const void* base, int C, int64_t rows, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t lp_tc_workspace_bytes(int Nf, int h, int w, int C) { return 2 * (size_t)Nf * h * w * C * 2 + 256; }

bool lp_tc_supported(int C, int k, float radius, int R, bool dense) {
    return !dense && C % 64 == 0 && C >= 64 && C <= 256 && k >= 1 && k <= 16 && (radius <= 0.f || TC_QW + 2 * R <= TC_NS);
}

int launch_lp_tc(const float* feats, int Nf, const LpTcArgs& a0, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    const int hw = a0.h * a0.w, C = a0.C;
    const size_t plane = (size_t)Nf * hw * C * 2;
    if (workspace_bytes < 2 * plane + 256) { set_error("lp_topk: workspace too small for the tensor-core path"); return CRW_ERR_SHAPE; }
    unsigned char* ws = (unsigned char*)workspace;
    unsigned* err = (unsigned*)ws;
    void* hi = ws + 256;
    void* lo = ws + 256 + plane;
    cudaMemsetAsync(err, 0, 4, (cudaStream_t)stream);
    const int64_t n4 = (int64_t)Nf * hw * C / 4;
    const int sgrid = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    lp_split_kernel<<<sgrid, 256, 0, (cudaStream_t)stream>>>((const float4*)feats, (uint2*)hi, (uint2*)lo, n4);
    int e = check_launch("lp_split");
    if (e != CRW_OK) return e;
    CUtensorMap mqh, mql, mkh, mkl;
    const int64_t rows = (int64_t)Nf * hw;
    if (!make_map(&mqh, hi, C, rows, TC_QW) || !make_map(&mql, lo, C, rows, TC_QW) || !make_map(&mkh, hi, C, rows, TC_NS) ||
        !make_map(&mkl, lo, C, rows, TC_NS)) {
        set_error("lp_topk: cuTensorMapEncodeTiled failed");
        return CRW_ERR_CUDA;
    }
    LpTcArgs a = a0;
    a.err = err;
    size_t ring = TC_STAGES * 2 * (size_t)TC_NT * C * 2, stagingq = 2 * (size_t)TC_M * C * 2;
    const size_t smem = 1024 + (ring > stagingq ? ring : stagingq) + 256;
    dim3 grid(((a.w + TC_QW - 1) / TC_QW) * ((a.h + TC_QH - 1) / TC_QH), a.Nt);
    if (a.k == 10) {
        auto k = lp_topk_tc_kernel<10>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(mqh, mql, mkh, mkl, a);
    } else {
        auto k = lp_topk_tc_kernel<16>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(mqh, mql, mkh, mkl, a);
    }
    return check_launch("lp_topk_tc");
}

#else   // CRW_SIM: the tensor-core path needs real hardware; the host simulator always takes the SIMT kernel

size_t lp_tc_workspace_bytes(int, int, int, int) { return 256; }
bool lp_tc_supported(int, int, float, int, bool) { return false; }
int launch_lp_tc(const float*, int, const LpTcArgs&, void*, size_t, crw_stream_t) { return CRW_ERR_UNSUPPORTED; }

#endif

}  // namespace crw
