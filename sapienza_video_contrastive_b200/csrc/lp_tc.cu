// lp_tc.cu - label-propagation affinity + streaming top-k on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
// Reference semantics: code/utils/test_utils.py:148-179 (affinity, mask, /T, topk, softmax) with the radius mask of
// code/utils/__init__.py:377-391 + code/test.py:118-122 evaluated as an integer test.  Same results as lp_simt.cu.
//
// Three kernels per call (round 2: one third of the tensor-core work of round 1, results certified exact):
//   1. lp_topk_tc_kernel<PRE>    pre-ranking on the fp16 "hi" planes only: ONE tcgen05.mma per 16-channel step.  With
//      x = hi + dx, |dx| <= 2^-11 |x|, a pre-score differs from the exact dot product by at most
//      eps = 1.05e-3 |q| max|k| (+ a denormal term).  The kernel keeps, per query, the 16 best pre-scores (the shortlist).
//   2. lp_rescore_kernel<CHECK>  exact fp32 scores of the 16 shortlisted keys (sequential fmaf over the channels, the very
//      order of the SIMT kernel), ranking by (score desc, index asc), softmax over the k winners -> Ws, Is.  Every key
//      outside the shortlist has pre-score <= the 16th, hence exact score <= pre16 + eps: when the k-th exact score
//      exceeds that bound the top-k is PROVEN identical to a full exact evaluation.  Otherwise the query's tile is put
//      on a device-side work list (exact ties - replicated first frames, vos.py:148-149 - always land here).
//   3. lp_topk_tc_kernel<EXACT> + lp_rescore_kernel<NOCHECK> over the listed tiles only: fp32-faithful scores from the fp16
//      hi/lo split, x = hi + lo * 2^-11, <q,k> = qhi.khi + (qhi.klo + qlo.khi) * 2^-11 with fp32 accumulation in two TMEM
//      accumulators (three tcgen05.mma per step, dropped term ~2^-22 relative, dyadic inputs exact) - round 1's kernel -
//      producing a shortlist that goes through the same exact re-ranking.  (k > 12 is left to the SIMT kernel.)
//
// One CTA = one target frame x one 16x8 block of 128 queries (the M = 128 rows / TMEM lanes of the MMA):
//   TMEM     the query tile itself (hi, and lo in EXACT mode; C/2 columns per plane) lives in tensor memory for the CTA's
//            lifetime (tcgen05.mma with the A operand in TMEM): it is re-used by every key tile, so it is never re-read
//            from shared memory, which would otherwise bound small-N MMAs.  256 further columns hold two accumulator
//            stages (main [+ correction], 64 keys each).
//   warp 0   TMA producer: the query tile once (staged in shared memory, copied to TMEM by the epilogue warps), then a
//            ring of 64-key tiles (128B-swizzled; 6 stages of the hi plane in PRE mode, 3 stages of hi+lo in EXACT mode).
//            A tile is two 32-key runs: restricted slots only visit the rows of the block's (16+2R) x (8+2R) window, one
//            run per window row; long-memory slots sweep the frame linearly.
//   warp 1   MMA issuer (one elected thread): per key tile 1 or 3 x C/16 tcgen05.mma (M128 N64 K16, kind::f16, A from
//            TMEM, B from shared memory) into one of two ping-pong accumulator stages.
//   warps 2-9  epilogue: thread = (query = TMEM lane, one 32-key run of the tile).  tcgen05.ld pulls the scores into
//            registers; validity + the radius test are one bit mask per run, folded into a branch-free running max, and
//            only when some lane's max beats its 16th best does the warp extract candidates into the register-resident
//            sorted list.  The two half-lists of a query are merged once at the end through shared memory; the affinity
//            matrix never leaves the SM.
#include "common.cuh"

#include "tc_common.cuh"
#include "lp_tc.cuh"

namespace crw {

constexpr int TC_M = 128;        // queries per CTA (16 rows x 8 cols)
constexpr int TC_QH = 16, TC_QW = 8;
constexpr int TC_NS = 32;        // keys per run (one TMA box)
constexpr int TC_NT = 64;        // keys per tile = MMA N (two runs)
constexpr int TC_EPI_WARPS = 8;  // two epilogue warps per TMEM lane quarter, each owning half of a sub-tile's columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int LP_SHORT = 16;     // shortlist length per query
constexpr int LP_HALF = 12;      // candidates each of a query's two epilogue threads keeps (>= LP_PRE_MAX_K: the k best may all sit in one half)
constexpr int LP_PRE_MAX_K = 12; // the pre-ranking pass needs spare shortlist entries below the k-th to certify a query
constexpr int MODE_PRE = 0, MODE_EXACT = 1;
constexpr unsigned LP_QX_CAP = 1024;   // up to this many uncertified queries are settled one by one (lp_exact_query_kernel)
// workspace header words
constexpr int HDR_ERR = 0, HDR_COUNT = 1, HDR_KNORM2 = 2, HDR_MAXABS = 3, HDR_FLAGGED = 4, HDR_BYTES = 1024;
// |pre-score - exact fp32 score| <= LP_EPS_REL |q| max|k| + LP_EPS_ABS (|q| + max|k|):  2 * 2^-11 (1 + 2^-12) from rounding
// both operands to fp16, 2e-5 for the tensor core's truncating fp32 accumulation over <= 256 channels, 1.6e-5 for the fp32
// rounding of the exact score itself; the absolute term covers fp16 subnormals (|x| < 6.1e-5: error <= 2^-25 per element)
constexpr float LP_EPS_REL = 1.05e-3f, LP_EPS_ABS = 1e-6f;

#ifndef CRW_SIM

// ---- split kernel: fp32 channel-last features -> fp16 hi / lo planes; also max squared row norm and max |x| ---------
__global__ void __launch_bounds__(256) lp_split_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                                       int64_t rows, int c4, unsigned* hdr) {
    const int lane = threadIdx.x & 31;
    float best_ss = 0.f, best_abs = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        float ss = 0.f;
        for (int i = lane; i < c4; i += 32) {
            const float4 v = __ldg(x + r * c4 + i);
            const float f[4] = {v.x, v.y, v.z, v.w};
            __half h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = __float2half_rn(f[j]);
                l[j] = __float2half_rn((f[j] - __half2float(h[j])) * 2048.0f);
                ss = fmaf(f[j], f[j], ss);
                best_abs = fmaxf(best_abs, fabsf(f[j]));
            }
            uint2 ho, lo_;
            ho.x = (unsigned)__half_as_ushort(h[0]) | ((unsigned)__half_as_ushort(h[1]) << 16);
            ho.y = (unsigned)__half_as_ushort(h[2]) | ((unsigned)__half_as_ushort(h[3]) << 16);
            lo_.x = (unsigned)__half_as_ushort(l[0]) | ((unsigned)__half_as_ushort(l[1]) << 16);
            lo_.y = (unsigned)__half_as_ushort(l[2]) | ((unsigned)__half_as_ushort(l[3]) << 16);
            hi[r * c4 + i] = ho;
            lo[r * c4 + i] = lo_;
        }
        best_ss = fmaxf(best_ss, warp_sum(ss));
    }
    best_abs = warp_max(best_abs);
    __shared__ float sh[16];
    if (lane == 0) { sh[threadIdx.x >> 5] = best_ss; sh[8 + (threadIdx.x >> 5)] = best_abs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a = fmaxf(a, sh[i]); b = fmaxf(b, sh[8 + i]); }
        atomicMax(hdr + HDR_KNORM2, __float_as_uint(a));             // non-negative floats order like their bit patterns
        atomicMax(hdr + HDR_MAXABS, __float_as_uint(b));
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

// bits j = 0..15 set where key column j of a 16-key part is admissible
__device__ __forceinline__ unsigned range_mask16(int jlo, int jhi) {
    jlo = max(jlo, 0);
    jhi = min(jhi, 15);
    return jhi < jlo ? 0u : ((0xffffu >> (15 - jhi)) & (0xffffu << jlo)) & 0xffffu;
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
lp_topk_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                  const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo, LpTcArgs a) {
    constexpr int K = LP_SHORT;
    constexpr int PLANES = MODE == MODE_EXACT ? 2 : 1;
    constexpr int STAGES = MODE == MODE_EXACT ? 3 : 6;
    // accumulator stages in tensor memory: the pre-ranking pass needs 64 columns per stage and its query tile 128, so SIX stages
    // fit - the MMA warp then runs up to six tiles ahead of the slowest epilogue warp, which turns "every tile waits for the
    // unluckiest warp's insertions" into "throughput = average epilogue rate"; the fp32-faithful pass (128 + 256 columns) has two
    constexpr unsigned ACC_STAGES = MODE == MODE_EXACT ? 2u : 6u, ACC_STRIDE = MODE == MODE_EXACT ? 128u : 64u;
    constexpr unsigned ACC_BASE = MODE == MODE_EXACT ? 256u : 128u;
    extern __shared__ unsigned char smem_dyn[];
    const int tiles_x = (a.w + TC_QW - 1) / TC_QW;
    const int tiles = tiles_x * ((a.h + TC_QH - 1) / TC_QH);
    // work items = (target frame, query tile).  Without a list: one per CTA (2-D grid).  With the work list written by the
    // certification pass: a small fixed grid strides over it (launching one early-exit CTA per possible tile cost 0.57 ms of
    // block scheduling per call for a list of 21 entries).
    unsigned work_first, work_end, work_step;
    if (a.list) {
        if (a.count[HDR_FLAGGED - HDR_COUNT] <= LP_QX_CAP) return;      // few open queries: lp_exact_query_kernel settles them
        work_first = blockIdx.x; work_end = *a.count; work_step = gridDim.x;
        if (work_first >= work_end) return;
    } else {
        work_first = blockIdx.y * (unsigned)tiles + blockIdx.x; work_end = work_first + 1u; work_step = 1u;
    }
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);     // stays in the shared state space: LDS / STS, not generic LD / ST
    const int C = a.C, KC = C / 64, hw = a.h * a.w;
    const unsigned q_chunk = (unsigned)TC_M * 128;                // staged query tile: one 64-channel chunk = 128 rows x 128 B
    const unsigned q_plane = (unsigned)KC * q_chunk;
    const unsigned k_chunk = (unsigned)TC_NT * 128;               // key tile: one chunk = 64 rows x 128 B
    const unsigned k_plane = (unsigned)KC * k_chunk;
    const unsigned k_stage = PLANES * k_plane;
    unsigned char* k_smem = smem;                                 // [stage][hi|lo][chunk][64 rows][128 B]
    unsigned char* q_smem = smem;                                 // staging of the query tile (aliases the first stages)
    float* park_s = reinterpret_cast<float*>(k_smem + STAGES * k_stage);      // [8][256 epilogue threads][4] scores of the run in hand
    float* thr_s = park_s + 32 * 32 * TC_EPI_WARPS;                           // [128 queries][2 halves] published 12th-best scores
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr_s + TC_M * 2);
    uint64_t* q_full = bars;
    uint64_t* q_ready = bars + 1;
    uint64_t* k_full = bars + 2;
    uint64_t* k_empty = k_full + STAGES;
    uint64_t* t_full = k_empty + STAGES;
    uint64_t* t_empty = t_full + ACC_STAGES;
    unsigned* tmem_base_smem = reinterpret_cast<unsigned*>(t_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_base_smem;
    const unsigned tm_qhi = tmem_base, tm_qlo = tmem_base + (unsigned)(C / 2), tm_acc = tmem_base + ACC_BASE;

    for (unsigned work = work_first; work < work_end; work += work_step) {
    const int wi = a.list ? a.list[work] : (int)work;
    const int n = wi / tiles, tile = wi - n * tiles;
    const int qy0 = (tile / tiles_x) * TC_QH, qx0 = (tile % tiles_x) * TC_QW;
    if (threadIdx.x == 0) {                                       // fresh barriers for every work item (all phases restart at 0)
        const bool again = work != work_first;
        const int nb = 2 + 2 * STAGES + 2 * (int)ACC_STAGES;
        if (again) for (int i = 0; i < nb; ++i) mbar_inval(bars + i);
        mbar_init(q_full, 1);
        mbar_init(q_ready, 32 * TC_EPI_WARPS);
        for (int s = 0; s < STAGES; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
        for (unsigned s = 0; s < ACC_STAGES; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ===================== TMA producer (whole warp: one lane per box, so a tile's loads issue in parallel) ==========
        {
            const int64_t qrow0 = a.query_frames[n] * (int64_t)hw;
            if (lane == 0) mbar_expect_tx(q_full, PLANES * q_plane);
            __syncwarp();
            for (int op = lane; op < TC_QH * KC * PLANES; op += 32) {
                const int yy = op / (KC * PLANES), rem = op - yy * (KC * PLANES), c = rem / PLANES, pl = rem - c * PLANES;
                const int row = (int)(qrow0 + (int64_t)(qy0 + yy) * a.w + qx0);
                tma_load_2d(q_smem + (pl ? q_plane : 0u) + c * q_chunk + yy * 1024, pl ? &map_q_lo : &map_q_hi, c * 64, row, q_full);
            }
            bool ok = mbar_wait(q_ready, 0, a.err);              // the staged query tile has moved to TMEM
            SubIter it;
            it.init(&a, qy0, qx0);
            for (unsigned t = 0; ok && !it.done(); ++t) {
                const unsigned st = t % STAGES, ph = (t / STAGES) & 1u;
                ok = mbar_wait(k_empty + st, ph ^ 1u, a.err);
                if (!ok) break;
                unsigned char* dst = k_smem + st * k_stage;
                int rows[2];
                int nrun = 0;
                for (; nrun < 2 && !it.done(); ++nrun, it.next())
                    rows[nrun] = (int)(a.key_frames[(int64_t)n * a.S + it.slot] * (int64_t)hw + it.kidx0());
                if (lane == 0) mbar_expect_tx(k_full + st, (unsigned)nrun * (unsigned)PLANES * (unsigned)KC * (TC_NS * 128u));
                __syncwarp();
                if (lane < nrun * KC * PLANES) {
                    const int r = lane / (KC * PLANES), rem = lane - r * (KC * PLANES), c = rem / PLANES, pl = rem - c * PLANES;
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
            const unsigned u_qhi = tb, u_qlo = tb + (unsigned)(C / 2), u_acc = tb + ACC_BASE;
            bool ok = mbar_wait(q_ready, 0, a.err);
            tc_fence_after();
            SubIter it;
            it.init(&a, qy0, qx0);
            const unsigned ntiles = (unsigned)((it.total_runs() + 1) / 2);
            const unsigned k_base = __shfl_sync(kFull, smem_u32(k_smem), 0);
            for (unsigned t = 0; ok && t < ntiles; ++t) {
                const unsigned as = t % ACC_STAGES, aph = (t / ACC_STAGES) & 1u;
                const unsigned st = t % STAGES, ph = (t / STAGES) & 1u;
                ok = mbar_wait(t_empty + as, aph ^ 1u, a.err);
                if (!ok) break;
                ok = mbar_wait(k_full + st, ph, a.err);
                if (!ok) break;
                tc_fence_after();
                const unsigned ka = k_base + st * k_stage;
                const unsigned d_main = u_acc + as * ACC_STRIDE, d_corr = d_main + 64u;
                if (elect_one()) {
                    unsigned lk_hi = desc_lo(ka), lk_lo = desc_lo(ka + k_plane);
                    unsigned a_hi = u_qhi, a_lo = u_qlo;
                    for (int cch = 0; cch < KC; ++cch) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t k_hi = desc_make(lk_hi + 2u * kk);
                            const unsigned accum = (cch | kk) ? 1u : 0u;
                            tc_mma_f16_ts(d_main, a_hi + 8u * kk, k_hi, idesc, accum);
                            if (MODE == MODE_EXACT) {
                                const uint64_t k_lo = desc_make(lk_lo + 2u * kk);
                                tc_mma_f16_ts(d_corr, a_hi + 8u * kk, k_lo, idesc, accum);
                                tc_mma_f16_ts(d_corr, a_lo + 8u * kk, k_hi, idesc, 1u);
                            }
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
            if (MODE == MODE_EXACT || half == 0) {
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
            }
            tc_fence_before();
            mbar_arrive(q_ready);
        }
        // Per-thread candidate list: the LP_HALF best (score, id) pairs of this thread's half of the key stream, sorted, in
        // registers.  Per 32-key run: (1) one compare per key builds a per-lane candidate bit mask (score > thr, admissible);
        // (2) if any lane has a candidate, the run's scores are parked in shared memory (so that a lane can fetch "its j-th
        // score" with a run-time j) and the warp loops: every lane with a candidate left pops its lowest bit and runs the
        // 12-stage compare-and-shift insertion - candidates of DIFFERENT lanes are inserted in the same pass.
        // thr = this thread's admission threshold = max(its own 12th best, the other half's 12th best, published through
        // shared memory once per run): a key below either bound cannot be among the query's 16 best.
        // What was tried and measured on the way (ncu: the epilogue warps, not the tensor pipe, set the pace - 1.8 M warp
        // instructions per CTA in round 1): extraction of the running maximum per 16-key part, 150 instructions per pass,
        // 473 us per CTA; lists in shared memory updated cooperatively by a half-warp per candidate, one serial shuffle /
        // vote / load chain per candidate, 1.5 ms; one fully unrolled insertion chain per key column, 36 KB of loop body
        // (instruction-cache misses), 1.2 ms.
        float tv[LP_HALF];
        int ti[LP_HALF];
#pragma unroll
        for (int r = 0; r < LP_HALF; ++r) { tv[r] = -INFINITY; ti[r] = 0x7fffffff; }
        volatile float* thr_pub = thr_s + (q * 2 + half);
        const volatile float* thr_other = thr_s + (q * 2 + (half ^ 1));
        *thr_pub = -INFINITY;
        asm volatile("bar.sync 1, %0;" :: "r"(32 * TC_EPI_WARPS) : "memory");      // (the other half's threshold is read below)
        float thr = -INFINITY;
        const int etid = threadIdx.x - 64;                             // 0..255 among the epilogue threads
        float4* park4 = reinterpret_cast<float4*>(park_s) + etid;      // [8 groups of 4 columns][256 threads] float4: conflict-free stores
        const float* park1 = park_s + etid * 4;
        // Run enumeration (must match SubIter, which the producer and the issuer use): slots in order, long-memory slots sweep
        // the frame in n_sweep runs, restricted slots visit the block's window rows wy0..wy1.  This warp owns runs
        // half, half + 2, ... of the concatenated list, i.e. one run per tile.
        const int n_sweep = (hw + TC_NS - 1) / TC_NS;
        const int wy0 = max(qy0 - a.R, 0), wy1 = min(qy0 + TC_QH - 1 + a.R, a.h - 1);
        const int n_rows = wy1 - wy0 + 1, wx0 = qx0 - a.R;
        const int n_unres = a.restricted ? a.n_long : a.S;
        const unsigned total_runs = (unsigned)(n_unres * n_sweep + (a.S - n_unres) * n_rows);
        const unsigned ntiles = (total_runs + 1u) / 2u;
        // rows of the window this warp's four query rows can reach at all (warp-uniform skip of the others)
        const int wq_lo = qy0 + quarter * 4 - a.R, wq_hi = qy0 + quarter * 4 + 3 + a.R;
        // half-width of the disc per row offset, 4 bits each (radius <= 12): dx^2 <= r2i - dy^2
        unsigned long long dxm_lut = 0ull;
        for (int d = 0; d <= a.R; ++d) dxm_lut |= (unsigned long long)((int)sqrtf((float)(a.r2i - d * d))) << (4 * d);
        int slot = 0, step = half;
        int nsteps = slot < n_unres ? n_sweep : n_rows;
        while (slot < a.S && step >= nsteps) { step -= nsteps; ++slot; nsteps = slot < n_unres ? n_sweep : n_rows; }
        for (unsigned t = 0; ok && t < ntiles; ++t) {
            const bool mine_exists = slot < a.S;                       // the last tile may hold a single run
            const unsigned as = t % ACC_STAGES, aph = (t / ACC_STAGES) & 1u;
            bool live = mine_exists;
            unsigned adm = 0u;
            int base_id = 0;
            if (mine_exists) {
                // admissible key columns of this run, as one 32-bit mask (bit j = key kidx0 + j)
                if (slot >= n_unres) {
                    const int wy = wy0 + step;
                    live = wy >= wq_lo && wy <= wq_hi;                 // warp-uniform
                    const int dy = abs(wy - qy);
                    if (dy <= a.R && qvalid) {
                        const int dxm = (int)((dxm_lut >> (4 * dy)) & 15ull);
                        const int jlo = max(qx - dxm, 0) - wx0, jhi = min(qx + dxm, a.w - 1) - wx0;      // 0 <= jlo <= jhi <= 29
                        adm = (0xffffffffu >> (31 - jhi)) & (0xffffffffu << jlo);
                    }
                    base_id = slot * hw + wy * a.w + wx0;
                } else {
                    const int nvalid = hw - step * TC_NS;
                    adm = !qvalid ? 0u : (nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u));
                    base_id = slot * hw + step * TC_NS;
                }
            }
            live = live && __any_sync(kFull, adm != 0u);
            ok = mbar_wait(t_full + as, aph, a.err);
            if (!ok) break;
            tc_fence_after();
            float sv[32];
            if (live) {
                // pull the whole 32-key run into registers, then release the accumulator stage BEFORE the selection work:
                // the MMA into this stage only waits for the loads, not for the insertions
                unsigned m0[16], m1[16];
                const unsigned col = tm_acc + as * ACC_STRIDE + (unsigned)half * 32u + lane_sel;
                tc_ld16(col, m0);
                tc_ld16(col + 16u, m1);
                if (MODE == MODE_EXACT) {
                    unsigned c0[16], c1[16];
                    tc_ld16(col + 64u, c0);
                    tc_ld16(col + 80u, c1);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        sv[j] = fmaf(__uint_as_float(c0[j]), 4.8828125e-4f, __uint_as_float(m0[j]));
                        sv[16 + j] = fmaf(__uint_as_float(c1[j]), 4.8828125e-4f, __uint_as_float(m1[j]));
                    }
                } else {
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { sv[j] = __uint_as_float(m0[j]); sv[16 + j] = __uint_as_float(m1[j]); }
                }
            }
            tc_fence_before();
            mbar_arrive(t_empty + as);
            if (live) {
                // (>= the other half's bound, > the own one: a key that TIES with the other half's 12th best may still carry the
                // lower index, so it must get in - nextafter turns ">=" into the single ">" the mask loop evaluates)
                thr = fmaxf(thr, nextafterf(*thr_other, -INFINITY));
                // candidate mask: bit j = sign(thr - score_j), shifted in from the top column down (two instructions per key)
                // (four independent 8-bit chains: one 32-long chain of dependent shifts stalled the two warps of a scheduler)
                unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                    c0 = __funnelshift_l(__float_as_uint(thr - sv[j]), c0, 1);
                    c1 = __funnelshift_l(__float_as_uint(thr - sv[8 + j]), c1, 1);
                    c2 = __funnelshift_l(__float_as_uint(thr - sv[16 + j]), c2, 1);
                    c3 = __funnelshift_l(__float_as_uint(thr - sv[24 + j]), c3, 1);
                }
                unsigned cm = (c0 | (c1 << 8) | (c2 << 16) | (c3 << 24)) & adm;
                if (__any_sync(kFull, cm != 0u)) {
                    __syncwarp();
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) park4[g4 * 256] = make_float4(sv[4 * g4], sv[4 * g4 + 1], sv[4 * g4 + 2], sv[4 * g4 + 3]);
                    while (__any_sync(kFull, cm != 0u)) {
                        if (cm != 0u) {
                            const int j = __ffs(cm) - 1;
                            cm &= cm - 1u;
                            const float x = park1[(j >> 2) * 1024 + (j & 3)];
                            const int id = base_id + j;
                            // sorted insert with every comparison made against the SAME x (no serial compare-and-shift chain):
                            // entries >= x stay (equal scores keep the earlier, lower id ahead), x lands behind them, the rest shift
                            bool ge[LP_HALF];
#pragma unroll
                            for (int r = 0; r < LP_HALF; ++r) ge[r] = tv[r] >= x;
#pragma unroll
                            for (int r = LP_HALF - 1; r >= 1; --r) {
                                tv[r] = ge[r] ? tv[r] : (ge[r - 1] ? x : tv[r - 1]);
                                ti[r] = ge[r] ? ti[r] : (ge[r - 1] ? id : ti[r - 1]);
                            }
                            tv[0] = ge[0] ? tv[0] : x;
                            ti[0] = ge[0] ? ti[0] : id;
                        }
                    }
                    thr = fmaxf(thr, tv[LP_HALF - 1]);
                    *thr_pub = tv[LP_HALF - 1];
                }
            }
            // advance to this warp's run of the next tile
            step += 2;
            while (slot < a.S && step >= nsteps) { step -= nsteps; ++slot; nsteps = slot < n_unres ? n_sweep : n_rows; }
        }
        // All MMAs are complete (the last t_full was observed), so the key ring's shared memory is free: park the lists there
        float* lists_v = reinterpret_cast<float*>(smem);                // [128 queries][2 halves][LP_HALF]
        int* lists_i = reinterpret_cast<int*>(smem + (size_t)TC_M * 2 * LP_HALF * sizeof(float));
        asm volatile("bar.sync 1, %0;" :: "r"(32 * TC_EPI_WARPS) : "memory");
#pragma unroll
        for (int r = 0; r < LP_HALF; ++r) { lists_v[(q * 2 + half) * LP_HALF + r] = tv[r]; lists_i[(q * 2 + half) * LP_HALF + r] = ti[r]; }
        // merge the two half-lists of each query (sorted by value desc; equal values: lower index first) into the 16-key
        // shortlist.  Its last score is stored as the BOUND on every key that is not in the shortlist: such a key was either
        // cut by the merge (score <= the 16th) or never admitted to a half-list (score <= that half's 12th best).
        asm volatile("bar.sync 1, %0;" :: "r"(32 * TC_EPI_WARPS) : "memory");
        if (ok && qvalid && half == 0) {
            const float* v0 = lists_v + (size_t)(q * 2 + 0) * LP_HALF;
            const float* v1 = lists_v + (size_t)(q * 2 + 1) * LP_HALF;
            const int* i0 = lists_i + (size_t)(q * 2 + 0) * LP_HALF;
            const int* i1 = lists_i + (size_t)(q * 2 + 1) * LP_HALF;
            int h0 = 0, h1 = 0;
            float vals[K];
            int ids[K];
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const float a0 = h0 < LP_HALF ? v0[h0] : -INFINITY, a1 = h1 < LP_HALF ? v1[h1] : -INFINITY;
                const int b0 = h0 < LP_HALF ? i0[h0] : 0x7fffffff, b1 = h1 < LP_HALF ? i1[h1] : 0x7fffffff;
                const bool first = a0 > a1 || (a0 == a1 && b0 <= b1);
                vals[r] = first ? a0 : a1;
                ids[r] = first ? b0 : b1;
                h0 += first ? 1 : 0;
                h1 += first ? 0 : 1;
            }
            vals[K - 1] = fmaxf(vals[K - 1], fmaxf(v0[LP_HALF - 1], v1[LP_HALF - 1]));     // (-inf while a half-list is not full)
            const int64_t obase = ((int64_t)n * hw + qy * a.w + qx) * K;
            float4* ov = reinterpret_cast<float4*>(a.short_v + obase);
            int4* oi = reinterpret_cast<int4*>(a.short_i + obase);
#pragma unroll
            for (int r = 0; r < K / 4; ++r) {
                ov[r] = make_float4(vals[4 * r], vals[4 * r + 1], vals[4 * r + 2], vals[4 * r + 3]);
                oi[r] = make_int4(ids[4 * r], ids[4 * r + 1], ids[4 * r + 2], ids[4 * r + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();                                              // every role is done with this work item: smem, TMEM and barriers are free
    tc_fence_after();
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- exact re-ranking of the shortlists ---------------------------------------------------------------------------------
// One warp = two queries x 16 shortlisted keys (lane = (query, candidate)).  The key rows are staged through shared memory
// 64 channels at a time with coalesced 256-byte row reads; each lane then runs the SAME sequential fmaf chain over the channels
// as lp_simt.cu, so the scores - and therefore the order, the ties and the softmax weights - are bit-identical to the SIMT
// kernel's.  CHECK: certify the query (see the file header) or put its tile on the work list.  !CHECK: one CTA per listed tile.
struct LpRescoreArgs {
    const float* feats;          // (Nf, hw, C) fp32
    const int64_t* key_frames;   // (Nt, S)
    const int64_t* query_frames; // (Nt)
    const float* short_v;        // (Nt*hw, 16) pre-scores, descending
    const int* short_i;          // (Nt*hw, 16) flat ids slot*hw + pos
    int Nt, S, h, w, C, k;
    float tau;
    float* Ws;
    int64_t* Is;
    unsigned* hdr;               // workspace header
    unsigned* flags;             // (Nt*tiles) tile already listed
    int* list;                   // (Nt*tiles) work list
    int2* qlist;                 // (LP_QX_CAP) uncertified queries (target, position)
    float* qpart_v;              // (LP_QX_CAP, QX_SPLIT, QX_K) partial winners of the per-query kernel
    int* qpart_i;
    unsigned* qticket;           // (LP_QX_CAP) CTAs of a query that have delivered
    int n_unres, restricted, R, r2i;
};

constexpr int RS_WARPS = 4, RS_CH = 64, RS_LD = RS_CH + 4;
// Few uncertified queries (distinct frames: ~1 in 10 000) are settled one by one by lp_exact_query_kernel, an exact-fp32
// evaluation of ALL their admissible keys; only when more than LP_QX_CAP queries are open (replicated frames) is the
// fp32-faithful tensor-core pass over whole tiles the cheaper way.  Both are always launched; the one out of its regime exits.
constexpr int QX_WARPS = 8, QX_K = 12, QX_SPLIT = 8;      // CTAs per open query: each takes every 8th group of 256 keys

template <int CHECK>
__global__ void __launch_bounds__(32 * RS_WARPS) lp_rescore_kernel(LpRescoreArgs a) {
    __shared__ __align__(16) float rows_s[RS_WARPS][32][RS_LD];
    __shared__ __align__(16) float q_s[RS_WARPS][2][RS_CH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = a.h * a.w;
    const int tiles_x = (a.w + TC_QW - 1) / TC_QW;
    const int tiles = tiles_x * ((a.h + TC_QH - 1) / TC_QH);
    const int qq = lane >> 4, cand = lane & 15;
    const float kmax = sqrtf(__uint_as_float(a.hdr[HDR_KNORM2]));
    if (blockIdx.x == 0 && threadIdx.x == 0 && __uint_as_float(a.hdr[HDR_MAXABS]) >= 60000.f) atomicExch(a.hdr + HDR_ERR, 2u);

    int64_t pair0, pair1, stride;                                  // this warp's range of query pairs
    int tile_n = 0, tile_qy0 = 0, tile_qx0 = 0;
    if (CHECK) {
        const int64_t npairs = ((int64_t)a.Nt * hw + 1) / 2;
        pair0 = (int64_t)blockIdx.x * RS_WARPS + warp;
        pair1 = npairs;
        stride = (int64_t)gridDim.x * RS_WARPS;
    } else {
        pair0 = warp;
        pair1 = TC_M / 2;
        stride = RS_WARPS;
    }
    const unsigned n_items = CHECK ? 1u : (a.hdr[HDR_FLAGGED] <= LP_QX_CAP ? 0u : a.hdr[HDR_COUNT]);
    for (unsigned item = CHECK ? 0u : blockIdx.x; item < n_items; item += CHECK ? 1u : gridDim.x) {
    if (!CHECK) {
        const int e = a.list[item];
        tile_n = e / tiles;
        const int tile = e - tile_n * tiles;
        tile_qy0 = (tile / tiles_x) * TC_QH;
        tile_qx0 = (tile % tiles_x) * TC_QW;
    }
    for (int64_t pr = pair0; pr < pair1; pr += stride) {
        // the query of this half-warp
        int n, qpos;
        bool qok;
        if (CHECK) {
            const int64_t g = pr * 2 + qq;
            qok = g < (int64_t)a.Nt * hw;
            n = qok ? (int)(g / hw) : 0;
            qpos = qok ? (int)(g - (int64_t)n * hw) : 0;
        } else {
            const int ql = (int)pr * 2 + qq;                           // 0..127 inside the 16x8 block
            const int qy = tile_qy0 + (ql >> 3), qx = tile_qx0 + (ql & 7);
            qok = qy < a.h && qx < a.w;
            n = tile_n;
            qpos = qok ? qy * a.w + qx : 0;
        }
        const int64_t sbase = ((int64_t)n * hw + qpos) * LP_SHORT + cand;
        const int id = qok ? a.short_i[sbase] : 0x7fffffff;
        const float pre = qok ? a.short_v[sbase] : -INFINITY;
        const bool valid = id != 0x7fffffff;
        const int slot = valid ? id / hw : 0;
        const int pos = valid ? id - slot * hw : 0;
        const float* krow = a.feats + (a.key_frames[(int64_t)n * a.S + slot] * (int64_t)hw + pos) * a.C;
        const float* qrow = a.feats + (a.query_frames[n] * (int64_t)hw + qpos) * a.C;
        float acc = 0.f, qn2 = 0.f;
        for (int c0 = 0; c0 < a.C; c0 += RS_CH) {
            __syncwarp();
            // stage 32 key rows x 64 channels: a half-warp reads one row's 256 bytes
#pragma unroll 8
            for (int i = 0; i < 16; ++i) {
                const int r = 2 * i + (lane >> 4);
                const float* src = reinterpret_cast<const float*>(__shfl_sync(kFull, (unsigned long long)krow, r));
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + c0) + (lane & 15));
                *reinterpret_cast<float4*>(&rows_s[warp][r][(lane & 15) * 4]) = v;
            }
            {
                const float* src = reinterpret_cast<const float*>(__shfl_sync(kFull, (unsigned long long)qrow, (lane >> 4) * 16));
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + c0) + (lane & 15));
                *reinterpret_cast<float4*>(&q_s[warp][lane >> 4][(lane & 15) * 4]) = v;
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < RS_CH; c += 4) {
                const float4 kv = *reinterpret_cast<const float4*>(&rows_s[warp][lane][c]);
                const float4 qv = *reinterpret_cast<const float4*>(&q_s[warp][qq][c]);
                acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
                qn2 = fmaf(qv.x, qv.x, qn2); qn2 = fmaf(qv.y, qv.y, qn2); qn2 = fmaf(qv.z, qv.z, qn2); qn2 = fmaf(qv.w, qv.w, qn2);
            }
        }
        const float raw = valid ? acc : -INFINITY;
        const float v = valid ? acc / a.tau : -INFINITY;              // the SIMT kernel ranks sc / tau
        // rank inside the half-warp by (value desc, index asc, lane asc)
        int rank = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float ov = __shfl_sync(kFull, v, (lane & 16) + i);
            const int oi = __shfl_sync(kFull, id, (lane & 16) + i);
            rank += (ov > v || (ov == v && (oi < id || (oi == id && i < cand)))) ? 1 : 0;
        }
        // softmax over the k winners, summed in rank order like the SIMT kernel
        const unsigned half_mask = 0xffffu << (lane & 16);
        const unsigned b0 = __ballot_sync(kFull, rank == 0) & half_mask;
        const float mxv = __shfl_sync(kFull, v, __ffs(b0) - 1);
        const float e = rank < a.k ? expf(v - mxv) : 0.f;
        float den = 0.f;
        for (int r = 0; r < a.k; ++r) {
            const unsigned b = __ballot_sync(kFull, rank == r) & half_mask;
            den += __shfl_sync(kFull, e, __ffs(b) - 1);
        }
        if (qok && rank < a.k) {
            const int64_t o = ((int64_t)n * a.k + rank) * hw + qpos;
            a.Ws[o] = e / den;
            a.Is[o] = valid ? id : 0;
        }
        if (CHECK) {
            // certification: every key outside the shortlist has pre-score <= pre16, hence exact score <= pre16 + eps
            const unsigned bk = __ballot_sync(kFull, rank == a.k - 1) & half_mask;
            const float tk = __shfl_sync(kFull, raw, __ffs(bk) - 1);
            const float pre16 = __shfl_sync(kFull, pre, (lane & 16) + 15);
            const float qn = sqrtf(qn2);
            const float eps = LP_EPS_REL * qn * kmax + LP_EPS_ABS * (qn + kmax);
            const bool certified = pre16 == -INFINITY || tk > pre16 + eps;
            if (qok && cand == 0 && !certified) {
                const int qy = qpos / a.w, qx = qpos - qy * a.w;
                const int e_ = n * tiles + (qy / TC_QH) * tiles_x + qx / TC_QW;
                const unsigned qi = atomicAdd(a.hdr + HDR_FLAGGED, 1u);
                if (qi < LP_QX_CAP) a.qlist[qi] = make_int2(n, qpos);
                if (atomicExch(a.flags + e_, 1u) == 0u) a.list[atomicAdd(a.hdr + HDR_COUNT, 1u)] = e_;
            }
        }
    }
    }
}

// ---- exact evaluation of single queries ---------------------------------------------------------------------------------
// One CTA = one uncertified query.  Every admissible key (all positions of the long-memory slots, the disc of the restricted
// ones) is scored in exact fp32 with the same staging and the same sequential fmaf chain as lp_rescore_kernel / lp_simt.cu
// (lane = key), each lane keeps its k best in registers, the 256 sorted lane lists are merged by k rounds of a block-wide
// arg-max in (score desc, index asc) order, then the softmax of the SIMT kernel.  ~14 k keys x 1 KB per query: ~0.1 ms.
__global__ void __launch_bounds__(32 * QX_WARPS) lp_exact_query_kernel(LpRescoreArgs a) {
    extern __shared__ __align__(16) unsigned char qx_smem[];
    float (*rows_s)[32][RS_LD] = reinterpret_cast<float (*)[32][RS_LD]>(qx_smem);                       // [warp][32 keys][64 + 4]
    float* q_s = reinterpret_cast<float*>(qx_smem + sizeof(float) * QX_WARPS * 32 * RS_LD);             // [C]
    float* lv = q_s + a.C;                                                                              // [256][QX_K]
    int* li = reinterpret_cast<int*>(lv + 32 * QX_WARPS * QX_K);
    __shared__ int row_start[32], row_x0[32], row_y[32];
    __shared__ int n_rows_s, disc_s;
    __shared__ float red_v[QX_WARPS];
    __shared__ int red_i[QX_WARPS], red_t[QX_WARPS];
    __shared__ float win_v[QX_K];
    __shared__ int win_i[QX_K];
    const unsigned qcount = a.hdr[HDR_FLAGGED];
    if (qcount > LP_QX_CAP) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hw = a.h * a.w;
    __shared__ unsigned last_s;
    for (unsigned work = blockIdx.x; work < qcount * QX_SPLIT; work += gridDim.x) {
        const unsigned item = work / QX_SPLIT, part = work - item * QX_SPLIT;
        const int2 nq = a.qlist[item];
        const int n = nq.x, qpos = nq.y;
        const int qy = qpos / a.w, qx = qpos - qy * a.w;
        __syncthreads();
        if (tid == 0) {
            int cnt = 0, nr = 0;
            if (a.restricted)
                for (int dy = -a.R; dy <= a.R; ++dy) {
                    const int y = qy + dy, rem = a.r2i - dy * dy;
                    if (y < 0 || y >= a.h || rem < 0) continue;
                    const int dxm = (int)sqrtf((float)rem);
                    const int x0 = max(qx - dxm, 0), x1 = min(qx + dxm, a.w - 1);
                    row_start[nr] = cnt; row_x0[nr] = x0; row_y[nr] = y;
                    cnt += x1 - x0 + 1;
                    ++nr;
                }
            n_rows_s = nr;
            disc_s = cnt;
        }
        const float* qrow = a.feats + (a.query_frames[n] * (int64_t)hw + qpos) * a.C;
        for (int c = tid; c < a.C; c += blockDim.x) q_s[c] = qrow[c];
        __syncthreads();
        const int disc = disc_s, nr = n_rows_s;
        const int64_t total = (int64_t)a.n_unres * hw + (int64_t)(a.S - a.n_unres) * disc;
        float tv[QX_K];
        int ti[QX_K];
#pragma unroll
        for (int r = 0; r < QX_K; ++r) { tv[r] = -INFINITY; ti[r] = 0x7fffffff; }
        for (int64_t base = ((int64_t)part * QX_WARPS + warp) * 32; base < total; base += (int64_t)QX_SPLIT * QX_WARPS * 32) {
            const int64_t c = base + lane;
            const bool valid = c < total;
            int slot = 0, pos = 0;
            if (valid) {
                if (c < (int64_t)a.n_unres * hw) { slot = (int)(c / hw); pos = (int)(c - (int64_t)slot * hw); }
                else {
                    const int64_t c2 = c - (int64_t)a.n_unres * hw;
                    slot = a.n_unres + (int)(c2 / disc);
                    const int d = (int)(c2 - (int64_t)(slot - a.n_unres) * disc);
                    int r = 0;
                    while (r + 1 < nr && row_start[r + 1] <= d) ++r;
                    pos = row_y[r] * a.w + row_x0[r] + (d - row_start[r]);
                }
            }
            const float* krow = a.feats + (a.key_frames[(int64_t)n * a.S + slot] * (int64_t)hw + pos) * a.C;
            float acc = 0.f;
            for (int c0 = 0; c0 < a.C; c0 += RS_CH) {
                __syncwarp();
#pragma unroll 8
                for (int i = 0; i < 16; ++i) {
                    const int r = 2 * i + (lane >> 4);
                    const float* src = reinterpret_cast<const float*>(__shfl_sync(kFull, (unsigned long long)krow, r));
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + c0) + (lane & 15));
                    *reinterpret_cast<float4*>(&rows_s[warp][r][(lane & 15) * 4]) = v;
                }
                __syncwarp();
#pragma unroll
                for (int cc = 0; cc < RS_CH; cc += 4) {
                    const float4 kv = *reinterpret_cast<const float4*>(&rows_s[warp][lane][cc]);
                    const float4 qv = *reinterpret_cast<const float4*>(&q_s[c0 + cc]);
                    acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
                }
            }
            if (valid) {
                float x = acc / a.tau;
                int id = slot * hw + pos;
                if (x > tv[QX_K - 1]) {                       // ids ascend along a lane's stream: strict > keeps the lower index ahead
#pragma unroll
                    for (int r = 0; r < QX_K; ++r) {
                        const bool better = x > tv[r];
                        const float ov = tv[r];
                        const int oi = ti[r];
                        tv[r] = better ? x : ov;
                        ti[r] = better ? id : oi;
                        x = better ? ov : x;
                        id = better ? oi : id;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < QX_K; ++r) { lv[tid * QX_K + r] = tv[r]; li[tid * QX_K + r] = ti[r]; }
        int head = 0;
        for (int r = 0; r < a.k; ++r) {                       // k rounds of a block-wide arg-max over the list heads
            float v = head < QX_K ? lv[tid * QX_K + head] : -INFINITY;
            int id = head < QX_K ? li[tid * QX_K + head] : 0x7fffffff;
            int who = tid;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(kFull, v, o);
                const int oi = __shfl_xor_sync(kFull, id, o), ow = __shfl_xor_sync(kFull, who, o);
                if (ov > v || (ov == v && oi < id)) { v = ov; id = oi; who = ow; }
            }
            if (lane == 0) { red_v[warp] = v; red_i[warp] = id; red_t[warp] = who; }
            __syncthreads();
            float bv = red_v[0];
            int bi = red_i[0], bt = red_t[0];
            for (int w2 = 1; w2 < QX_WARPS; ++w2)
                if (red_v[w2] > bv || (red_v[w2] == bv && red_i[w2] < bi)) { bv = red_v[w2]; bi = red_i[w2]; bt = red_t[w2]; }
            if (tid == bt) ++head;
            if (tid == 0) { win_v[r] = bv; win_i[r] = bi; }
            __syncthreads();
        }
        // deliver this CTA's k winners; the CTA that delivers last merges the QX_SPLIT partial lists (any delivery order gives
        // the same result: the merge ranks by (score desc, index asc)) and writes the query's output
        if (tid == 0) {
            for (int r = 0; r < a.k; ++r) {
                a.qpart_v[((size_t)item * QX_SPLIT + part) * QX_K + r] = win_v[r];
                a.qpart_i[((size_t)item * QX_SPLIT + part) * QX_K + r] = win_i[r];
            }
            __threadfence();
            last_s = atomicAdd(a.qticket + item, 1u) == QX_SPLIT - 1 ? 1u : 0u;
        }
        __syncthreads();
        if (last_s && tid == 0) {
            __threadfence();
            const volatile float* pv = a.qpart_v + (size_t)item * QX_SPLIT * QX_K;
            const volatile int* pi = a.qpart_i + (size_t)item * QX_SPLIT * QX_K;
            int heads[QX_SPLIT];
            for (int p2 = 0; p2 < QX_SPLIT; ++p2) heads[p2] = 0;
            float vals[QX_K];
            int ids[QX_K];
            for (int r = 0; r < a.k; ++r) {
                float bv = -INFINITY;
                int bi = 0x7fffffff, bp = 0;
                for (int p2 = 0; p2 < QX_SPLIT; ++p2) {
                    if (heads[p2] >= a.k) continue;
                    const float v = pv[p2 * QX_K + heads[p2]];
                    const int id = pi[p2 * QX_K + heads[p2]];
                    if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; bp = p2; }
                }
                heads[bp]++;
                vals[r] = bv;
                ids[r] = bi;
            }
            const int64_t o = (int64_t)n * a.k * hw + qpos;
            float den = 0.f;
            const float mxv = vals[0];
            for (int r = 0; r < a.k; ++r) { vals[r] = expf(vals[r] - mxv); den += vals[r]; }
            for (int r = 0; r < a.k; ++r) {
                a.Ws[o + (int64_t)r * hw] = vals[r] / den;
                a.Is[o + (int64_t)r * hw] = ids[r] == 0x7fffffff ? 0 : ids[r];
            }
            a.qticket[item] = 0u;                             // ready for the next call
        }
    }
}

__global__ void lp_fill_list_kernel(int* list, unsigned* count, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = i;
    if (i == 0) { count[0] = (unsigned)n; count[HDR_FLAGGED - HDR_COUNT] = LP_QX_CAP + 1u; }      // every tile, on the tensor cores
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

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static int lp_tiles(int h, int w) { return ((w + TC_QW - 1) / TC_QW) * ((h + TC_QH - 1) / TC_QH); }

size_t lp_tc_workspace_bytes(int Nf, int Nt, int h, int w, int C) {
    const size_t plane = align256((size_t)Nf * h * w * C * 2);
    const size_t sl = align256((size_t)Nt * h * w * LP_SHORT * 4);
    const size_t tl = align256((size_t)Nt * lp_tiles(h, w) * 4);
    return HDR_BYTES + 2 * plane + 2 * sl + 2 * tl + align256(LP_QX_CAP * sizeof(int2)) + 2 * align256((size_t)LP_QX_CAP * QX_SPLIT * QX_K * 4) +
           align256(LP_QX_CAP * 4);
}

bool lp_tc_supported(int C, int k, float radius, int R, bool dense) {
    // k <= 12: the 16-key shortlist must hold spare entries below the k-th for the certification to mean anything
    return !dense && C % 64 == 0 && C >= 64 && C <= 256 && k >= 1 && k <= LP_PRE_MAX_K && (radius <= 0.f || TC_QW + 2 * R <= TC_NS);
}

template <int MODE>
static int launch_tc_pass(const CUtensorMap& mqh, const CUtensorMap& mql, const CUtensorMap& mkh, const CUtensorMap& mkl,
                          const LpTcArgs& a, cudaStream_t st) {
    constexpr int PLANES = MODE == MODE_EXACT ? 2 : 1, STAGES = MODE == MODE_EXACT ? 3 : 6;
    const size_t ring = (size_t)STAGES * PLANES * TC_NT * a.C * 2, stagingq = (size_t)PLANES * TC_M * a.C * 2;
    const size_t lists = (size_t)TC_M * 2 * LP_HALF * 8;                 // parked in the ring once the MMAs are done
    size_t body = ring > stagingq ? ring : stagingq;
    body = (body > lists ? body : lists) + (size_t)32 * 32 * TC_EPI_WARPS * 4 + (size_t)TC_M * 2 * 4;
    const size_t smem = 1024 + body + 256;
    if (smem > 227 * 1024) { set_error("lp_topk: shared memory budget exceeded (%zu bytes)", smem); return CRW_ERR_UNSUPPORTED; }
    auto k = lp_topk_tc_kernel<MODE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = lp_tiles(a.h, a.w);
    const int listed_grid = tiles * a.Nt < 148 ? tiles * a.Nt : 148;
    dim3 grid = a.list ? dim3((unsigned)listed_grid, 1) : dim3((unsigned)tiles, (unsigned)a.Nt);
    k<<<grid, TC_THREADS, smem, st>>>(mqh, mql, mkh, mkl, a);
    return check_launch(MODE == MODE_EXACT ? "lp_topk_tc<exact>" : "lp_topk_tc<pre>");
}

int launch_lp_tc(const float* feats, int Nf, const LpTcArgs& a0, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    const int hw = a0.h * a0.w, C = a0.C;
    cudaStream_t st = (cudaStream_t)stream;
    if (workspace_bytes < lp_tc_workspace_bytes(Nf, a0.Nt, a0.h, a0.w, C)) {
        set_error("lp_topk: workspace too small for the tensor-core path");
        return CRW_ERR_SHAPE;
    }
    const size_t plane = align256((size_t)Nf * hw * C * 2);
    const size_t sl = align256((size_t)a0.Nt * hw * LP_SHORT * 4);
    const int tiles = lp_tiles(a0.h, a0.w);
    const size_t tl = align256((size_t)a0.Nt * tiles * 4);
    unsigned char* ws = (unsigned char*)workspace;
    unsigned* hdr = (unsigned*)ws;
    void* hi = ws + HDR_BYTES;
    void* lo = ws + HDR_BYTES + plane;
    float* short_v = (float*)(ws + HDR_BYTES + 2 * plane);
    int* short_i = (int*)(ws + HDR_BYTES + 2 * plane + sl);
    unsigned* flags = (unsigned*)(ws + HDR_BYTES + 2 * plane + 2 * sl);
    int* list = (int*)(ws + HDR_BYTES + 2 * plane + 2 * sl + tl);
    unsigned char* qbase = ws + HDR_BYTES + 2 * plane + 2 * sl + 2 * tl;
    int2* qlist = (int2*)qbase;
    const size_t qp = align256((size_t)LP_QX_CAP * QX_SPLIT * QX_K * 4);
    float* qpart_v = (float*)(qbase + align256(LP_QX_CAP * sizeof(int2)));
    int* qpart_i = (int*)(qbase + align256(LP_QX_CAP * sizeof(int2)) + qp);
    unsigned* qticket = (unsigned*)(qbase + align256(LP_QX_CAP * sizeof(int2)) + 2 * qp);
    cudaMemsetAsync(qticket, 0, LP_QX_CAP * 4, st);
    cudaMemsetAsync(hdr, 0, HDR_BYTES, st);
    cudaMemsetAsync(flags, 0, (size_t)a0.Nt * tiles * 4, st);
    const int64_t rows = (int64_t)Nf * hw;
    const int sgrid = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
    lp_split_kernel<<<sgrid, 256, 0, st>>>((const float4*)feats, (uint2*)hi, (uint2*)lo, rows, C / 4, hdr);
    int e = check_launch("lp_split");
    if (e != CRW_OK) return e;
    CUtensorMap mqh, mql, mkh, mkl;
    if (!make_map(&mqh, hi, C, rows, TC_QW) || !make_map(&mql, lo, C, rows, TC_QW) || !make_map(&mkh, hi, C, rows, TC_NS) ||
        !make_map(&mkl, lo, C, rows, TC_NS)) {
        set_error("lp_topk: cuTensorMapEncodeTiled failed");
        return CRW_ERR_CUDA;
    }
    LpTcArgs a = a0;
    a.err = hdr + HDR_ERR;
    a.short_v = short_v;
    a.short_i = short_i;
    LpRescoreArgs r{};
    r.feats = feats; r.key_frames = a.key_frames; r.query_frames = a.query_frames; r.short_v = short_v; r.short_i = short_i;
    r.Nt = a.Nt; r.S = a.S; r.h = a.h; r.w = a.w; r.C = C; r.k = a.k; r.tau = a.tau; r.Ws = a.Ws; r.Is = a.Is;
    r.hdr = hdr; r.flags = flags; r.list = list; r.qlist = qlist; r.qpart_v = qpart_v; r.qpart_i = qpart_i; r.qticket = qticket;
    r.n_unres = a.restricted ? a.n_long : a.S; r.restricted = a.restricted; r.R = a.R; r.r2i = a.r2i;
    const bool pre = !(a.flags & CRW_LP_EXACT_ONLY);
    if (pre) {
        a.list = nullptr; a.count = nullptr;
        e = launch_tc_pass<MODE_PRE>(mqh, mql, mkh, mkl, a, st);
        if (e != CRW_OK) return e;
        const int64_t npairs = ((int64_t)a.Nt * hw + 1) / 2;
        const int64_t want = (npairs + RS_WARPS - 1) / RS_WARPS;
        const int rgrid = (int)(want < 148 * 12 ? want : 148 * 12);
        lp_rescore_kernel<1><<<rgrid, 32 * RS_WARPS, 0, st>>>(r);
        e = check_launch("lp_rescore<check>");
        if (e != CRW_OK) return e;
        {   // few open queries: exact fp32, one CTA each
            const size_t qsm = sizeof(float) * ((size_t)QX_WARPS * 32 * RS_LD + C + 32 * QX_WARPS * QX_K) + sizeof(int) * 32 * QX_WARPS * QX_K;
            cudaFuncSetAttribute(lp_exact_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsm);
            lp_exact_query_kernel<<<148 * 4, 32 * QX_WARPS, qsm, st>>>(r);
            e = check_launch("lp_exact_query");
            if (e != CRW_OK) return e;
        }
        a.list = list; a.count = hdr + HDR_COUNT;
        e = launch_tc_pass<MODE_EXACT>(mqh, mql, mkh, mkl, a, st);
        if (e != CRW_OK) return e;
        lp_rescore_kernel<0><<<tiles * a.Nt < 148 * 4 ? tiles * a.Nt : 148 * 4, 32 * RS_WARPS, 0, st>>>(r);
        return check_launch("lp_rescore<listed>");
    }
    // every tile on the fp32-faithful path (on request): fill the work list with all tiles first
    a.list = nullptr; a.count = nullptr;
    e = launch_tc_pass<MODE_EXACT>(mqh, mql, mkh, mkl, a, st);
    if (e != CRW_OK) return e;
    lp_fill_list_kernel<<<(tiles * a.Nt + 255) / 256, 256, 0, st>>>(list, hdr + HDR_COUNT, tiles * a.Nt);
    lp_rescore_kernel<0><<<tiles * a.Nt < 148 * 4 ? tiles * a.Nt : 148 * 4, 32 * RS_WARPS, 0, st>>>(r);
    return check_launch("lp_rescore<all>");
}

#else   // CRW_SIM: the tensor-core path needs real hardware; the host simulator always takes the SIMT kernel

size_t lp_tc_workspace_bytes(int, int, int, int, int) { return 256; }
bool lp_tc_supported(int, int, float, int, bool) { return false; }
int launch_lp_tc(const float*, int, const LpTcArgs&, void*, size_t, crw_stream_t) { return CRW_ERR_UNSUPPORTED; }

#endif

}  // namespace crw
