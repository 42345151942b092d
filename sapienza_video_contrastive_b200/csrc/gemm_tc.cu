// gemm_tc.cu - batched fp32-faithful GEMM on tcgen05 for the large-graph walk (superpixel / fine-stride graphs, N >= 128):
// the chain products P_j = P_{j-1} X_j, W_j = P_j S_j, the reverse sweep and the affinity / dQ contractions of
// code/model.py:366-413 when the transition matrices no longer fit one SM (walk_general.cu).
//
//   C[z] (M x N, fp32) (+)= sum over terms t of  A_t[z] (M x K_t)  *  B_t[z] (K_t x N)
//
// Operands arrive as fp32 with arbitrary (row, column) strides.  A pre-pass (`tc_split_kernel`) rewrites each operand as
// two K-major fp16 planes, hi = fp16(s x) and lo = fp16((s x - hi) 2^11), where s is a per-ROW power of two chosen from
// the row's max |x|, so that gradient-scale rows keep their relative precision in fp16's narrow exponent range;
// transposed operands are transposed through shared memory in the same pass.  The GEMM kernel then runs
// a.b = ahi.bhi + (ahi.blo + alo.bhi) 2^-11  with fp32 accumulation in two TMEM accumulators (three tcgen05.mma per
// 16-wide K step, kind::f16, M128 N128), fed by a 3-stage TMA ring of 128B-swizzled 64-wide K chunks, and the epilogue
// un-scales (exactly: powers of two) and writes fp32.  Dropped lo.lo term: ~2^-22 relative to |a||b|.
#include "tc_common.cuh"
#include "gemm_tc.cuh"

namespace crw {

constexpr int kSplitMaxK = 4096;          // rows of a transposed operand are staged in shared memory, 8 at a time

#ifndef CRW_SIM

constexpr int GT_M = 128, GT_N = 128, GT_K = 64, GT_STAGES = 3, GT_THREADS = 192;

// ---- fp32 (strided, possibly transposed) -> two K-major fp16 planes [z][R][Kp] + one power-of-two exponent per row -------
// `tc_split_kernel` finds each row's max |x| = f 2^e (f in [0.5, 1)) and writes hi = fp16(x 2^-e), lo = fp16((x 2^-e - hi) 2^11).
// Per-ROW scaling keeps every row of either operand at full relative precision however far apart the rows' magnitudes are
// (loss gradients next to probabilities), and the epilogue undoes it exactly: C[r][c] = acc * 2^(ea[r] + eb[c]).
__device__ __forceinline__ int tc_exponent_of(int maxbits) {
    const int biased = (maxbits >> 23) & 0xff;
    if (biased == 0 || biased == 255) return 0;              // zero / subnormal / non-finite rows stay unscaled
    const int e = biased - 126;
    return e < -100 ? -100 : (e > 100 ? 100 : e);
}
__device__ __forceinline__ float tc_pow2(int e) { return __int_as_float((127 + e) << 23); }

struct SplitSide {
    TcOperand x[2][2];           // [group][term], as a (rows x K_t) operand
    int R;                       // rows of this operand's K-major plane (M for A, N for B)
    int* expo;                   // [Z][R] row exponents e: the plane holds x 2^-e
    __half* hi;
    __half* lo;
};
struct SplitArgs {
    SplitSide s[2];
    int K[2], Kp[2];             // per term: K and K rounded up to 8; a plane row is [term 0 | term 1], Kcat = Kp[0] + Kp[1] wide
    int Kcat, nj, Z, zpg;        // Z = matrices over all groups, zpg = matrices per group
};

__device__ __forceinline__ void tc_split_store(__half* hi, __half* lo, int64_t o, float v0, float v1) {
    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
    const __half l0 = __float2half_rn((v0 - __half2float(h0)) * 2048.0f), l1 = __float2half_rn((v1 - __half2float(h1)) * 2048.0f);
    *reinterpret_cast<__half2*>(hi + o) = __halves2half2(h0, h1);
    *reinterpret_cast<__half2*>(lo + o) = __halves2half2(l0, l1);
}

// One launch converts both operands of every group.  grid (ceil(Rmax / 8), 2 Z), 256 threads: a block owns 8 plane rows
// over the whole K range of all terms, one row per warp, so the row maxima never leave the warp and the source is read
// exactly once.  A term's row of up to 64 kRegs floats stays in registers between the max and the split.  Transposed
// operands (rows are the source's fast dimension) are first staged through shared memory with 32-byte-sector reads
// (8 rows x 4 k per warp instruction).
__device__ __forceinline__ int tc_split_stride(int K) { return ((K + 31) & ~31) + 4; }      // = 4 mod 32: conflict-free staging

template <int kRegs, int NTERMS>
__global__ void __launch_bounds__(256) tc_split_kernel(SplitArgs a) {
    extern __shared__ float stage[];                          // row-fast sources only: [term][8][tc_split_stride(K_t)]
    const int side = blockIdx.y >= a.Z, z = blockIdx.y - side * a.Z;
    const SplitSide& S = a.s[side];
    const int r0 = blockIdx.x * 8;
    if (r0 >= S.R) return;
    const int grp = z / a.zpg, zl = z - grp * a.zpg;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* row[NTERMS];
    int64_t cs[NTERMS];
    bool staged = false;
    int soff = 0;
#pragma unroll
    for (int t = 0; t < NTERMS; ++t) {
        const TcOperand& x = S.x[grp][t];
        const float* base = x.p + (int64_t)(zl / a.nj) * x.sb + (int64_t)(zl % a.nj) * x.sj;
        if (x.rs == 1 && x.cs != 1) {
            const int stride = tc_split_stride(a.K[t]);
            const int rr = threadIdx.x & 7, kk = threadIdx.x >> 3;
            if (r0 + rr < S.R)
                for (int k = kk; k < a.K[t]; k += 32) stage[soff + rr * stride + k] = __ldg(base + r0 + rr + (int64_t)k * x.cs);
            row[t] = stage + soff + ty * stride;
            cs[t] = 1;
            soff += 8 * stride;
            staged = true;
        } else {
            row[t] = base + (int64_t)(r0 + ty) * x.rs;
            cs[t] = x.cs;
        }
    }
    if (staged) __syncthreads();
    const int r = r0 + ty;
    if (r >= S.R) return;
    float2 v[NTERMS][kRegs];
    float m = 0.f;
    bool inreg[NTERMS];
#pragma unroll
    for (int t = 0; t < NTERMS; ++t) {
        inreg[t] = a.Kp[t] <= 64 * kRegs;
        if (inreg[t]) {
#pragma unroll
            for (int q = 0; q < kRegs; ++q) {
                const int k = q * 64 + 2 * tx;
                v[t][q].x = k < a.K[t] ? row[t][(int64_t)k * cs[t]] : 0.f;
                v[t][q].y = k + 1 < a.K[t] ? row[t][(int64_t)(k + 1) * cs[t]] : 0.f;
                m = fmaxf(m, fmaxf(fabsf(v[t][q].x), fabsf(v[t][q].y)));
            }
        } else {
            for (int k = tx; k < a.K[t]; k += 32) m = fmaxf(m, fabsf(row[t][(int64_t)k * cs[t]]));
        }
    }
    m = warp_max(m);
    const int e = tc_exponent_of(__float_as_int(m));
    const float sc = tc_pow2(-e);
    if (tx == 0) S.expo[(int64_t)z * S.R + r] = e;
    int64_t o = ((int64_t)z * S.R + r) * a.Kcat;
#pragma unroll
    for (int t = 0; t < NTERMS; ++t) {
        if (inreg[t]) {
#pragma unroll
            for (int q = 0; q < kRegs; ++q) {
                const int k = q * 64 + 2 * tx;
                if (k < a.Kp[t]) tc_split_store(S.hi, S.lo, o + k, v[t][q].x * sc, v[t][q].y * sc);   // Kp is even: the pair is inside the plane
            }
        } else {
            for (int k = 2 * tx; k < a.Kp[t]; k += 64) {
                const float v0 = k < a.K[t] ? row[t][(int64_t)k * cs[t]] : 0.f, v1 = k + 1 < a.K[t] ? row[t][(int64_t)(k + 1) * cs[t]] : 0.f;
                tc_split_store(S.hi, S.lo, o + k, v0 * sc, v1 * sc);
            }
        }
        o += a.Kp[t];
    }
}

struct GtArgs {
    float* C[2];                 // per group
    int64_t csb[2], csj[2], ldc[2];
    int accumulate[2];
    int M, N, nj, zpg, K;        // K = concatenated width of the operand planes
    const int* aexp;             // per-row exponents of the operands, [z][M] and [z][N]
    const int* bexp;
    unsigned* err;
};

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap a_hi0, const __grid_constant__ CUtensorMap a_lo0,
               const __grid_constant__ CUtensorMap b_hi0, const __grid_constant__ CUtensorMap b_lo0, GtArgs g) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);     // stays in the shared state space: LDS / STS, not generic LD / ST
    constexpr unsigned kTile = GT_M * 128;                 // one plane of one operand: 128 rows x 128 B
    constexpr unsigned kStage = 4 * kTile;                 // A hi, A lo, B hi, B lo
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_STAGES * kStage);
    uint64_t* full = bars;
    uint64_t* empty = bars + GT_STAGES;
    uint64_t* done = empty + GT_STAGES;
    unsigned* tmem_base_smem = reinterpret_cast<unsigned*>(done + 1);
    float* ub_s = reinterpret_cast<float*>(tmem_base_smem + 2);        // 128 column un-scale factors of this tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.z, m0 = blockIdx.y * GT_M, n0 = blockIdx.x * GT_N;
#ifdef CRW_GT_TRACE
    const bool tracer = (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
    long long* trace = reinterpret_cast<long long*>(g.err) + 8;
#define GT_MARK(i) do { if (tracer && lane == 0) trace[i] = clock64(); } while (0)
#else
#define GT_MARK(i) do { } while (0)
#endif
    if (warp == 0) GT_MARK(0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_base_smem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_base_smem;
    const int nchunks = (g.K + GT_K - 1) / GT_K;
    if (warp == 0) GT_MARK(1);

    if (warp == 0) {
        // ===================== TMA producer: lanes 0..3 issue the four planes of a stage =====================
        bool ok = true;
        for (int i = 0; ok && i < nchunks; ++i) {
            const unsigned st = i % GT_STAGES, ph = (i / GT_STAGES) & 1u;
            ok = mbar_wait(empty + st, ph ^ 1u, g.err);
            if (!ok) break;
            if (lane == 0) mbar_expect_tx(full + st, kStage);
            __syncwarp();
            if (lane < 4) {
                const CUtensorMap* m = lane == 0 ? &a_hi0 : lane == 1 ? &a_lo0 : lane == 2 ? &b_hi0 : &b_lo0;
                tma_load_3d(smem + st * kStage + lane * kTile, m, i * GT_K, lane < 2 ? m0 : n0, z, full + st);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        const unsigned idesc = umma_idesc_f16(GT_M, GT_N);
        const unsigned tb = __shfl_sync(kFull, tmem_base, 0);
        const unsigned sbase = __shfl_sync(kFull, smem_u32(smem), 0);
        bool ok = true;
        for (int i = 0; ok && i < nchunks; ++i) {
            const unsigned st = i % GT_STAGES, ph = (i / GT_STAGES) & 1u;
            ok = mbar_wait(full + st, ph, g.err);
            if (!ok) break;
            tc_fence_after();
            if (i == 0) GT_MARK(2);
            if (elect_one()) {
                const unsigned a = sbase + st * kStage;
                const unsigned la_hi = desc_lo(a), la_lo = desc_lo(a + kTile), lb_hi = desc_lo(a + 2 * kTile), lb_lo = desc_lo(a + 3 * kTile);
                const int ksteps = min(4, (g.K - i * GT_K + 15) >> 4);      // the last chunk may be mostly zero fill: skip its empty k-steps
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (kk >= ksteps) break;
                    const unsigned accum = (i | kk) ? 1u : 0u;
                    tc_mma_f16_ss(tb, desc_make(la_hi + 2u * kk), desc_make(lb_hi + 2u * kk), idesc, accum);
                    tc_mma_f16_ss(tb + 128u, desc_make(la_hi + 2u * kk), desc_make(lb_lo + 2u * kk), idesc, accum);
                    tc_mma_f16_ss(tb + 128u, desc_make(la_lo + 2u * kk), desc_make(lb_hi + 2u * kk), idesc, 1u);
                }
                tc_commit(empty + st);
                if (i == nchunks - 1) tc_commit(done);
            }
            __syncwarp();
        }
        GT_MARK(3);
    } else {
        // ===================== epilogue: TMEM -> registers -> un-scale -> fp32 C =====================
        const int quarter = warp & 3;
        const int row = m0 + quarter * 32 + lane;
        {   // while the main loop runs: this tile's column factors 2^eb -> shared memory (one per epilogue thread)
            const int cl = quarter * 32 + lane, col = n0 + cl;
            ub_s[cl] = col < g.N ? tc_pow2(__ldg(g.bexp + (int64_t)z * g.N + col)) : 1.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        bool ok = mbar_wait(done, 0, g.err);
        tc_fence_after();
        if (warp == 2) GT_MARK(4);
        const float ua = row < g.M ? tc_pow2(__ldg(g.aexp + (int64_t)z * g.M + row)) : 1.f;      // powers of two: exact
        const int grp = z / g.zpg, zl = z - grp * g.zpg;
        const int64_t ldc = g.ldc[grp];
        const int accumulate = g.accumulate[grp];
        float* cbase = g.C[grp] + (int64_t)(zl / g.nj) * g.csb[grp] + (int64_t)(zl % g.nj) * g.csj[grp];
        float* crow = cbase + (int64_t)row * ldc;
        const unsigned lane_addr = tmem_base + ((unsigned)(quarter * 32) << 16);
        const bool vec = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(cbase) & 15) == 0;
        if (ok) {
            for (int c0 = 0; c0 < GT_N; c0 += 32) {
                unsigned m[32], c[32];
                tc_ld32(lane_addr + (unsigned)c0, m);
                tc_ld32(lane_addr + 128u + (unsigned)c0, c);
                tc_wait_ld();
                if (row < g.M) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = fmaf(__uint_as_float(c[j]), 4.8828125e-4f, __uint_as_float(m[j])) * ua * ub_s[c0 + j];
                    const int col = n0 + c0;
                    if (vec && col + 32 <= g.N) {
                        float4* o = reinterpret_cast<float4*>(crow + col);
                        if (accumulate) {                    // all eight loads in flight before the first add
                            float4 old[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) old[j] = o[j];
#pragma unroll
                            for (int j = 0; j < 8; ++j) { v[4 * j] += old[j].x; v[4 * j + 1] += old[j].y; v[4 * j + 2] += old[j].z; v[4 * j + 3] += old[j].w; }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col + j < g.N) crow[col + j] = accumulate ? crow[col + j] + v[j] : v[j];
                    }
                }
            }
        }
    }
    if (warp == 2) GT_MARK(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) GT_MARK(6);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256u) : "memory");
    }
}

// the planes live at fixed workspace offsets, so the same few maps recur every step: memoise the encodes
struct MapKey {
    const void* base;
    int K, Kp, R, Z;
    bool operator==(const MapKey& o) const { return base == o.base && K == o.K && Kp == o.Kp && R == o.R && Z == o.Z; }
};
struct MapSlot {
    MapKey key;
    CUtensorMap map;
    bool used;
};
static thread_local MapSlot g_map_cache[64];

static bool make_map3_uncached(CUtensorMap* m, const void* base, int K, int Kp, int R, int Z);

static bool make_map3(CUtensorMap* m, const void* base, int K, int Kp, int R, int Z) {
    const MapKey key{base, K, Kp, R, Z};
    const size_t h = ((reinterpret_cast<uintptr_t>(base) >> 8) * 0x9E3779B97F4A7C15ull + (size_t)K * 1315423911u + (size_t)R * 2654435761u + (size_t)Z) >> 7;
    MapSlot& s = g_map_cache[h & 63];
    if (s.used && s.key == key) { *m = s.map; return true; }
    if (!make_map3_uncached(m, base, K, Kp, R, Z)) return false;
    s.key = key; s.map = *m; s.used = true;
    return true;
}

static bool make_map3_uncached(CUtensorMap* m, const void* base, int K, int Kp, int R, int Z) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)R, (cuuint64_t)Z};
    cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)Kp * 2 * (cuuint64_t)R};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#endif  // !CRW_SIM

// ---- host entry used by walk_general.cu ------------------------------------------------------------------------------
static inline size_t gt_align(size_t x) { return (x + 255) / 256 * 256; }

size_t gemm_tc_workspace_bytes(int M, int N, int Kcat, int Z) {
    // A planes 2 * Z*M*Kcat halves, B planes 2 * Z*N*Kcat halves, row exponents Z*(M+N) words; + error word
    return 256 + gt_align(2 * (size_t)Z * M * Kcat * 2) + gt_align(2 * (size_t)Z * N * Kcat * 2) + gt_align((size_t)Z * (M + N) * 4);
}

bool gemm_tc_eligible(int M, int N, int Kmin, int Kmax) {
#ifdef CRW_SIM
    (void)M; (void)N; (void)Kmin; (void)Kmax;
    return false;
#else
    return M >= 64 && N >= 64 && Kmin >= 64 && Kmax <= kSplitMaxK;
#endif
}

#ifndef CRW_SIM
template <int kRegs>
static void launch_split(const SplitArgs& sa, int nterms, dim3 grid, size_t smem, cudaStream_t st) {
    if (nterms == 1) tc_split_kernel<kRegs, 1><<<grid, 256, smem, st>>>(sa);
    else tc_split_kernel<kRegs, 2><<<grid, 256, smem, st>>>(sa);
}
#endif

int gemm_tc_run(const TcGemmCall& c, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
#ifdef CRW_SIM
    (void)c; (void)workspace; (void)workspace_bytes; (void)stream;
    return CRW_ERR_UNSUPPORTED;
#else
    if (c.ngroups < 1 || c.ngroups > 2 || c.nterms < 1 || c.nterms > 2) { set_error("gemm_tc: bad call"); return CRW_ERR_SHAPE; }
    const int zpg = c.nb * c.nj, Z = c.ngroups * zpg;
    int Kmax = 0, Kcat = 0;
    for (int t = 0; t < c.nterms; ++t) { Kmax = c.K[t] > Kmax ? c.K[t] : Kmax; Kcat += (c.K[t] + 7) & ~7; }
    if (Kmax > kSplitMaxK) { set_error("gemm_tc: K = %d exceeds %d", Kmax, kSplitMaxK); return CRW_ERR_UNSUPPORTED; }
    if (workspace_bytes < gemm_tc_workspace_bytes(c.M, c.N, Kcat, Z)) { set_error("gemm_tc: workspace too small"); return CRW_ERR_SHAPE; }
    unsigned char* ws = (unsigned char*)workspace;
    unsigned* err = (unsigned*)ws;
    size_t o = 256;
    __half* a_hi = (__half*)(ws + o); __half* a_lo = a_hi + (size_t)Z * c.M * Kcat; o += gt_align(2 * (size_t)Z * c.M * Kcat * 2);
    __half* b_hi = (__half*)(ws + o); __half* b_lo = b_hi + (size_t)Z * c.N * Kcat; o += gt_align(2 * (size_t)Z * c.N * Kcat * 2);
    int* aexp = (int*)(ws + o); int* bexp = aexp + (size_t)Z * c.M;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = 1024 + GT_STAGES * 4 * (size_t)GT_M * 128 + 256 + 512;
    auto k = gemm_tc_kernel;
    static thread_local int attr_device = -1;                // once per (thread, device)
    int device = 0;
    cudaGetDevice(&device);
    if (attr_device != device) {
        const int split_smem_max = 2 * 8 * (kSplitMaxK + 4) * 4;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(tc_split_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, split_smem_max / 2) != cudaSuccess ||
            cudaFuncSetAttribute(tc_split_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(tc_split_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * (1024 + 4) * 4) != cudaSuccess) {
            set_error("gemm_tc: %s", cudaGetErrorString(cudaGetLastError()));
            return CRW_ERR_CUDA;
        }
        attr_device = device;
    }
    // every term's operands -> one pair of K-concatenated planes per operand, one launch
    SplitArgs sa{};
    sa.s[0].R = c.M; sa.s[0].expo = aexp; sa.s[0].hi = a_hi; sa.s[0].lo = a_lo;
    sa.s[1].R = c.N; sa.s[1].expo = bexp; sa.s[1].hi = b_hi; sa.s[1].lo = b_lo;
    size_t split_smem_a = 0, split_smem_b = 0;
    for (int t = 0; t < c.nterms; ++t) {
        sa.K[t] = c.K[t]; sa.Kp[t] = (c.K[t] + 7) & ~7;
        bool ra = false, rb = false;
        for (int gi = 0; gi < c.ngroups; ++gi) {
            sa.s[0].x[gi][t] = c.grp[gi].A[t];
            TcOperand Bt = c.grp[gi].B[t];                   // as an (N x K) row operand: element (c, k) = p[c * cs + k * rs]
            const int64_t tmp = Bt.rs; Bt.rs = Bt.cs; Bt.cs = tmp;
            sa.s[1].x[gi][t] = Bt;
            ra |= c.grp[gi].A[t].rs == 1 && c.grp[gi].A[t].cs != 1;
            rb |= Bt.rs == 1 && Bt.cs != 1;
        }
        const size_t one = 8 * (size_t)(((c.K[t] + 31) & ~31) + 4) * sizeof(float);
        if (ra) split_smem_a += one;
        if (rb) split_smem_b += one;
    }
    sa.Kcat = Kcat; sa.nj = c.nj; sa.Z = Z; sa.zpg = zpg;
    const size_t split_smem = split_smem_a > split_smem_b ? split_smem_a : split_smem_b;
    if (split_smem > 200 * 1024) { set_error("gemm_tc: transposed operands too wide to stage (K = %d)", Kmax); return CRW_ERR_UNSUPPORTED; }
    const dim3 sg(((c.M > c.N ? c.M : c.N) + 7) / 8, 2 * Z);
    const int Kpm = (Kmax + 7) & ~7;
    if (Kpm <= 256) launch_split<4>(sa, c.nterms, sg, split_smem, st);
    else if (Kpm <= 512) launch_split<8>(sa, c.nterms, sg, split_smem, st);
    else if (Kpm <= 1024) launch_split<16>(sa, c.nterms, sg, split_smem, st);
    else launch_split<32>(sa, c.nterms, sg, split_smem, st);
    int e = check_launch("gemm_tc_split");
    if (e != CRW_OK) return e;
    CUtensorMap maps[4];
    if (!make_map3(&maps[0], a_hi, Kcat, Kcat, c.M, Z) || !make_map3(&maps[1], a_lo, Kcat, Kcat, c.M, Z) ||
        !make_map3(&maps[2], b_hi, Kcat, Kcat, c.N, Z) || !make_map3(&maps[3], b_lo, Kcat, Kcat, c.N, Z)) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed");
        return CRW_ERR_CUDA;
    }
    GtArgs g{};
    for (int gi = 0; gi < c.ngroups; ++gi) {
        g.C[gi] = c.grp[gi].C; g.csb[gi] = c.grp[gi].csb; g.csj[gi] = c.grp[gi].csj; g.ldc[gi] = c.grp[gi].ldc;
        g.accumulate[gi] = c.grp[gi].accumulate;
    }
    g.M = c.M; g.N = c.N; g.nj = c.nj; g.zpg = zpg; g.K = Kcat; g.aexp = aexp; g.bexp = bexp; g.err = err;
    dim3 grid((c.N + GT_N - 1) / GT_N, (c.M + GT_M - 1) / GT_M, Z);
    k<<<grid, GT_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], g);
    return check_launch("gemm_tc");
#endif
}

}  // namespace crw

using namespace crw;

extern "C" size_t crw_bmm_tc_workspace_bytes(int Z, int M, int N, int K) { return gemm_tc_workspace_bytes(M, N, (K + 7) & ~7, Z); }

extern "C" int crw_bmm_tc(const float* A, const float* B, float* C, int Z, int M, int N, int K, int trans_a, int trans_b,
                          int accumulate, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    if (Z < 0 || M <= 0 || N <= 0 || K <= 0) { set_error("bmm_tc: bad shape Z=%d M=%d N=%d K=%d", Z, M, N, K); return CRW_ERR_SHAPE; }
    if (Z == 0) return CRW_OK;
    if (!gemm_tc_eligible(M, N, K, K)) { set_error("bmm_tc: needs M >= 64, N >= 64, 64 <= K <= 4096 (got %d, %d, %d)", M, N, K); return CRW_ERR_UNSUPPORTED; }
    TcGemmCall c{};
    c.ngroups = 1; c.nterms = 1; c.K[0] = K; c.M = M; c.N = N; c.nb = Z; c.nj = 1;
    c.grp[0].accumulate = accumulate;
    c.grp[0].A[0] = TcOperand{A, (int64_t)M * K, 0, trans_a ? 1 : K, trans_a ? M : 1};
    c.grp[0].B[0] = TcOperand{B, (int64_t)K * N, 0, trans_b ? 1 : N, trans_b ? K : 1};
    c.grp[0].C = C; c.grp[0].csb = (int64_t)M * N; c.grp[0].csj = 0; c.grp[0].ldc = N;
    int e = gemm_tc_run(c, workspace, workspace_bytes, stream);
    return e;
}
