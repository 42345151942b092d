// gemm_tf32.cu - batched fp32-faithful GEMM on tcgen05 kind::tf32 with the operand preparation INSIDE the kernel, for the
// mid-size graphs of the walk (72 <= N < 512: superpixel graphs, fine-stride patch grids), where a product is too small to
// amortise a separate operand-split launch (gemm_tc.cu: ~8 us of split + ~12 us of GEMM per product at N = 196).
//
//   C_g[z] (M x N, fp32) (+)= sum over terms t of  A_gt[z] (M x K_t)  *  B_gt[z] (K_t x N)         (same call shape as gemm_tc.cu)
//
// TMA reads the fp32 operands where they lie (4-D tensor maps over the caller's strides; an operand may be K-major or
// MN-major - transposed uses need no copy, the UMMA descriptor carries the major-ness) into 128B-swizzled stage buffers
// (SWIZZLE_128B for K-major, the 32-byte-atom variant that MN-major 32-bit operands require).
// Eight converter warps then rewrite a landed stage in place as  big = tf32(x)  and, into a twin buffer with the identical
// layout,  small = tf32(x - big)  (round to nearest: unbiased; x = big + small to 2^-22 relative), fence the generic writes towards the
// async proxy and hand the stage to the MMA warp, which issues  big.big + big.small + small.big  (three tcgen05.mma per
// 8-wide k-step, all into ONE fp32 TMEM accumulator - the smalls are true values, no re-scaling, and TF32 has fp32's
// exponent range, so there are no per-row exponents, no row maxima and no epilogue un-scaling).  The terms of a two-term
// product simply continue the same main loop with the next pair of tensor maps.
//
// Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (warp-uniform loop, one elected lane), warps 2-9 = converters, then
// the epilogue (two warps per TMEM lane quarter; tcgen05.ld -> fp32 stores, optional accumulate).  3 stages x 64 KB.
#include "tc_common.cuh"
#include "gemm_tc.cuh"

namespace crw {

#ifndef CRW_SIM

constexpr int G3_M = 128, G3_N = 128, G3_K = 32, G3_STAGES = 3, G3_CONV_WARPS = 8, G3_THREADS = 64 + 32 * G3_CONV_WARPS;
constexpr unsigned kG3Tile = G3_M * 128;               // one operand tile: 128 rows (or 4 x 32 k-rows) x 128 B = 16 KB
constexpr unsigned kG3Stage = 4 * kG3Tile;             // A big, A small, B big, B small

struct G3Maps {
    CUtensorMap m[2][2][2];                            // [group][term][operand A / B]
};
struct G3Args {
    float* C[2];
    int64_t csb[2], csj[2], ldc[2];
    int accumulate[2];
    int M, N, nj, zpg, nterms;
    int K[2];
    int a_mn[2][2], b_mn[2][2];                        // [group][term]: operand is MN-major (its rows are the contiguous dimension)
    unsigned* err;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ss(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout <<61.
// K-major: SWIZZLE_128B (2): rows of 128 B (32 floats of K), 8-row groups 1024 B apart (SBO), LBO unused.
// MN-major 32-bit operands only exist as SWIZZLE_128B_BASE32B (1; TMA: SWIZZLE_128B_ATOM_32B): k-rows of 128 B (32 floats of
// M/N) swizzled in 32-byte units over 4 rows, 4-k groups 512 B apart (SBO), 32-wide M/N blocks 4096 B apart (LBO).
__device__ __forceinline__ uint64_t g3_desc(unsigned smem_addr, bool mn_major) {
    const unsigned lbo = mn_major ? (4096u >> 4) : 1u, sbo = mn_major ? (512u >> 4) : (1024u >> 4);
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)lbo << 16) | ((uint64_t)sbo << 32) |
           ((uint64_t)1u << 46) | ((uint64_t)(mn_major ? 1u : 2u) << 61);
}
// kind::tf32 instruction descriptor: D fp32 (1 << 4), A / B tf32 (2 << 7, 2 << 10), major-ness bits 15 / 16, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ unsigned g3_idesc(bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((unsigned)(G3_N >> 3) << 17) | ((unsigned)(G3_M >> 4) << 24);
}
// round to TF32 (10 explicit mantissa bits), ties away from zero - what cvt.rna.tf32.f32 computes, but on the integer
// pipe: the conversion instruction runs at a quarter of the ALU rate and made the converters the bottleneck (measured)
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

__global__ void __launch_bounds__(G3_THREADS, 1) gemm_tf32_kernel(const __grid_constant__ G3Maps maps, G3Args g) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);     // stays in the shared state space: LDS / STS, not generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G3_STAGES * kG3Stage);
    uint64_t* full = bars;                             // TMA landed
    uint64_t* conv = bars + G3_STAGES;                 // converters done
    uint64_t* empty = bars + 2 * G3_STAGES;            // MMAs of the stage retired
    uint64_t* done = bars + 3 * G3_STAGES;
    unsigned* tmem_base_smem = reinterpret_cast<unsigned*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.z, m0 = blockIdx.y * G3_M, n0 = blockIdx.x * G3_N;
    const int grp = z / g.zpg, zl = z - grp * g.zpg;
    const int zb = zl / g.nj, zj = zl - zb * g.nj;

    if (warp == 0 && lane < 2 * g.nterms)                  // the descriptors this CTA's TMA loads will need
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&maps.m[grp][lane >> 1][lane & 1])) : "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < G3_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(conv + s, 32 * G3_CONV_WARPS); mbar_init(empty + s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_base_smem)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_base_smem;
    int nch[2] = {(g.K[0] + G3_K - 1) / G3_K, g.nterms > 1 ? (g.K[1] + G3_K - 1) / G3_K : 0};
    const int total = nch[0] + nch[1];

    if (warp == 0) {
        // ===================== TMA producer =====================
        bool ok = true;
        for (int it = 0; ok && it < total; ++it) {
            const int t = it >= nch[0], i = it - t * nch[0];
            const unsigned st = it % G3_STAGES, ph = (it / G3_STAGES) & 1u;
            ok = mbar_wait(empty + st, ph ^ 1u, g.err);
            if (!ok) break;
            if (lane == 0) mbar_expect_tx(full + st, 2 * kG3Tile);
            __syncwarp();
            unsigned char* sa = smem + st * kG3Stage;
            unsigned char* sb = sa + 2 * kG3Tile;
            // lanes 0-3: the A tile (one 32k x 128-row box, or four 32-row x 32-k boxes when MN-major); lanes 4-7: the B tile
            if (lane < 8) {
                const int side = lane >> 2, q = lane & 3;
                const CUtensorMap* mp = &maps.m[grp][t][side];
                const bool mn = side ? g.b_mn[grp][t] : g.a_mn[grp][t];
                const int r0 = side ? n0 : m0;
                unsigned char* dst = side ? sb : sa;
                if (mn) tma_load_4d(dst + q * 4096, mp, r0 + 32 * q, i * G3_K, zj, zb, full + st);
                else if (q == 0) tma_load_4d(dst, mp, i * G3_K, r0, zj, zb, full + st);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const unsigned tb = __shfl_sync(kFull, tmem_base, 0);
        const unsigned sbase = __shfl_sync(kFull, smem_u32(smem), 0);
        bool ok = true;
        for (int it = 0; ok && it < total; ++it) {
            const int t = it >= nch[0], i = it - t * nch[0];
            const unsigned st = it % G3_STAGES, ph = (it / G3_STAGES) & 1u;
            ok = mbar_wait(conv + st, ph, g.err);
            if (!ok) break;
            tc_fence_after();
            if (elect_one()) {
                const bool amn = g.a_mn[grp][t] != 0, bmn = g.b_mn[grp][t] != 0;
                const unsigned idesc = g3_idesc(amn, bmn);
                const unsigned a = sbase + st * kG3Stage;
                const int ksteps = min(4, (g.K[t] - i * G3_K + 7) >> 3);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (kk >= ksteps) break;
                    const unsigned ao = amn ? 1024u * kk : 32u * kk, bo = bmn ? 1024u * kk : 32u * kk;
                    const uint64_t ab = g3_desc(a + ao, amn), as = g3_desc(a + kG3Tile + ao, amn);
                    const uint64_t bb = g3_desc(a + 2 * kG3Tile + bo, bmn), bs = g3_desc(a + 3 * kG3Tile + bo, bmn);
                    tc_mma_tf32_ss(tb, ab, bb, idesc, (it | kk) ? 1u : 0u);
                    tc_mma_tf32_ss(tb, ab, bs, idesc, 1u);
                    tc_mma_tf32_ss(tb, as, bb, idesc, 1u);
                }
                tc_commit(empty + st);
                if (it == total - 1) tc_commit(done);
            }
            __syncwarp();
        }
    } else {
        // ===================== converters (warps 2-5), then the epilogue =====================
        constexpr int NCT = 32 * G3_CONV_WARPS;
        const int ct = threadIdx.x - 64;                                  // 0..NCT-1
        bool ok = true;
        for (int it = 0; ok && it < total; ++it) {
            const unsigned st = it % G3_STAGES, ph = (it / G3_STAGES) & 1u;
            ok = mbar_wait(full + st, ph, g.err);
            if (!ok) break;
            float4* sa = reinterpret_cast<float4*>(smem + st * kG3Stage);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                float4* big = sa + side * (2 * kG3Tile / 16);
                float4* small = big + kG3Tile / 16;
#pragma unroll
                for (int u = 0; u < 1024 / NCT; ++u) {
                    const float4 x = big[ct + NCT * u];
                    float4 b, s;
                    b.x = tf32_rna(x.x); b.y = tf32_rna(x.y); b.z = tf32_rna(x.z); b.w = tf32_rna(x.w);
                    s.x = tf32_rna(x.x - b.x); s.y = tf32_rna(x.y - b.y); s.z = tf32_rna(x.z - b.z); s.w = tf32_rna(x.w - b.w);
                    big[ct + NCT * u] = b;
                    small[ct + NCT * u] = s;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core's async proxy
            mbar_arrive(conv + st);
        }
        const int quarter = warp & 3;
        const int row = m0 + quarter * 32 + lane;
        const int chalf = (warp - 2) >> 2;                               // two epilogue warps per TMEM lane quarter: 64 columns each
        ok = ok && mbar_wait(done, 0, g.err);
        tc_fence_after();
        const int64_t ldc = g.ldc[grp];
        const int accumulate = g.accumulate[grp];
        float* cbase = g.C[grp] + (int64_t)zb * g.csb[grp] + (int64_t)zj * g.csj[grp];
        float* crow = cbase + (int64_t)row * ldc;
        const unsigned lane_addr = tmem_base + ((unsigned)(quarter * 32) << 16);
        const bool vec = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(cbase) & 15) == 0;
        if (ok) {
            for (int c0 = chalf * (G3_N / 2); c0 < (chalf + 1) * (G3_N / 2); c0 += 32) {
                if (n0 + c0 >= g.N) break;
                unsigned m[32];
                tc_ld32(lane_addr + (unsigned)c0, m);
                tc_wait_ld();
                if (row < g.M) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(m[j]);
                    const int col = n0 + c0;
                    if (vec && col + 32 <= g.N) {
                        float4* o = reinterpret_cast<float4*>(crow + col);
                        if (accumulate) {
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
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128u) : "memory");
    }
}

// the operand as a (rows x K) matrix over the caller's strides; rows or K must be the unit-stride dimension
static bool g3_operand_ok(const TcOperand& x, int nb, int nj, bool* mn_major) {
    if ((reinterpret_cast<uintptr_t>(x.p) & 15) != 0) return false;
    const bool kmaj = x.cs == 1, mnmaj = x.rs == 1 && x.cs != 1;
    if (!kmaj && !mnmaj) return false;
    const int64_t outer = kmaj ? x.rs : x.cs;
    if (outer <= 0 || (outer & 3) != 0) return false;
    if (nj > 1 && (x.sj <= 0 || (x.sj & 3) != 0)) return false;
    if (nb > 1 && (x.sb <= 0 || (x.sb & 3) != 0)) return false;
    *mn_major = mnmaj;
    return true;
}

static bool g3_make_map(CUtensorMap* m, const TcOperand& x, int R, int K, int nb, int nj, bool mn_major) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const int64_t outer = mn_major ? x.cs : x.rs;
    cuuint64_t dims[4] = {(cuuint64_t)(mn_major ? R : K), (cuuint64_t)(mn_major ? K : R), (cuuint64_t)nj, (cuuint64_t)nb};
    // a dimension of extent 1 is only ever addressed at 0: its stride just has to be well-formed
    cuuint64_t strides[3] = {(cuuint64_t)outer * 4, (cuuint64_t)(nj > 1 ? x.sj : outer) * 4, (cuuint64_t)(nb > 1 ? x.sb : outer) * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(mn_major ? 32 : 128), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x.p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#endif  // !CRW_SIM

// the call can run on the fused tf32 kernel: every operand TMA-addressable in place (16-byte aligned base and strides, one
// unit-stride dimension)
bool gemm_tf32_eligible(const TcGemmCall& c) {
#ifdef CRW_SIM
    (void)c;
    return false;
#else
    if (c.ngroups < 1 || c.ngroups > 2 || c.nterms < 1 || c.nterms > 2 || c.M < 64 || c.N < 64) return false;
    for (int t = 0; t < c.nterms; ++t) {
        if (c.K[t] < 32) return false;
        for (int gi = 0; gi < c.ngroups; ++gi) {
            bool amn, bmn;
            TcOperand Bt = c.grp[gi].B[t];
            const int64_t tmp = Bt.rs; Bt.rs = Bt.cs; Bt.cs = tmp;          // as an (N x K) row operand
            if (!g3_operand_ok(c.grp[gi].A[t], c.nb, c.nj, &amn) || !g3_operand_ok(Bt, c.nb, c.nj, &bmn)) return false;
        }
    }
    return true;
#endif
}

int gemm_tf32_run(const TcGemmCall& c, unsigned* err_word, crw_stream_t stream) {
#ifdef CRW_SIM
    (void)c; (void)err_word; (void)stream;
    return CRW_ERR_UNSUPPORTED;
#else
    if (!gemm_tf32_eligible(c)) { set_error("gemm_tf32: operands not addressable by TMA in place"); return CRW_ERR_UNSUPPORTED; }
    G3Maps maps;
    G3Args g{};
    for (int t = 0; t < c.nterms; ++t) {
        g.K[t] = c.K[t];
        for (int gi = 0; gi < c.ngroups; ++gi) {
            bool amn = false, bmn = false;
            TcOperand Bt = c.grp[gi].B[t];
            const int64_t tmp = Bt.rs; Bt.rs = Bt.cs; Bt.cs = tmp;
            g3_operand_ok(c.grp[gi].A[t], c.nb, c.nj, &amn);
            g3_operand_ok(Bt, c.nb, c.nj, &bmn);
            g.a_mn[gi][t] = amn; g.b_mn[gi][t] = bmn;
            if (!g3_make_map(&maps.m[gi][t][0], c.grp[gi].A[t], c.M, c.K[t], c.nb, c.nj, amn) ||
                !g3_make_map(&maps.m[gi][t][1], Bt, c.N, c.K[t], c.nb, c.nj, bmn)) {
                set_error("gemm_tf32: cuTensorMapEncodeTiled failed");
                return CRW_ERR_CUDA;
            }
        }
    }
    for (int gi = 0; gi < c.ngroups; ++gi) {
        g.C[gi] = c.grp[gi].C; g.csb[gi] = c.grp[gi].csb; g.csj[gi] = c.grp[gi].csj; g.ldc[gi] = c.grp[gi].ldc;
        g.accumulate[gi] = c.grp[gi].accumulate;
    }
    g.M = c.M; g.N = c.N; g.nj = c.nj; g.zpg = c.nb * c.nj; g.nterms = c.nterms; g.err = err_word;
    const size_t smem = 1024 + G3_STAGES * (size_t)kG3Stage + 256;
    auto k = gemm_tf32_kernel;
    static thread_local int attr_device = -1;
    int device = 0;
    cudaGetDevice(&device);
    if (attr_device != device) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("gemm_tf32: %s", cudaGetErrorString(cudaGetLastError()));
            return CRW_ERR_CUDA;
        }
        attr_device = device;
    }
    dim3 grid((c.N + G3_N - 1) / G3_N, (c.M + G3_M - 1) / G3_M, c.ngroups * c.nb * c.nj);
    k<<<grid, G3_THREADS, smem, (cudaStream_t)stream>>>(maps, g);
    return check_launch("gemm_tf32");
#endif
}

}  // namespace crw

using namespace crw;

// test / utility entry: C[z] (M,N) (+)= op(A[z]) op(B[z]) on the fused tf32 kernel (same conventions as crw_bmm_tc)
extern "C" int crw_bmm_tf32(const float* A, const float* B, float* C, int Z, int M, int N, int K, int trans_a, int trans_b,
                            int accumulate, unsigned* err_word, crw_stream_t stream) {
    if (Z < 0 || M <= 0 || N <= 0 || K <= 0) { set_error("bmm_tf32: bad shape Z=%d M=%d N=%d K=%d", Z, M, N, K); return CRW_ERR_SHAPE; }
    if (Z == 0) return CRW_OK;
    TcGemmCall c{};
    c.ngroups = 1; c.nterms = 1; c.K[0] = K; c.M = M; c.N = N; c.nb = Z; c.nj = 1;
    c.grp[0].accumulate = accumulate;
    c.grp[0].A[0] = TcOperand{A, (int64_t)M * K, 0, trans_a ? 1 : K, trans_a ? M : 1};
    c.grp[0].B[0] = TcOperand{B, (int64_t)K * N, 0, trans_b ? 1 : N, trans_b ? K : 1};
    c.grp[0].C = C; c.grp[0].csb = (int64_t)M * N; c.grp[0].csj = 0; c.grp[0].ldc = N;
    return gemm_tf32_run(c, err_word, stream);
}
