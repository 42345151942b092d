// walk_chain_cluster.cu - the per-clip chain of walk_fused.cu (model.py:376-413 re-associated, forward + reverse sweep)
// spread over a thread-block cluster of 4 CTAs = 4 SMs per clip.
//
// The chain is the one sequential part of the small-graph walk: ~20 dependent N x N products per clip (T = 4), which a
// single CTA executes issue-bound on one SM while most of the GPU idles (20 clips -> 20 of 148 SMs).  Here every product
// is split by OUTPUT ROWS over the 4 CTAs of a cluster: CTA r computes rows [r RB, (r+1) RB) of every matrix, which needs
// its own rows (or columns, for A^T B products) of the left operand and the whole right operand.  Every CTA therefore
// keeps a full copy of every matrix in its shared memory (same layout as the single-CTA kernel): it computes its rows
// locally, one `barrier.cluster` (release / acquire) per dependency level publishes them, and every CTA then PULLS the other
// three row blocks out of its peers' shared memory with 128-bit distributed-shared-memory loads (one per thread).  Pushing
// each computed value to the three remote copies instead was measured 3x slower: a warp's remote 4-byte stores cost ~250
// cycles each to issue.  The transition stacks F_i, G_i arrive in every CTA by TMA bulk copies.  Results that leave the kernel (dX_j, dY_j) are
// written straight to the workspace by the CTA that owns the rows.  Row sums of the loss stay per CTA and are folded by
// the last CTA of the launch in (clip, rank) order: deterministic.
#include "walk.cuh"

#ifndef CRW_SIM
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#endif

namespace crw {

#ifndef CRW_SIM

constexpr int kCcThreads = 512;
constexpr int kCcWarps = kCcThreads / 32;

#define CC_COMP4(v, kk) ((kk) == 0 ? (v).x : (kk) == 1 ? (v).y : (kk) == 2 ? (v).z : (v).w)

// rows r0 .. r0+TM-1 (clipped to rhi) of C = A1 B1 (+ A2 B2); lane owns columns 2 lane, 2 lane + 1
template <int TM>
__device__ __forceinline__ void cc_nn(float* C, const float* A1, const float* B1, const float* A2, const float* B2, int N, int NP,
                                      int r0, int rhi, int lane) {
    if (r0 >= rhi) return;
    const int c0 = min(2 * lane, NP - 2);
    float2 acc[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i] = make_float2(0.f, 0.f);
    const int K4 = N >> 2;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float* A = t ? A2 : A1;
        const float* B = t ? B2 : B1;
        if (!A) continue;
        const float* arow[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) arow[i] = A + min(r0 + i, rhi - 1) * NP;
        const float* bp = B + c0;
#pragma unroll 4
        for (int k4 = 0; k4 < K4; ++k4) {
            float4 a4[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a4[i] = *reinterpret_cast<const float4*>(arow[i] + k4 * 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float2 b2 = *reinterpret_cast<const float2*>(bp + (k4 * 4 + kk) * NP);
#pragma unroll
                for (int i = 0; i < TM; ++i) acc[i] = ffma2(CC_COMP4(a4[i], kk), b2, acc[i]);
            }
        }
        for (int k = K4 * 4; k < N; ++k) {
            const float2 b2 = *reinterpret_cast<const float2*>(bp + k * NP);
#pragma unroll
            for (int i = 0; i < TM; ++i) acc[i] = ffma2(arow[i][k], b2, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = r0 + i;
        if (r >= rhi) continue;
        if (2 * lane < N) C[r * NP + 2 * lane] = acc[i].x;
        if (2 * lane + 1 < N) C[r * NP + 2 * lane + 1] = acc[i].y;
    }
}

// rows of C = A1 B1^T (+ A2 B2^T); lane owns columns lane and lane + 32 (rows of B)
template <int TM>
__device__ __forceinline__ void cc_nt(float* C, const float* A1, const float* B1, const float* A2, const float* B2, int N, int NP,
                                      int r0, int rhi, int lane) {
    if (r0 >= rhi) return;
    float2 acc[TM][2];                    // (even k, odd k) partial sums
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
    const int K4 = N >> 2;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float* A = t ? A2 : A1;
        const float* B = t ? B2 : B1;
        if (!A) continue;
        const float* arow[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) arow[i] = A + min(r0 + i, rhi - 1) * NP;
        const float* brow[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) brow[j] = B + min(lane + 32 * j, N - 1) * NP;
#pragma unroll 4
        for (int k4 = 0; k4 < K4; ++k4) {
            float4 b4[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) b4[j] = *reinterpret_cast<const float4*>(brow[j] + k4 * 4);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float4 a4 = *reinterpret_cast<const float4*>(arow[i] + k4 * 4);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float2 s = acc[i][j];
                    s = ffma2v(make_float2(a4.x, a4.y), make_float2(b4[j].x, b4[j].y), s);
                    s = ffma2v(make_float2(a4.z, a4.w), make_float2(b4[j].z, b4[j].w), s);
                    acc[i][j] = s;
                }
            }
        }
        for (int k = K4 * 4; k < N; ++k) {
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j].x = fmaf(arow[i][k], brow[j][k], acc[i][j].x);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = r0 + i;
        if (r >= rhi) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = lane + 32 * j;
            if (c < N) C[r * NP + c] = acc[i][j].x + acc[i][j].y;
        }
    }
}

// rows of C = A1^T B1 (+ A2^T B2): C[r][c] = sum_k A[k][r] B[k][c]; lane owns columns 2 lane, 2 lane + 1
template <int TM>
__device__ __forceinline__ void cc_tn(float* C, const float* A1, const float* B1, const float* A2, const float* B2, int N, int NP,
                                      int r0, int rhi, int lane) {
    if (r0 >= rhi) return;
    const int c0 = min(2 * lane, NP - 2);
    float2 acc[TM];
    int rr[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) { acc[i] = make_float2(0.f, 0.f); rr[i] = min(r0 + i, rhi - 1); }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float* A = t ? A2 : A1;
        const float* B = t ? B2 : B1;
        if (!A) continue;
        const float* bp = B + c0;
#pragma unroll 4
        for (int k = 0; k < N; ++k) {
            const float2 b2 = *reinterpret_cast<const float2*>(bp + k * NP);
#pragma unroll
            for (int i = 0; i < TM; ++i) acc[i] = ffma2(A[k * NP + rr[i]], b2, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = r0 + i;
        if (r >= rhi) continue;
        if (2 * lane < N) C[r * NP + 2 * lane] = acc[i].x;
        if (2 * lane + 1 < N) C[r * NP + 2 * lane + 1] = acc[i].y;
    }
}


__global__ void __launch_bounds__(kCcThreads, 1) walk_chain_cluster_kernel(WalkParams p) {
    CRW_DYN_SMEM(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    float* smem = reinterpret_cast<float*>(smem_raw);
    const int N = p.N, T = p.T;
    const FusedLayout L = fused_layout(N, T, p.D);
    const int NP = L.NP, MS = L.MS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / kChainCluster;
    const bool flip = (p.flags & CRW_WALK_FLIP) != 0;
    const int RB = (N + kChainCluster - 1) / kChainCluster;
    const int rlo = min(rank * RB, N), rhi = min(rlo + RB, N);
    float* Fm = smem;                               // same carve-up as walk_chain_kernel
    float* Gm = Fm + (T - 1) * MS;
    float* Pm = Gm + (T - 1) * MS;
    float* Sm = Pm + (T - 2) * MS;
    float* scratch = Sm + (T - 2) * MS;
    float* red = scratch + 3 * MS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * kCcWarps + 2);
    const float* peer[kChainCluster];               // the four shared-memory windows of the cluster, by rank
#pragma unroll
    for (int i = 0; i < kChainCluster; ++i) peer[i] = i == rank ? smem : cluster.map_shared_rank(smem, i);
    // after a cluster barrier: fetch the row blocks the three peers computed into this CTA's copy (128-bit DSMEM loads)
    auto pull = [&](float* m0, float* m1) {
        const int q4 = NP >> 2, per = N * q4;
        for (int v = tid; v < (m1 ? 2 : 1) * per; v += kCcThreads) {
            const int which = v >= per, u = v - which * per;
            const int r = u / q4, c4 = u - r * q4, owner = r / RB;
            if (owner == rank) continue;
            float* m = which ? m1 : m0;
            const int off = (int)(m - smem) + r * NP + 4 * c4;
            *reinterpret_cast<float4*>(smem + off) = *reinterpret_cast<const float4*>(peer[owner] + off);
        }
        __syncthreads();
    };

    const int64_t cm = (int64_t)b * (T - 1) * MS;
    bulk_bar_init(bar, tid);
    {
        const unsigned mbytes = (unsigned)(MS * sizeof(float));
        bulk_expect(bar, 2u * (unsigned)(T - 1) * mbytes, tid);
        for (int i = 0; i < T - 1; ++i) {
            bulk_copy(Fm + i * MS, p.ws_F + cm + (int64_t)i * MS, mbytes, bar, tid);
            bulk_copy(Gm + i * MS, p.ws_G + cm + (int64_t)i * MS, mbytes, bar, tid);
        }
        bulk_wait(bar, 0);
    }
    cluster.sync();                                 // every CTA of the cluster is resident before the first remote store

    float* Xm = flip ? Gm : Fm;
    float* Ym = flip ? Fm : Gm;
    float* dXg = (flip ? p.ws_dG : p.ws_dF) + cm;
    float* dYg = (flip ? p.ws_dF : p.ws_dG) + cm;
    auto Pj = [&](int j) { return j == 0 ? Xm : Pm + (j - 1) * MS; };
    auto Sj = [&](int j) { return j == 0 ? Ym : Sm + (j - 1) * MS; };
    const int grp = warp >> 3, gw = warp & 7;       // two 8-warp groups run independent products, two rows per warp
    const int r2 = rlo + 2 * gw;                    // this warp's rows when a group of 8 warps covers the block (RB <= 16)
    const int r1 = rlo + warp;                      // this warp's row when all 16 warps cover the block
    const float* kNone = nullptr;

    for (int j = 1; j <= T - 2; ++j) {
        if (grp == 0) cc_nn<2>(Pj(j), Pj(j - 1), Xm + j * MS, kNone, kNone, N, NP, r2, rhi, lane);
        else          cc_nn<2>(Sj(j), Ym + j * MS, Sj(j - 1), kNone, kNone, N, NP, r2, rhi, lane);
        cluster.sync();
        pull(Pj(j), Sj(j));
    }

    float* freeb[2 * kFusedMaxT + 4];
    int nfree = 0;
    freeb[nfree++] = scratch;
    freeb[nfree++] = scratch + MS;
    freeb[nfree++] = scratch + 2 * MS;
    float* gP = nullptr;
    float* gS = nullptr;
    const float cgrad = 1.0f / ((float)(T - 2) * (float)p.B * (float)N);
    float* dWpre = nullptr;                          // W_j rows already computed in the tail of the previous round
    for (int j = T - 2; j >= 1; --j) {
        float* dW;
        {
            if (dWpre) dW = dWpre;
            else {
                dW = freeb[--nfree];
                cc_nn<1>(dW, Pj(j), Sj(j), kNone, kNone, N, NP, r1, rhi, lane);        // own rows of W_j
                __syncwarp();
            }
            // loss of the warp's row (model.py:395-397) and dW in place
            float lsum = 0.f, asum = 0.f;
            const int n = r1;
            if (n < rhi) {
                float w[2];
                float rs = 0.f, best = -INFINITY;
                int bi = 0x7fffffff;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = lane + 32 * h;
                    w[h] = 0.f;
                    if (m < N) {
                        w[h] = dW[n * NP + m];
                        rs += w[h] + kEpsLog;
                        if (w[h] > best) { best = w[h]; bi = m; }
                    }
                }
                rs = warp_sum(rs);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {       // argmax, first maximum wins
                    const float ob = __shfl_xor_sync(kFull, best, o);
                    const int oi = __shfl_xor_sync(kFull, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                const float dg = __shfl_sync(kFull, (n >> 5) ? w[1] : w[0], n & 31) + kEpsLog;
                if (lane == 0) {
                    lsum = logf(rs) - logf(dg);
                    asum = (bi == n) ? 1.f : 0.f;
                }
                const float ir = 1.0f / rs, idg = 1.0f / dg;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = lane + 32 * h;
                    if (m < N) dW[n * NP + m] = cgrad * (ir - (m == n ? idg : 0.f));
                }
            }
            if (lane == 0) { red[warp] = lsum; red[kCcWarps + warp] = asum; }
            cluster.sync();                                 // every CTA's rows of dW_j are final (and red[] visible locally)
            pull(dW, nullptr);
            if (tid == 0) {
                float l = 0.f, a = 0.f;
                for (int w2 = 0; w2 < kCcWarps; ++w2) { l += red[w2]; a += red[kCcWarps + w2]; }
                float* part = p.ws_partial + (((int64_t)b * (T - 2) + (j - 1)) * kChainCluster + rank) * 2;
                part[0] = l;
                part[1] = a;
            }
        }
        // gP_j = dW_j S_j^T + gP_{j+1} X_{j+1}^T ;  gS_j = P_j^T dW_j + Y_{j+1}^T gS_{j+1}
        float* nP = freeb[--nfree];
        float* nS = freeb[--nfree];
        if (grp == 0)
            cc_nt<2>(nP, dW, Sj(j), gP ? gP : kNone, gP ? Xm + (j + 1) * MS : kNone, N, NP, r2, rhi, lane);
        else
            cc_tn<2>(nS, Pj(j), dW, gS ? Ym + (j + 1) * MS : kNone, gS ? gS : kNone, N, NP, r2, rhi, lane);
        cluster.sync();                                     // every CTA's rows of gP_j, gS_j are final
        pull(nP, nS);
        freeb[nfree++] = dW;
        if (gP) freeb[nfree++] = gP;
        if (gS) freeb[nfree++] = gS;
        freeb[nfree++] = Pj(j);
        freeb[nfree++] = Sj(j);
        gP = nP;
        gS = nS;
        // Tail of the round, four independent products on four warp groups at once (a product's latency, not its
        // throughput, is what a round costs):  dX_j = P_{j-1}^T gP_j  and  dY_j = gS_j S_{j-1}^T  go straight to the
        // workspace; next to them either the next round's W_{j-1} = P_{j-1} S_{j-1} or, after the last round, the walk's
        // first-frame gradients gP_0 = gP_1 X_1^T and gS_0 = Y_1^T gS_1 (= dX_0, dY_0).
        // No barrier before it: the buffers just freed are read by nobody any more (peers finished pulling them before the
        // barrier above), and the ones read here are not written in the next round.
        const bool last = j == 1;
        float* dWn = last ? nullptr : freeb[--nfree];
        const int q4 = warp & 3, r4 = rlo + 4 * q4;
        if (warp < 4)       cc_tn<4>(dXg + (int64_t)j * MS, Pj(j - 1), gP, kNone, kNone, N, NP, r4, rhi, lane);
        else if (warp < 8)  cc_nt<4>(dYg + (int64_t)j * MS, gS, Sj(j - 1), kNone, kNone, N, NP, r4, rhi, lane);
        else if (!last)     cc_nn<2>(dWn, Pj(j - 1), Sj(j - 1), kNone, kNone, N, NP, rlo + 2 * (warp - 8), rhi, lane);
        else if (warp < 12) cc_nt<4>(dXg, gP, Xm + MS, kNone, kNone, N, NP, r4, rhi, lane);
        else                cc_tn<4>(dYg, Ym + MS, gS, kNone, kNone, N, NP, r4, rhi, lane);
        if (last) break;
        __syncthreads();                                    // W_{j-1} rows (two per warp of the upper half) -> the row-owning warps
        dWpre = dWn;
    }

    // cross-clip reduction of the per-CTA sums by the last CTA to finish, in (clip, rank) order (deterministic);
    // xent[T-2] receives the loss itself, sum_j xent_j / (T-2) (model.py:413)
    unsigned* s_flag = reinterpret_cast<unsigned*>(red + 2 * kCcWarps);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        *s_flag = atomicAdd(p.ws_counter, 1u) == (unsigned)(p.B * kChainCluster) - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (*s_flag && warp == 0) {
        __threadfence();
        const float inv = 1.0f / ((float)p.B * (float)N);
        float tot = 0.f;
        for (int j = 0; j < T - 2; ++j) {
            float l = 0.f, a = 0.f;
            for (int b0 = 0; b0 < p.B; b0 += 32) {                     // fixed order: ranks inside a clip, chunks of 32 clips
                const int bb = b0 + lane;
                float lv = 0.f, av = 0.f;
                if (bb < p.B) {
                    const float* part = p.ws_partial + ((int64_t)bb * (T - 2) + j) * kChainCluster * 2;
#pragma unroll
                    for (int r = 0; r < kChainCluster; ++r) { lv += ld_cg(part + 2 * r); av += ld_cg(part + 2 * r + 1); }
                }
                l += warp_sum(lv);
                a += warp_sum(av);
            }
            if (lane == 0) { p.xent[j] = l * inv; p.acc[j] = a * inv; }
            tot += l * inv;
        }
        if (lane == 0) {
            p.xent[T - 2] = tot / (float)(T - 2);
            if (p.dev_state && p.rate > 0.f && !p.u12)
                p.dev_state[1] = ld_cg64(p.dev_state + 1) + (uint64_t)p.pinc * 2u * (unsigned)(T - 1);
            *p.ws_counter = 0u;
        }
    }
    cluster.sync();                                 // nobody exits while a peer may still store into its shared memory
}


#endif  // !CRW_SIM

// rows per CTA must fit one pass of 8 warps x 2 rows / 16 warps x 1 row
bool chain_cluster_fits(int N, int T) {
#ifdef CRW_SIM
    (void)N; (void)T;
    return false;
#else
    return T >= 3 && N >= 8 && (N + kChainCluster - 1) / kChainCluster <= 16;
#endif
}

int launch_walk_chain_cluster(const WalkParams& p, size_t smem_bytes, crw_stream_t stream) {
#ifdef CRW_SIM
    (void)p; (void)smem_bytes; (void)stream;
    return CRW_ERR_UNSUPPORTED;
#else
    auto k = walk_chain_cluster_kernel;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
        set_error("walk_chain_cluster: %s", cudaGetErrorString(cudaGetLastError()));
        return CRW_ERR_CUDA;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.B * kChainCluster));
    cfg.blockDim = dim3(kCcThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kChainCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k, p) != cudaSuccess) {
        set_error("walk_chain_cluster: %s", cudaGetErrorString(cudaGetLastError()));
        return CRW_ERR_CUDA;
    }
    return check_launch("walk_chain_cluster");
#endif
}

}  // namespace crw
