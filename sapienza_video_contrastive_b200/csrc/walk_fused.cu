// walk_fused.cu - one CTA per clip: the whole contrastive random walk of code/model.py:366-413, forward and
// backward, with every N x N transition matrix resident in shared memory.
//
//   phase 0  L2-normalise node vectors (model.py:118), write q
//   phase 1  per frame pair i: A_i = Q_i Q_{i+1}^T (model.py:68), edge dropout (model.py:81, incl. the
//            in-place union mask of SURVEY F4 and the transposed draw layout of F6), ZeroSoftmax / softmax rows
//            (utils/__init__.py:418-422) in both directions -> F_i (A12) and G_i (A21)
//   phase 2  prefix / suffix chains P_j = P_{j-1} X_j, S_j = Y_j S_{j-1}  (model.py:376-380 re-associated)
//   phase 3  per walk j = T-2..1: W_j = P_j S_j, log-diagonal cross-entropy + argmax accuracy (model.py:395-397),
//            dW_j, reverse accumulation through the chains, ZeroSoftmax backward, dQ, and finally the
//            normalisation backward -> grad_feats
//
// Raw affinities and the dropout codes of each pair go to a small per-clip global workspace (L2 resident)
// because they are only re-read once, elementwise, in the backward.  Used when the clip fits in 227 KB of
// shared memory (N <= ~60 at T = 4); larger clips take the multi-kernel path in walk_general.cu.
#include "walk.cuh"

namespace crw {

constexpr int kFusedThreads = 512;

// ---- small dense products on shared-memory matrices --------------------------------------------------------
// C[r][c] (+)= sum_k A(r,k) * B(k,c), 0 <= r,c,k < N.  A(r,k) = A[r*ars + k*aks], B(k,c) = B[k*bks + c*bcs],
// C row-major with stride NP.  Work is split over `nthr` threads (tid in [0,nthr)); each owns a TM x TN set of
// outputs with STRIDED rows/columns so neighbouring lanes touch neighbouring columns (bank-conflict free for
// row-major B, broadcast for A).
template <int TM, int TN>
__device__ __forceinline__ void mm_smem(float* C, int NP, const float* A, int ars, int aks, const float* B, int bks,
                                        int bcs, int N, bool accumulate, int tid, int nthr) {
    const int RS = (N + TM - 1) / TM, CS = (N + TN - 1) / TN;
    for (int t = tid; t < RS * CS; t += nthr) {
        const int tr = t / CS, tc = t - tr * CS;
        int ro[TM], co[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) ro[i] = min(tr + i * RS, N - 1) * ars;
#pragma unroll
        for (int j = 0; j < TN; ++j) co[j] = min(tc + j * CS, N - 1) * bcs;
        float acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int k = 0; k < N; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = A[ro[i] + k * aks];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = B[k * bks + co[j]];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = tr + i * RS;
            if (r >= N) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int c = tc + j * CS;
                if (c >= N) continue;
                float* p = C + r * NP + c;
                *p = accumulate ? (*p + acc[i][j]) : acc[i][j];
            }
        }
    }
}

// A_i = Qa Qb^T with rows staged in shared memory (stride DP floats, DP/4 odd, 16-byte aligned), K = D.
__device__ __forceinline__ void affinity_smem(float* C, int NP, const float* Qa, const float* Qb, int DP, int D, int N,
                                              int tid, int nthr) {
    constexpr int TM = 2, TN = 4;
    const int RS = (N + TM - 1) / TM, CS = (N + TN - 1) / TN;
    for (int t = tid; t < RS * CS; t += nthr) {
        const int tr = t / CS, tc = t - tr * CS;
        const float4* pa[TM];
        const float4* pb[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) pa[i] = reinterpret_cast<const float4*>(Qa + min(tr + i * RS, N - 1) * DP);
#pragma unroll
        for (int j = 0; j < TN; ++j) pb[j] = reinterpret_cast<const float4*>(Qb + min(tc + j * CS, N - 1) * DP);
        float acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
        for (int k = 0; k < D / 4; ++k) {
            float4 a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = pa[i][k];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = pb[j][k];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    float s = acc[i][j];
                    s = fmaf(a[i].x, b[j].x, s);
                    s = fmaf(a[i].y, b[j].y, s);
                    s = fmaf(a[i].z, b[j].z, s);
                    s = fmaf(a[i].w, b[j].w, s);
                    acc[i][j] = s;
                }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = tr + i * RS;
            if (r >= N) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int c = tc + j * CS;
                if (c < N) C[r * NP + c] = acc[i][j];
            }
        }
    }
}

// dQ accumulation into the global gradient buffer (row stride gs floats between nodes).
//   transposed == false: G[n][:] += sum_m Z[n][m] * Q[m][:]
//   transposed == true : G[m][:] += sum_n Z[n][m] * Q[n][:]
__device__ __forceinline__ void dq_update(float* G, const float* Q, int64_t gs, const float* Z, int NP, int N, int D,
                                          bool transposed, int tid, int nthr) {
    constexpr int TM = 2;
    const int RS = (N + TM - 1) / TM, CS = D / 4;
    const int zrs = transposed ? 1 : NP, zks = transposed ? NP : 1;
    for (int t = tid; t < RS * CS; t += nthr) {
        const int tr = t / CS, d4 = t - tr * CS;
        int zo[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) zo[i] = min(tr + i * RS, N - 1) * zrs;
        float4 acc[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int k = 0; k < N; ++k) {
            const float4 qv = *reinterpret_cast<const float4*>(Q + (int64_t)k * gs + d4 * 4);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float z = Z[zo[i] + k * zks];
                acc[i].x = fmaf(z, qv.x, acc[i].x);
                acc[i].y = fmaf(z, qv.y, acc[i].y);
                acc[i].z = fmaf(z, qv.z, acc[i].z);
                acc[i].w = fmaf(z, qv.w, acc[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = tr + i * RS;
            if (r >= N) continue;
            float4* p = reinterpret_cast<float4*>(G + (int64_t)r * gs + d4 * 4);
            float4 o = *p;
            o.x += acc[i].x; o.y += acc[i].y; o.z += acc[i].z; o.w += acc[i].w;
            *p = o;
        }
    }
}

__global__ void __launch_bounds__(kFusedThreads, 1) walk_fused_kernel(WalkParams p) {
    CRW_DYN_SMEM(smem_raw);
    if (p.dev_state) { p.seed = ld_cg64(p.dev_state); p.offset = ld_cg64(p.dev_state + 1); }
    float* smem = reinterpret_cast<float*>(smem_raw);
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kFusedThreads / 32;
    const int N = p.N, T = p.T, D = p.D;
    const FusedLayout L = fused_layout(N, T, D);
    const int NP = L.NP, MS = L.MS, DP = L.DP;
    const bool softmax = (p.flags & CRW_WALK_SOFTMAX) != 0;
    const bool flip = (p.flags & CRW_WALK_FLIP) != 0;
    const float tau = p.tau;

    float* Fm = smem + L.off_F;                 // F_i = A12_i, i < T-1
    float* Gm = smem + L.off_G;                 // G_i = A21_i
    float* R = smem + L.off_R;                  // phase-dependent region
    float* s12 = smem + L.off_stat;             // (T-1)*N row denominators of F
    float* s21 = s12 + (T - 1) * N;             // (T-1)*N row denominators of G
    float* invn = s21 + (T - 1) * N;            // T*N  1/max(norm,eps)
    float* nrm = invn + T * N;                  // T*N  norm
    float* red = nrm + T * N;                   // 2*NW scratch for block reductions
    unsigned char* codes = reinterpret_cast<unsigned char*>(smem + L.off_codes);   // N*N dropout codes of the current pair

    const int64_t gs = (int64_t)T * D;          // stride between nodes in feats / q / grad
    const float* fb = p.feats + (int64_t)b * N * gs;
    float* qb = p.q + (int64_t)b * N * gs;
    float* gb = p.grad ? p.grad + (int64_t)b * N * gs : nullptr;
    float* araw = p.ws_araw + (int64_t)b * (T - 1) * N * N;
    unsigned char* gcodes = p.ws_codes + (int64_t)b * (T - 1) * N * N;

    // ---- phase 0: normalise --------------------------------------------------------------------------------
    for (int row = warp; row < N * T; row += NW) {           // row = n*T + t
        const float* src = fb + (int64_t)row * D;
        float ss = 0.f;
        for (int d = lane * 4; d < D; d += 128) {
            const float4 v = *reinterpret_cast<const float4*>(src + d);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        const float nr = sqrtf(ss);
        const float den = fmaxf(nr, kEpsNorm);
        for (int d = lane * 4; d < D; d += 128) {
            float4 v = *reinterpret_cast<const float4*>(src + d);
            v.x /= den; v.y /= den; v.z /= den; v.w /= den;
            *reinterpret_cast<float4*>(qb + (int64_t)row * D + d) = v;
            if (gb) *reinterpret_cast<float4*>(gb + (int64_t)row * D + d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (lane == 0) {
            const int n = row / T, t = row - n * T;
            invn[t * N + n] = 1.0f / den;
            nrm[t * N + n] = nr;
        }
    }
    __syncthreads();

    // ---- phase 1: affinities -> transition matrices -----------------------------------------------------------
    {
        float* At = R;                         // raw affinity of the current pair
        float* Qs0 = R + MS;                   // staged frames, ping-pong
        float* Qs1 = Qs0 + N * DP;
        auto stage = [&](float* dst, int t) {
            for (int v = tid; v < N * (D / 4); v += kFusedThreads) {
                const int n = v / (D / 4), d4 = v - n * (D / 4);
                *reinterpret_cast<float4*>(dst + n * DP + d4 * 4) =
                    *reinterpret_cast<const float4*>(qb + (int64_t)n * gs + (int64_t)t * D + d4 * 4);
            }
        };
        stage(Qs0, 0);
        const int64_t numel = (int64_t)p.B * N * N;
        for (int i = 0; i < T - 1; ++i) {
            float* Qa = (i & 1) ? Qs1 : Qs0;
            float* Qb = (i & 1) ? Qs0 : Qs1;
            stage(Qb, i + 1);
            __syncthreads();
            affinity_smem(At, NP, Qa, Qb, DP, D, N, tid, kFusedThreads);
            __syncthreads();
            // forward rows (A12): row n over m.  Also produces the dropout codes (bit0: forward draw, bit1: backward draw)
            float* Fi = Fm + i * MS;
            float* Gi = Gm + i * MS;
            for (int n = warp; n < N; n += NW) {
                float xv[2], ev[2];
                float mx = -INFINITY;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = lane + 32 * h;
                    xv[h] = 0.f;
                    if (m < N) {
                        const float a = At[n * NP + m];
                        unsigned code = 0;
                        if (p.rate > 0.f) {
                            const int64_t e = ((int64_t)b * N + n) * N + m;
                            float u1, u2;
                            if (p.u12) {
                                u1 = p.u12[(int64_t)i * numel + e];
                                u2 = p.u21p[(int64_t)i * numel + e];
                            } else {
                                u1 = torch_uniform(p.seed, p.offset + (uint64_t)p.pinc * i, p.pthreads, (uint64_t)e);
                                u2 = torch_uniform(p.seed, p.offset + (uint64_t)p.pinc * (T - 1 + i), p.pthreads, (uint64_t)e);
                            }
                            code = (u1 < p.rate ? 1u : 0u) | (u2 < p.rate ? 2u : 0u);
                        }
                        codes[n * N + m] = (unsigned char)code;
                        gcodes[(int64_t)i * N * N + n * N + m] = (unsigned char)code;
                        araw[(int64_t)i * N * N + n * N + m] = a;
                        xv[h] = ((code & 1u) ? kNegDrop : a) / tau;
                        mx = fmaxf(mx, xv[h]);
                    }
                }
                float s = 0.f;
                if (softmax) {
                    mx = warp_max(mx);
#pragma unroll
                    for (int h = 0; h < 2; ++h) { ev[h] = (lane + 32 * h < N) ? expf(xv[h] - mx) : 0.f; s += ev[h]; }
                    s = warp_sum(s);
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float E = expf(xv[h]) - 1.0f;
                        ev[h] = (lane + 32 * h < N) ? E * E : 0.f;
                        s += ev[h];
                    }
                    s = warp_sum(s) + kEpsZs;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) { const int m = lane + 32 * h; if (m < N) Fi[n * NP + m] = ev[h] / s; }
                if (lane == 0) s12[i * N + n] = s;
            }
            __syncthreads();
            // backward rows (A21): row m of G = column m of A over n, union mask
            for (int m = warp; m < N; m += NW) {
                float xv[2], ev[2];
                float mx = -INFINITY;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = lane + 32 * h;
                    xv[h] = 0.f;
                    if (n < N) {
                        const float a = At[n * NP + m];
                        xv[h] = (codes[n * N + m] ? kNegDrop : a) / tau;
                        mx = fmaxf(mx, xv[h]);
                    }
                }
                float s = 0.f;
                if (softmax) {
                    mx = warp_max(mx);
#pragma unroll
                    for (int h = 0; h < 2; ++h) { ev[h] = (lane + 32 * h < N) ? expf(xv[h] - mx) : 0.f; s += ev[h]; }
                    s = warp_sum(s);
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float E = expf(xv[h]) - 1.0f;
                        ev[h] = (lane + 32 * h < N) ? E * E : 0.f;
                        s += ev[h];
                    }
                    s = warp_sum(s) + kEpsZs;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) { const int n = lane + 32 * h; if (n < N) Gi[m * NP + n] = ev[h] / s; }
                if (lane == 0) s21[i * N + m] = s;
            }
            __syncthreads();
        }
    }

    if (T < 3) {
        if (p.dev_state && tid == 0) {       // still consume the 2(T-1) draws
            __threadfence();
            if (atomicAdd(p.ws_counter, 1u) == (unsigned)p.B - 1u) {
                p.dev_state[1] = p.offset + (uint64_t)p.pinc * 2u * (unsigned)(T - 1);
                *p.ws_counter = 0u;
            }
        }
        return;
    }

    // ---- phase 2: prefix / suffix chains ---------------------------------------------------------------------
    // X_i / Y_i: forward / backward lists (swapped by --flip, model.py:380-382)
    float* Xm = flip ? Gm : Fm;
    float* Ym = flip ? Fm : Gm;
    float* Pm = R;                               // P_j at Pm + (j-1)*MS, j = 1..T-2   (P_0 = X_0)
    float* Sm = R + (T - 2) * MS;                // S_j at Sm + (j-1)*MS               (S_0 = Y_0)
    float* scratch = R + 2 * (T - 2) * MS;       // 3 matrices
    auto Pj = [&](int j) { return j == 0 ? Xm : Pm + (j - 1) * MS; };
    auto Sj = [&](int j) { return j == 0 ? Ym : Sm + (j - 1) * MS; };
    const int grp = tid >> 8, gtid = tid & 255;  // two 256-thread groups run independent products
    for (int j = 1; j <= T - 2; ++j) {
        if (grp == 0) mm_smem<4, 4>(Pj(j), NP, Pj(j - 1), NP, 1, Xm + j * MS, NP, 1, N, false, gtid, 256);
        else          mm_smem<4, 4>(Sj(j), NP, Ym + j * MS, NP, 1, Sj(j - 1), NP, 1, N, false, gtid, 256);
        __syncthreads();
    }

    // ---- phase 3: losses and the reverse sweep -------------------------------------------------------------
    float* freeb[2 * kFusedMaxT + 4];   // grows by two buffers per level (P_j, S_j are recycled)
    int nfree = 0;
    freeb[nfree++] = scratch;
    freeb[nfree++] = scratch + MS;
    freeb[nfree++] = scratch + 2 * MS;
    float* gP = nullptr;
    float* gS = nullptr;
    const float cgrad = 1.0f / ((float)(T - 2) * (float)p.B * (float)N);
    for (int j = T - 2; j >= 0; --j) {
        float* dW = nullptr;
        if (j >= 1) {
            dW = freeb[--nfree];
            mm_smem<2, 4>(dW, NP, Pj(j), NP, 1, Sj(j), NP, 1, N, false, tid, kFusedThreads);
            __syncthreads();
            float lsum = 0.f, asum = 0.f;
            for (int n = warp; n < N; n += NW) {
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
                const float dg = __shfl_sync(kFull, w[n >> 5], n & 31) + kEpsLog;
                if (lane == 0) {
                    lsum += logf(rs) - logf(dg);
                    asum += (bi == n) ? 1.f : 0.f;
                }
                const float ir = 1.0f / rs, idg = 1.0f / dg;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = lane + 32 * h;
                    if (m < N) dW[n * NP + m] = cgrad * (ir - (m == n ? idg : 0.f));
                }
            }
            if (lane == 0) { red[warp] = lsum; red[NW + warp] = asum; }
            __syncthreads();
            if (tid == 0) {
                float l = 0.f, a = 0.f;
                for (int w2 = 0; w2 < NW; ++w2) { l += red[w2]; a += red[NW + w2]; }
                p.ws_partial[((int64_t)b * (T - 2) + (j - 1)) * 2 + 0] = l;
                p.ws_partial[((int64_t)b * (T - 2) + (j - 1)) * 2 + 1] = a;
            }
            if (!gb) { freeb[nfree++] = dW; __syncthreads(); continue; }
        }
        if (!gb) continue;
        // gP_j = dW_j S_j^T + gP_{j+1} X_{j+1}^T ;  gS_j = P_j^T dW_j + Y_{j+1}^T gS_{j+1}
        float* nP = freeb[--nfree];
        float* nS = freeb[--nfree];
        if (grp == 0) {
            if (dW) mm_smem<4, 4>(nP, NP, dW, NP, 1, Sj(j), 1, NP, N, false, gtid, 256);
            if (gP) mm_smem<4, 4>(nP, NP, gP, NP, 1, Xm + (j + 1) * MS, 1, NP, N, dW != nullptr, gtid, 256);
        } else {
            if (dW) mm_smem<4, 4>(nS, NP, Pj(j), 1, NP, dW, NP, 1, N, false, gtid, 256);
            if (gS) mm_smem<4, 4>(nS, NP, Ym + (j + 1) * MS, 1, NP, gS, NP, 1, N, dW != nullptr, gtid, 256);
        }
        __syncthreads();
        if (dW) freeb[nfree++] = dW;
        if (gP) freeb[nfree++] = gP;
        if (gS) freeb[nfree++] = gS;
        if (j >= 1) { freeb[nfree++] = Pj(j); freeb[nfree++] = Sj(j); }
        gP = nP;
        gS = nS;
        // dX_j = P_{j-1}^T gP_j ; dY_j = gS_j S_{j-1}^T   (j = 0: dX_0 = gP_0, dY_0 = gS_0)
        float* dX = gP;
        float* dY = gS;
        if (j >= 1) {
            dX = freeb[--nfree];
            dY = freeb[--nfree];
            if (grp == 0) mm_smem<4, 4>(dX, NP, Pj(j - 1), 1, NP, gP, NP, 1, N, false, gtid, 256);
            else          mm_smem<4, 4>(dY, NP, gS, NP, 1, Sj(j - 1), 1, NP, N, false, gtid, 256);
            __syncthreads();
        }
        // transition-matrix backward for pair j -> Z = dA_j
        float* Z = freeb[--nfree];
        const float* dF = flip ? dY : dX;
        const float* dG = flip ? dX : dY;
        const float* Fi = Fm + j * MS;
        const float* Gi = Gm + j * MS;
        const float* ar = araw + (int64_t)j * N * N;
        const unsigned char* gc = gcodes + (int64_t)j * N * N;
        for (int n = warp; n < N; n += NW) {
            float y[2], dy[2];
            float dot = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = lane + 32 * h;
                y[h] = dy[h] = 0.f;
                if (m < N) { y[h] = Fi[n * NP + m]; dy[h] = dF[n * NP + m]; dot += y[h] * dy[h]; }
            }
            dot = warp_sum(dot);
            const float den = s12[j * N + n];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = lane + 32 * h;
                if (m < N) {
                    float g = 0.f;
                    if (!(gc[n * N + m] & 1u)) {
                        if (softmax) g = y[h] * (dy[h] - dot) / tau;
                        else { const float E = expf(ar[n * N + m] / tau); g = (dy[h] - dot) / den * (2.0f * (E - 1.0f) * E) / tau; }
                    }
                    Z[n * NP + m] = g;
                }
            }
        }
        __syncthreads();
        for (int m = warp; m < N; m += NW) {
            float y[2], dy[2];
            float dot = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = lane + 32 * h;
                y[h] = dy[h] = 0.f;
                if (n < N) { y[h] = Gi[m * NP + n]; dy[h] = dG[m * NP + n]; dot += y[h] * dy[h]; }
            }
            dot = warp_sum(dot);
            const float den = s21[j * N + m];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = lane + 32 * h;
                if (n < N && !gc[n * N + m]) {
                    float g;
                    if (softmax) g = y[h] * (dy[h] - dot) / tau;
                    else { const float E = expf(ar[n * N + m] / tau); g = (dy[h] - dot) / den * (2.0f * (E - 1.0f) * E) / tau; }
                    Z[n * NP + m] += g;
                }
            }
        }
        __syncthreads();
        // dQ_j += Z Q_{j+1} ; dQ_{j+1} += Z^T Q_j
        dq_update(gb + (int64_t)j * D, qb + (int64_t)(j + 1) * D, gs, Z, NP, N, D, false, tid, kFusedThreads);
        dq_update(gb + (int64_t)(j + 1) * D, qb + (int64_t)j * D, gs, Z, NP, N, D, true, tid, kFusedThreads);
        __syncthreads();
        freeb[nfree++] = Z;
        if (j >= 1) { freeb[nfree++] = dX; freeb[nfree++] = dY; }
    }

    // ---- normalisation backward: df = (dq - q (q . dq)) / max(|f|, eps) --------------------------------------
    if (gb) {
        for (int row = warp; row < N * T; row += NW) {
            const int n = row / T, t = row - n * T;
            float4 qv[2], gv[2];
            float dot = 0.f;
            int c = 0;
            for (int d = lane * 4; d < D; d += 128, ++c) {
                qv[c] = *reinterpret_cast<const float4*>(qb + (int64_t)row * D + d);
                gv[c] = *reinterpret_cast<const float4*>(gb + (int64_t)row * D + d);
                dot += qv[c].x * gv[c].x + qv[c].y * gv[c].y + qv[c].z * gv[c].z + qv[c].w * gv[c].w;
            }
            dot = warp_sum(dot);
            const float in = invn[t * N + n];
            if (!(nrm[t * N + n] > kEpsNorm)) dot = 0.f;       // clamp active: q = f / eps, no projection term
            c = 0;
            for (int d = lane * 4; d < D; d += 128, ++c) {
                float4 o;
                o.x = (gv[c].x - qv[c].x * dot) * in;
                o.y = (gv[c].y - qv[c].y * dot) * in;
                o.z = (gv[c].z - qv[c].z * dot) * in;
                o.w = (gv[c].w - qv[c].w * dot) * in;
                *reinterpret_cast<float4*>(gb + (int64_t)row * D + d) = o;
            }
        }
    }

    // ---- cross-clip reduction of the per-clip sums by the last CTA to finish, in clip order (deterministic) ---
    if (tid == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(p.ws_counter, 1u);
        if (ticket == (unsigned)p.B - 1u) {
            __threadfence();
            const float inv = 1.0f / ((float)p.B * (float)N);
            for (int j = 0; j < T - 2; ++j) {
                float l = 0.f, a = 0.f;
                for (int bb = 0; bb < p.B; ++bb) {
                    l += ld_cg(p.ws_partial + ((int64_t)bb * (T - 2) + j) * 2 + 0);
                    a += ld_cg(p.ws_partial + ((int64_t)bb * (T - 2) + j) * 2 + 1);
                }
                p.xent[j] = l * inv;
                p.acc[j] = a * inv;
            }
            if (p.dev_state) p.dev_state[1] = p.offset + (uint64_t)p.pinc * 2u * (unsigned)(T - 1);
            *p.ws_counter = 0u;
        }
    }
}

int launch_walk_fused(const WalkParams& p, crw_stream_t stream) {
    const FusedLayout L = fused_layout(p.N, p.T, p.D);
    auto k = walk_fused_kernel;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes);
    CRW_LAUNCH(k, p.B, kFusedThreads, L.bytes, stream, p);
    return check_launch("walk_fused");
}

}  // namespace crw
