// walk_fused.cu - the contrastive random walk of code/model.py:366-413, forward and backward, for clips whose N x N
// transition matrices fit in shared memory (N <= 64; patch graphs: N = 49).
//
// Three launches, every matrix staged in shared memory inside each, exchanged through the L2-resident workspace:
//
//   walk_pairs_fwd   one CTA per (clip, frame pair): L2-normalise the two frames (model.py:118) while staging them,
//                    A_i = Q_i Q_{i+1}^T (model.py:68), edge dropout codes (model.py:81, in-place union mask F4,
//                    transposed draw layout F6; in-kernel Philox replay or supplied uniforms), ZeroSoftmax / softmax
//                    rows in both directions -> F_i (A12), G_i (A21)                       [B*(T-1) CTAs]
//   walk_chain       one CTA per clip: prefix / suffix chains P_j = P_{j-1} X_j, S_j = Y_j S_{j-1} (model.py:376-380
//                    re-associated, 3(T-2) products), W_j = P_j S_j, log-diagonal cross-entropy + argmax accuracy
//                    (model.py:395-397), dW_j and the reverse sweep -> dF_i, dG_i            [B CTAs]
//   walk_pairs_bwd   one CTA per (clip, pair): transition-matrix backward -> dA_i, dQ_i += dA Q_{i+1},
//                    dQ_{i+1} += dA^T Q_i; the last CTA of a clip applies the normalisation backward  [B*(T-1) CTAs]
//
// The sequential part (the chain) is the only one left on a single SM per clip; the two pair kernels spread the
// K = 128 contractions over (T-1) times more SMs.  All small products use 128-bit shared-memory operand loads.
#include "walk.cuh"

namespace crw {

constexpr int kFusedThreads = 512;
constexpr int NW = kFusedThreads / 32;

// ---- small dense products on shared-memory matrices (row stride NP, NP % 4 == 0, (NP/4) odd) ------------------------
// "Warp-row" mapping: a warp owns TMW = ceil(N / nw) consecutive output rows, a lane owns two output columns.  The A
// operand of a row is then the same address for all 32 lanes (one broadcast wavefront per 128-bit load) and the B
// operand is a contiguous, conflict-free sweep, which keeps these N ~ 49 products FMA-issue bound instead of
// shared-memory-bandwidth bound (a 4x4 register tile per thread needs ~2x the wavefronts per FMA).
// C may live in shared or global memory (row stride ldc).  Warps whose rows fall outside N simply return.
#define CRW_COMP4(v, kk) ((kk) == 0 ? (v).x : (kk) == 1 ? (v).y : (kk) == 2 ? (v).z : (v).w)

// C[r][c] (+)= sum_k A[r][k] B[k][c]
template <int TMAX>
__device__ __forceinline__ void wr_nn(float* C, int ldc, const float* A, const float* B, int N, int NP, bool accumulate,
                                      int w, int nw, int lane) {
    constexpr int TMW = TMAX;              // rows per warp; warps beyond ceil(N / TMAX) idle
    const int r0 = w * TMW;
    (void)nw;
    if (r0 >= N) return;
    const int c0 = min(2 * lane, NP - 2);
    float2 acc[TMAX];                      // the lane's two adjacent columns ride in one packed FMA
    const float* arow[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) { acc[i] = make_float2(0.f, 0.f); arow[i] = A + min(r0 + i, N - 1) * NP; }
    const float* bp = B + c0;
    const int K4 = N >> 2;
    for (int k4 = 0; k4 < K4; ++k4) {
        float4 a4[TMAX];
#pragma unroll
        for (int i = 0; i < TMAX; ++i) a4[i] = *reinterpret_cast<const float4*>(arow[i] + k4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float2 b2 = *reinterpret_cast<const float2*>(bp + (k4 * 4 + kk) * NP);
#pragma unroll
            for (int i = 0; i < TMAX; ++i) acc[i] = ffma2(CRW_COMP4(a4[i], kk), b2, acc[i]);
        }
    }
    for (int k = K4 * 4; k < N; ++k) {
        const float2 b2 = *reinterpret_cast<const float2*>(bp + k * NP);
#pragma unroll
        for (int i = 0; i < TMAX; ++i) acc[i] = ffma2(arow[i][k], b2, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        const int r = r0 + i;
        if (i >= TMW || r >= N) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = 2 * lane + j;
            if (c >= N) continue;
            float* o = C + r * ldc + c;
            const float v = j ? acc[i].y : acc[i].x;
            *o = accumulate ? (*o + v) : v;
        }
    }
}

// C[r][c] (+)= sum_k A[r][k] B[c][k]      (lane's columns: lane and lane + 32 -> neighbouring lanes read neighbouring rows of B)
template <int TMAX>
__device__ __forceinline__ void wr_nt(float* C, int ldc, const float* A, int lda, const float* B, int ldb, int K, int N,
                                      bool accumulate, int w, int nw, int lane) {
    constexpr int TMW = TMAX;              // rows per warp; warps beyond ceil(N / TMAX) idle
    const int r0 = w * TMW;
    (void)nw;
    if (r0 >= N) return;
    float2 acc[TMAX][2];                   // (even k, odd k) partial sums: both operands of the packed FMA are adjacent pairs
    const float* arow[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) { acc[i][0] = acc[i][1] = make_float2(0.f, 0.f); arow[i] = A + min(r0 + i, N - 1) * lda; }
    const float* brow[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) brow[j] = B + min(lane + 32 * j, N - 1) * ldb;
    const int K4 = K >> 2;
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
        float4 b4[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) b4[j] = *reinterpret_cast<const float4*>(brow[j] + k4 * 4);
#pragma unroll
        for (int i = 0; i < TMAX; ++i) {
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
    for (int k = K4 * 4; k < K; ++k) {
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j].x = fmaf(arow[i][k], brow[j][k], acc[i][j].x);
    }
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        const int r = r0 + i;
        if (i >= TMW || r >= N) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = lane + 32 * j;
            if (c >= N) continue;
            float* o = C + r * ldc + c;
            const float v = acc[i][j].x + acc[i][j].y;
            *o = accumulate ? (*o + v) : v;
        }
    }
}

// C[r][c] (+)= sum_k A[k][r] B[k][c]
template <int TMAX>
__device__ __forceinline__ void wr_tn(float* C, int ldc, const float* A, const float* B, int N, int NP, bool accumulate,
                                      int w, int nw, int lane) {
    constexpr int TMW = TMAX;              // rows per warp; warps beyond ceil(N / TMAX) idle
    const int r0 = w * TMW;
    (void)nw;
    if (r0 >= N) return;
    const int c0 = min(2 * lane, NP - 2);
    float2 acc[TMAX];
    int ro[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) { acc[i] = make_float2(0.f, 0.f); ro[i] = min(r0 + i, N - 1); }
    const float* bp = B + c0;
    const float* ap = A + r0;                      // r0 is a multiple of TMAX (4 or 8): 16-byte aligned row chunks,
    (void)ro;                                      // the chunk may run into the pad columns (< NP), never past the row
#pragma unroll 4
    for (int k = 0; k < N; ++k) {
        const float2 b2 = *reinterpret_cast<const float2*>(bp + k * NP);
        float av[TMAX];
#pragma unroll
        for (int i4 = 0; i4 < TMAX / 4; ++i4) {
            const float4 a4 = *reinterpret_cast<const float4*>(ap + k * NP + i4 * 4);
            av[i4 * 4 + 0] = a4.x; av[i4 * 4 + 1] = a4.y; av[i4 * 4 + 2] = a4.z; av[i4 * 4 + 3] = a4.w;
        }
#pragma unroll
        for (int i = 0; i < TMAX; ++i) acc[i] = ffma2(av[i], b2, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        const int r = r0 + i;
        if (i >= TMW || r >= N) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = 2 * lane + j;
            if (c >= N) continue;
            float* o = C + r * ldc + c;
            const float v = j ? acc[i].y : acc[i].x;
            *o = accumulate ? (*o + v) : v;
        }
    }
}

// O[r][:] = sum_k Zop(r,k) * Q[k][:]   (r < N, D <= 128*? columns: a lane owns 4 columns per 128-column chunk), Q staged in
// smem (stride DP), O global (stride D), overwrite.  transposed == false: Zop(r,k) = Z[r][k]; true: Zop(r,k) = Z[k][r]
template <int TMAX>
__device__ __forceinline__ void dq_wr(float* O, const float* Z, int NP, const float* Q, int DP, int N, int D, bool transposed,
                                      int w, int nw, int lane) {
    constexpr int TMW = TMAX;              // rows per warp; warps beyond ceil(N / TMAX) idle
    const int r0 = w * TMW;
    (void)nw;
    if (r0 >= N) return;
    for (int d0 = lane * 4; d0 < D; d0 += 128) {
        float4 acc[TMAX];
        int ro[TMAX];
#pragma unroll
        for (int i = 0; i < TMAX; ++i) { acc[i] = make_float4(0.f, 0.f, 0.f, 0.f); ro[i] = min(r0 + i, N - 1); }
        const float* qp = Q + d0;
        if (transposed) {
            const float* zp = Z + r0;                  // aligned chunk of TMAX (multiple of 4) columns of row k
#pragma unroll 4
            for (int k = 0; k < N; ++k) {
                const float4 qv = *reinterpret_cast<const float4*>(qp + k * DP);
                float zz[TMAX];
#pragma unroll
                for (int i4 = 0; i4 < TMAX / 4; ++i4) {
                    const float4 z4 = *reinterpret_cast<const float4*>(zp + k * NP + i4 * 4);
                    zz[i4 * 4 + 0] = z4.x; zz[i4 * 4 + 1] = z4.y; zz[i4 * 4 + 2] = z4.z; zz[i4 * 4 + 3] = z4.w;
                }
#pragma unroll
                for (int i = 0; i < TMAX; ++i) {
                    acc[i].x = fmaf(zz[i], qv.x, acc[i].x);
                    acc[i].y = fmaf(zz[i], qv.y, acc[i].y);
                    acc[i].z = fmaf(zz[i], qv.z, acc[i].z);
                    acc[i].w = fmaf(zz[i], qv.w, acc[i].w);
                }
            }
        } else {
            const int K4 = N >> 2;
            for (int k4 = 0; k4 < K4; ++k4) {
                float4 z4[TMAX];
#pragma unroll
                for (int i = 0; i < TMAX; ++i) z4[i] = *reinterpret_cast<const float4*>(Z + ro[i] * NP + k4 * 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 qv = *reinterpret_cast<const float4*>(qp + (k4 * 4 + kk) * DP);
#pragma unroll
                    for (int i = 0; i < TMAX; ++i) {
                        const float z = CRW_COMP4(z4[i], kk);
                        acc[i].x = fmaf(z, qv.x, acc[i].x);
                        acc[i].y = fmaf(z, qv.y, acc[i].y);
                        acc[i].z = fmaf(z, qv.z, acc[i].z);
                        acc[i].w = fmaf(z, qv.w, acc[i].w);
                    }
                }
            }
            for (int k = K4 * 4; k < N; ++k) {
                const float4 qv = *reinterpret_cast<const float4*>(qp + k * DP);
#pragma unroll
                for (int i = 0; i < TMAX; ++i) {
                    const float z = Z[ro[i] * NP + k];
                    acc[i].x = fmaf(z, qv.x, acc[i].x);
                    acc[i].y = fmaf(z, qv.y, acc[i].y);
                    acc[i].z = fmaf(z, qv.z, acc[i].z);
                    acc[i].w = fmaf(z, qv.w, acc[i].w);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TMAX; ++i) {
            const int r = r0 + i;
            if (i < TMW && r < N) *reinterpret_cast<float4*>(O + (int64_t)r * D + d0) = acc[i];
        }
    }
}

// row-stochastic pass over one row held as two values per lane (N <= 64)
__device__ __forceinline__ float stoch_row(float (&xv)[2], float (&ev)[2], int lane, int N, bool softmax) {
    float s = 0.f;
    if (softmax) {
        float mx = -INFINITY;
#pragma unroll
        for (int h = 0; h < 2; ++h) if (lane + 32 * h < N) mx = fmaxf(mx, xv[h]);
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
    return s;
}

__device__ __forceinline__ float4 ld_cg4(const float* p) {
#ifdef CRW_SIM
    return *reinterpret_cast<const float4*>(p);
#else
    return __ldcg(reinterpret_cast<const float4*>(p));
#endif
}

// ---- kernel 1: per (clip, pair) -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFusedThreads, 1) walk_pairs_fwd_kernel(WalkParams p) {
    CRW_DYN_SMEM(smem_raw);
    if (p.dev_state) { p.seed = ld_cg64(p.dev_state); p.offset = ld_cg64(p.dev_state + 1); }
    float* smem = reinterpret_cast<float*>(smem_raw);
    const int N = p.N, T = p.T, D = p.D;
    const FusedLayout L = fused_layout(N, T, D);
    const int NP = L.NP, MS = L.MS, DP = L.DP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / (T - 1), i = blockIdx.x - b * (T - 1);
    float* Qa = smem;                         // N x DP
    float* Qb = Qa + N * DP;
    float* At = Qb + N * DP;                  // N x NP raw affinity
    unsigned char* codes = reinterpret_cast<unsigned char*>(At + MS);
    const bool softmax = (p.flags & CRW_WALK_SOFTMAX) != 0;
    const float tau = p.tau;
    const int64_t gs = (int64_t)T * D;
    const float* fb = p.feats + (int64_t)b * N * gs;
    float* qb = p.q + (int64_t)b * N * gs;

    // stage + normalise frames i and i+1; the pair owns frame i (and the last pair also frame T-1).  A warp owns rows
    // warp, warp + 16, ...: all of their global loads are issued before the first use (one memory round trip, not seven)
    {
        constexpr int RP = (2 * 64 + NW - 1) / NW;          // row passes for N <= 64
        float4 v[RP][2];
#pragma unroll
        for (int rp = 0; rp < RP; ++rp) {
            const int row = warp + NW * rp;
            const int which = row >= N, n = row - which * N, t = i + which;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int d = lane * 4 + 128 * c;
                v[rp][c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < 2 * N && d < D) v[rp][c] = *reinterpret_cast<const float4*>(fb + (int64_t)n * gs + (int64_t)t * D + d);
            }
        }
#pragma unroll
        for (int rp = 0; rp < RP; ++rp) {
            const int row = warp + NW * rp;
            if (row >= 2 * N) break;
            const int which = row >= N, n = row - which * N, t = i + which;
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) ss += v[rp][c].x * v[rp][c].x + v[rp][c].y * v[rp][c].y + v[rp][c].z * v[rp][c].z + v[rp][c].w * v[rp][c].w;
            ss = warp_sum(ss);
            const float nr = sqrtf(ss), den = fmaxf(nr, kEpsNorm);
            const bool owner = !which || i == T - 2;
            float* dst = (which ? Qb : Qa) + n * DP;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int d = lane * 4 + 128 * c;
                if (d < D) {
                    float4 o = v[rp][c];
                    o.x /= den; o.y /= den; o.z /= den; o.w /= den;
                    *reinterpret_cast<float4*>(dst + d) = o;
                    if (owner) *reinterpret_cast<float4*>(qb + (int64_t)n * gs + (int64_t)t * D + d) = o;
                }
            }
            if (owner && lane == 0) {
                p.ws_invn[((int64_t)b * T + t) * N + n] = 1.0f / den;
                p.ws_nrm[((int64_t)b * T + t) * N + n] = nr;
            }
        }
    }
    __syncthreads();
    wr_nt<4>(At, NP, Qa, DP, Qb, DP, D, N, false, warp, NW, lane);
    __syncthreads();

    const int64_t pm = ((int64_t)b * (T - 1) + i) * MS;             // this pair's matrices in the workspace (stride NP)
    const int64_t pc = ((int64_t)b * (T - 1) + i) * N * N;          // ... and its codes (stride N)
    float* Fg = p.ws_F + pm;
    float* Gg = p.ws_G + pm;
    float* araw = p.ws_araw + pm;
    unsigned char* gcodes = p.ws_codes + pc;
    const int64_t numel = (int64_t)p.B * N * N;
    // forward rows (A12): row n over m; also the dropout codes (bit0 forward draw, bit1 backward draw)
    for (int n = warp; n < N; n += NW) {
        float xv[2], ev[2];
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
                gcodes[n * N + m] = (unsigned char)code;
                araw[n * NP + m] = a;
                xv[h] = ((code & 1u) ? kNegDrop : a) / tau;
            }
        }
        const float s = stoch_row(xv, ev, lane, N, softmax);
#pragma unroll
        for (int h = 0; h < 2; ++h) { const int m = lane + 32 * h; if (m < N) Fg[n * NP + m] = ev[h] / s; }
        if (lane == 0) p.ws_s12[((int64_t)b * (T - 1) + i) * N + n] = s;
    }
    __syncthreads();
    // backward rows (A21): row m of G = column m of A over n, union mask
    for (int m = warp; m < N; m += NW) {
        float xv[2], ev[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = lane + 32 * h;
            xv[h] = 0.f;
            if (n < N) xv[h] = (codes[n * N + m] ? kNegDrop : At[n * NP + m]) / tau;
        }
        const float s = stoch_row(xv, ev, lane, N, softmax);
#pragma unroll
        for (int h = 0; h < 2; ++h) { const int n = lane + 32 * h; if (n < N) Gg[m * NP + n] = ev[h] / s; }
        if (lane == 0) p.ws_s21[((int64_t)b * (T - 1) + i) * N + m] = s;
    }
}

__global__ void advance_philox_state_kernel(uint64_t* state, uint64_t inc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state[1] += inc;
}

// ---- kernel 2: per clip ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFusedThreads, 1) walk_chain_kernel(WalkParams p) {
    CRW_DYN_SMEM(smem_raw);
    float* smem = reinterpret_cast<float*>(smem_raw);
    const int N = p.N, T = p.T;
    const FusedLayout L = fused_layout(N, T, p.D);
    const int NP = L.NP, MS = L.MS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const bool flip = (p.flags & CRW_WALK_FLIP) != 0;
    float* Fm = smem;                               // (T-1) matrices
    float* Gm = Fm + (T - 1) * MS;
    float* Pm = Gm + (T - 1) * MS;                  // P_j at Pm + (j-1)*MS, j = 1..T-2 (P_0 = X_0)
    float* Sm = Pm + (T - 2) * MS;
    float* scratch = Sm + (T - 2) * MS;             // 3 matrices
    float* red = scratch + 3 * MS;                  // 2*NW floats
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * NW + 2);

    // F_i, G_i of this clip: two TMA bulk copies (identical layout in the workspace and in shared memory)
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

    float* Xm = flip ? Gm : Fm;                     // forward / backward lists (swapped by --flip, model.py:380-382)
    float* Ym = flip ? Fm : Gm;
    float* dXg = (flip ? p.ws_dG : p.ws_dF) + cm;   // gradients w.r.t. X_j / Y_j go to the F / G slots they came from
    float* dYg = (flip ? p.ws_dF : p.ws_dG) + cm;
    auto Pj = [&](int j) { return j == 0 ? Xm : Pm + (j - 1) * MS; };
    auto Sj = [&](int j) { return j == 0 ? Ym : Sm + (j - 1) * MS; };
    const int grp = warp >> 3, gw = warp & 7;       // two 8-warp groups run independent products
    for (int j = 1; j <= T - 2; ++j) {
        if (grp == 0) wr_nn<8>(Pj(j), NP, Pj(j - 1), Xm + j * MS, N, NP, false, gw, 8, lane);
        else          wr_nn<8>(Sj(j), NP, Ym + j * MS, Sj(j - 1), N, NP, false, gw, 8, lane);
        __syncthreads();
    }

    float* freeb[2 * kFusedMaxT + 4];
    int nfree = 0;
    freeb[nfree++] = scratch;
    freeb[nfree++] = scratch + MS;
    freeb[nfree++] = scratch + 2 * MS;
    float* gP = nullptr;
    float* gS = nullptr;
    const bool need_grad = p.grad != nullptr;
    const float cgrad = 1.0f / ((float)(T - 2) * (float)p.B * (float)N);
    for (int j = T - 2; j >= 0; --j) {
        float* dW = nullptr;
        if (j >= 1) {
            dW = freeb[--nfree];
            wr_nn<4>(dW, NP, Pj(j), Sj(j), N, NP, false, warp, NW, lane);
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
                const float dg = __shfl_sync(kFull, (n >> 5) ? w[1] : w[0], n & 31) + kEpsLog;
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
            if (!need_grad) { freeb[nfree++] = dW; __syncthreads(); continue; }
        }
        if (!need_grad) continue;
        // gP_j = dW_j S_j^T + gP_{j+1} X_{j+1}^T ;  gS_j = P_j^T dW_j + Y_{j+1}^T gS_{j+1}.  For j = 0 these ARE dX_0 / dY_0.
        float* nP = j == 0 ? dXg : freeb[--nfree];
        float* nS = j == 0 ? dYg : freeb[--nfree];
        if (grp == 0) {
            if (dW) wr_nt<8>(nP, NP, dW, NP, Sj(j), NP, N, N, false, gw, 8, lane);
            if (gP) wr_nt<8>(nP, NP, gP, NP, Xm + (j + 1) * MS, NP, N, N, dW != nullptr, gw, 8, lane);
        } else {
            if (dW) wr_tn<8>(nS, NP, Pj(j), dW, N, NP, false, gw, 8, lane);
            if (gS) wr_tn<8>(nS, NP, Ym + (j + 1) * MS, gS, N, NP, dW != nullptr, gw, 8, lane);
        }
        __syncthreads();
        if (j == 0) break;
        if (dW) freeb[nfree++] = dW;
        if (gP) freeb[nfree++] = gP;
        if (gS) freeb[nfree++] = gS;
        freeb[nfree++] = Pj(j);
        freeb[nfree++] = Sj(j);
        gP = nP;
        gS = nS;
        // dX_j = P_{j-1}^T gP_j ; dY_j = gS_j S_{j-1}^T  -> straight to the workspace
        if (grp == 0) wr_tn<8>(dXg + (int64_t)j * MS, NP, Pj(j - 1), gP, N, NP, false, gw, 8, lane);
        else          wr_nt<8>(dYg + (int64_t)j * MS, NP, gS, NP, Sj(j - 1), NP, N, N, false, gw, 8, lane);
        __syncthreads();
    }

    // cross-clip reduction of the per-clip sums by the last CTA to finish, in clip order (deterministic);
    // xent[T-2] receives the loss itself, sum_j xent_j / (T-2) (model.py:413)
    unsigned* s_flag = reinterpret_cast<unsigned*>(red + 2 * NW);
    if (tid == 0) {
        __threadfence();
        *s_flag = atomicAdd(p.ws_counter, 1u) == (unsigned)p.B - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (*s_flag && warp == 0) {
        __threadfence();
        const float inv = 1.0f / ((float)p.B * (float)N);
        float tot = 0.f;
        for (int j = 0; j < T - 2; ++j) {
            float l = 0.f, a = 0.f;
            for (int b0 = 0; b0 < p.B; b0 += 32) {                     // fixed order: chunks of 32 clips, butterfly inside
                const int bb = b0 + lane;
                float lv = 0.f, av = 0.f;
                if (bb < p.B) {
                    lv = ld_cg(p.ws_partial + ((int64_t)bb * (T - 2) + j) * 2 + 0);
                    av = ld_cg(p.ws_partial + ((int64_t)bb * (T - 2) + j) * 2 + 1);
                }
                l += warp_sum(lv);
                a += warp_sum(av);
            }
            if (lane == 0) { p.xent[j] = l * inv; p.acc[j] = a * inv; }
            tot += l * inv;
        }
        if (lane == 0) {
            p.xent[T - 2] = tot / (float)(T - 2);
            // every pair CTA of the previous launch has consumed the device-resident Philox state: advance it
            if (p.dev_state && p.rate > 0.f && !p.u12)
                p.dev_state[1] = ld_cg64(p.dev_state + 1) + (uint64_t)p.pinc * 2u * (unsigned)(T - 1);
            *p.ws_counter = 0u;
        }
    }
}

// ---- kernel 3: per (clip, pair) ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFusedThreads, 1) walk_pairs_bwd_kernel(WalkParams p) {
    CRW_DYN_SMEM(smem_raw);
    float* smem = reinterpret_cast<float*>(smem_raw);
    const int N = p.N, T = p.T, D = p.D;
    const FusedLayout L = fused_layout(N, T, D);
    const int NP = L.NP, MS = L.MS, DP = L.DP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / (T - 1), i = blockIdx.x - b * (T - 1);
    float* Qa = smem;
    float* Qb = Qa + N * DP;
    float* Z = Qb + N * DP;                   // N x NP: dA
    float* Ar = Z + MS;                       // N x NP: raw affinity
    float* Fs = Ar + MS;                      // F, dF, G, dG of the pair: staged with the other inputs, one memory round trip
    float* dFs = Fs + MS;
    float* Gs = dFs + MS;
    float* dGs = Gs + MS;
    unsigned char* codes = reinterpret_cast<unsigned char*>(dGs + MS);
    unsigned* s_last = reinterpret_cast<unsigned*>(codes + ((N * N + 15) & ~15));   // 2 flags
    const bool softmax = (p.flags & CRW_WALK_SOFTMAX) != 0;
    const float tau = p.tau;
    const int64_t gs = (int64_t)T * D;
    const float* qb = p.q + (int64_t)b * N * gs;
    const int64_t pm = ((int64_t)b * (T - 1) + i) * MS;
    const int64_t pc = ((int64_t)b * (T - 1) + i) * N * N;

    for (int v = tid; v < 2 * N * (D / 4); v += kFusedThreads) {
        const int which = v >= N * (D / 4), r = v - which * N * (D / 4);
        const int n = r / (D / 4), d4 = r - n * (D / 4);
        *reinterpret_cast<float4*>((which ? Qb : Qa) + n * DP + d4 * 4) =
            ld_cg4(qb + (int64_t)n * gs + (int64_t)(i + which) * D + d4 * 4);
    }
    for (int e = tid; e < MS / 4; e += kFusedThreads) {
        *reinterpret_cast<float4*>(Ar + e * 4) = ld_cg4(p.ws_araw + pm + e * 4);
        *reinterpret_cast<float4*>(Fs + e * 4) = ld_cg4(p.ws_F + pm + e * 4);
        *reinterpret_cast<float4*>(dFs + e * 4) = ld_cg4(p.ws_dF + pm + e * 4);
        *reinterpret_cast<float4*>(Gs + e * 4) = ld_cg4(p.ws_G + pm + e * 4);
        *reinterpret_cast<float4*>(dGs + e * 4) = ld_cg4(p.ws_dG + pm + e * 4);
    }
    for (int e = tid; e < N * N; e += kFusedThreads) codes[e] = p.ws_codes[pc + e];
    __syncthreads();
    // rows of F: Z[n][m] = d loss / d A[n][m] through the forward matrix
    for (int n = warp; n < N; n += NW) {
        float y[2], dy[2];
        float dot = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = lane + 32 * h;
            y[h] = dy[h] = 0.f;
            if (m < N) { y[h] = Fs[n * NP + m]; dy[h] = dFs[n * NP + m]; dot += y[h] * dy[h]; }
        }
        dot = warp_sum(dot);
        const float den = ld_cg(p.ws_s12 + ((int64_t)b * (T - 1) + i) * N + n);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = lane + 32 * h;
            if (m < N) {
                float g = 0.f;
                if (!(codes[n * N + m] & 1u)) {
                    if (softmax) g = y[h] * (dy[h] - dot) / tau;
                    else { const float E = expf(Ar[n * NP + m] / tau); g = (dy[h] - dot) / den * (2.0f * (E - 1.0f) * E) / tau; }
                }
                Z[n * NP + m] = g;
            }
        }
    }
    __syncthreads();
    // rows of G (columns of A): Z[n][m] += contribution through the backward matrix
    for (int m = warp; m < N; m += NW) {
        float y[2], dy[2];
        float dot = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = lane + 32 * h;
            y[h] = dy[h] = 0.f;
            if (n < N) { y[h] = Gs[m * NP + n]; dy[h] = dGs[m * NP + n]; dot += y[h] * dy[h]; }
        }
        dot = warp_sum(dot);
        const float den = ld_cg(p.ws_s21 + ((int64_t)b * (T - 1) + i) * N + m);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = lane + 32 * h;
            if (n < N && !codes[n * N + m]) {
                float g;
                if (softmax) g = y[h] * (dy[h] - dot) / tau;
                else { const float E = expf(Ar[n * NP + m] / tau); g = (dy[h] - dot) / den * (2.0f * (E - 1.0f) * E) / tau; }
                Z[n * NP + m] += g;
            }
        }
    }
    __syncthreads();
    // dQ_i (from this pair) = Z Q_{i+1} ; dQ_{i+1} (from this pair) = Z^T Q_i
    float* dqa = p.ws_dqa + ((int64_t)b * (T - 1) + i) * N * D;
    float* dqb = p.ws_dqb + ((int64_t)b * (T - 1) + i) * N * D;
    dq_wr<4>(dqa, Z, NP, Qb, DP, N, D, false, warp, NW, lane);
    dq_wr<4>(dqb, Z, NP, Qa, DP, N, D, true, warp, NW, lane);

    // Frame t receives dqa from pair t and dqb from pair t-1.  Whichever pair CTA delivers the last contribution of a
    // frame applies the normalisation backward to it: df = (dq - q (q.dq)) / max(|f|, eps)   (model.py:118)
    __threadfence();
    __syncthreads();
    if (tid < 2) {
        const int t = i + tid;
        const unsigned need = (t < T - 1 ? 1u : 0u) + (t > 0 ? 1u : 0u);
        const unsigned got = atomicAdd(p.ws_clipcnt + (int64_t)b * T + t, 1u) + 1u;
        s_last[tid] = got == need ? 1u : 0u;
        if (got == need) p.ws_clipcnt[(int64_t)b * T + t] = 0u;
    }
    __syncthreads();
    if (!s_last[0] && !s_last[1]) return;
    __threadfence();
    float* gb = p.grad + (int64_t)b * N * gs;
    // a warp owns rows warp, warp + 16, ...; two rows at a time, all their loads in flight before the first use
    for (int row0 = warp; row0 < 2 * N; row0 += 2 * NW) {
        float4 qv[2][2], ga[2][2], gbv[2][2];
        float in[2], nrm[2];
        bool live[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int row = row0 + NW * u;
            const int which = row >= N, n = row - which * N, t = i + which;
            live[u] = row < 2 * N && s_last[which] != 0;
            in[u] = nrm[u] = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int d = lane * 4 + 128 * c;
                qv[u][c] = ga[u][c] = gbv[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[u] && d < D) {
                    qv[u][c] = ld_cg4(qb + (int64_t)n * gs + (int64_t)t * D + d);
                    if (t < T - 1) ga[u][c] = ld_cg4(p.ws_dqa + (((int64_t)b * (T - 1) + t) * N + n) * D + d);
                    if (t > 0) gbv[u][c] = ld_cg4(p.ws_dqb + (((int64_t)b * (T - 1) + t - 1) * N + n) * D + d);
                }
            }
            if (live[u]) {
                in[u] = ld_cg(p.ws_invn + ((int64_t)b * T + t) * N + n);
                nrm[u] = ld_cg(p.ws_nrm + ((int64_t)b * T + t) * N + n);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!live[u]) continue;                         // warp-uniform
            const int row = row0 + NW * u;
            const int which = row >= N, n = row - which * N, t = i + which;
            const int64_t ro = (int64_t)n * gs + (int64_t)t * D;
            float4 gv[2];
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                gv[c] = make_float4(ga[u][c].x + gbv[u][c].x, ga[u][c].y + gbv[u][c].y, ga[u][c].z + gbv[u][c].z, ga[u][c].w + gbv[u][c].w);
                dot += qv[u][c].x * gv[c].x + qv[u][c].y * gv[c].y + qv[u][c].z * gv[c].z + qv[u][c].w * gv[c].w;
            }
            dot = warp_sum(dot);
            if (!(nrm[u] > kEpsNorm)) dot = 0.f;            // clamp active: q = f / eps
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int d = lane * 4 + 128 * c;
                if (d < D) {
                    float4 o;
                    o.x = (gv[c].x - qv[u][c].x * dot) * in[u];
                    o.y = (gv[c].y - qv[u][c].y * dot) * in[u];
                    o.z = (gv[c].z - qv[u][c].z * dot) * in[u];
                    o.w = (gv[c].w - qv[u][c].w * dot) * in[u];
                    *reinterpret_cast<float4*>(gb + ro + d) = o;
                }
            }
        }
    }
}

int launch_l2norm_rows(const float* f, float* q, float* invn, float* nrm, int64_t rows, int D, crw_stream_t stream);

int launch_walk_fused(const WalkParams& p, crw_stream_t stream) {
    const FusedLayout L = fused_layout(p.N, p.T, p.D);
    const int T = p.T;
    if (T < 2)   // a single frame: only the normalisation exists
        return launch_l2norm_rows(p.feats, p.q, p.ws_invn, p.ws_nrm, (int64_t)p.B * p.N, p.D, stream);
    auto k1 = walk_pairs_fwd_kernel;
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.pair_bytes);
    CRW_LAUNCH(k1, p.B * (T - 1), kFusedThreads, L.pair_bytes, stream, p);
    int e = check_launch("walk_pairs_fwd");
    if (e != CRW_OK) return e;
    if (T < 3) {
        if (p.dev_state && p.rate > 0.f && !p.u12) {     // no chain kernel to do it: advance the device-resident Philox state here
            CRW_LAUNCH(advance_philox_state_kernel, 1, 32, 0, stream, p.dev_state, (uint64_t)p.pinc * 2u * (unsigned)(T - 1));
            e = check_launch("advance_philox_state");
        }
        return e;
    }
    // a clip on 4 SMs only pays while the clusters of the whole batch are resident at once (B = 64: two waves of clusters
    // measured 8 % slower than one CTA per clip)
    static thread_local int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0) sm_count = 148;
    }
    if (p.grad && !(p.flags & CRW_WALK_NO_CLUSTER) && chain_cluster_fits(p.N, T) && p.B * kChainCluster <= sm_count) {
        e = launch_walk_chain_cluster(p, L.chain_bytes, stream);          // one clip = a cluster of 4 CTAs (walk_chain_cluster.cu)
    } else {
        auto k2 = walk_chain_kernel;
        cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.chain_bytes);
        CRW_LAUNCH(k2, p.B, kFusedThreads, L.chain_bytes, stream, p);
        e = check_launch("walk_chain");
    }
    if (e != CRW_OK || !p.grad) return e;
    auto k3 = walk_pairs_bwd_kernel;
    cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.pairb_bytes);
    CRW_LAUNCH(k3, p.B * (T - 1), kFusedThreads, L.pairb_bytes, stream, p);
    return check_launch("walk_pairs_bwd");
}

}  // namespace crw
