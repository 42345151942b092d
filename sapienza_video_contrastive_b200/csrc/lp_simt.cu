// lp_simt.cu - label-propagation evaluator (reference: code/utils/test_utils.py:148-179
// mem_efficient_batched_affinity, the radius mask of code/utils/__init__.py:377-391 + code/test.py:118-122, and
// the gather loop code/test.py:141-160), exact-fp32 SIMT implementation.
//
// lp_topk: one CTA per (target frame, 8x8 query tile).  The 64 query vectors stay in shared memory for the whole
// CTA; key vectors are streamed tile by tile (64 keys x 16 channels per step): all hw keys of each long-memory
// slot, and for each radius-restricted slot only the keys of the tile's (8+2R)^2 window - the out-of-radius 94 %
// of the dense affinity the reference computes and then masks with -1e10 is never formed.  Each 64x64 score tile
// goes through shared memory to a streaming per-query top-k kept in registers (4 partial lists per query,
// merged at the end), followed by the softmax over the k winners.  The affinity matrix never reaches HBM.
#include "common.cuh"
#include "lp_tc.cuh"

namespace crw {

constexpr int LQ = 64;          // queries per CTA (8 x 8 block)
constexpr int LK = 64;          // keys per step
constexpr int LC = 16;          // channels per step
constexpr int LPAD = 4;

struct LpArgs {
    const float* feats;         // (Nf, hw, C)
    const int64_t* key_frames;  // (Nt, S)
    const int64_t* query_frames;// (Nt)
    int Nt, S, n_long, h, w, C, k, R;
    float r2, inv_tau_unused, tau;
    int restricted;
    const float* dense_mask;    // (hw, hw) additive, or null
    float* Ws;                  // (Nt, k, hw)
    int64_t* Is;
};

template <int KCAP>
struct TopList {
    float v[KCAP];
    int idx[KCAP];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < KCAP; ++i) { v[i] = -INFINITY; idx[i] = 0x7fffffff; }
    }
    // keeps the list sorted by (value desc, index asc); candidates of one thread arrive with ascending index
    __device__ __forceinline__ void push(float x, int id) {
        if (!(x > v[KCAP - 1])) return;
#pragma unroll
        for (int i = 0; i < KCAP; ++i) {
            const bool better = x > v[i] || (x == v[i] && id < idx[i]);
            const float tv = better ? v[i] : x;
            const int ti = better ? idx[i] : id;
            v[i] = better ? x : v[i];
            idx[i] = better ? id : idx[i];
            x = tv;
            id = ti;
        }
    }
};

template <int KCAP>
__global__ void __launch_bounds__(256, 2) lp_topk_kernel(LpArgs a) {
    CRW_DYN_SMEM(smem_raw);
    float* Qs = reinterpret_cast<float*>(smem_raw);                 // [C][LQ + LPAD]
    const int C = a.C, hw = a.h * a.w;
    float* Ks = Qs + (size_t)C * (LQ + LPAD);                       // [LC][LK + LPAD]
    float* Ss = Ks + LC * (LK + LPAD);                              // [LQ][LK + 1]
    int* kpos = reinterpret_cast<int*>(Ss + LQ * (LK + 1));         // [LK] key position (or -1)
    int* kyx = kpos + LK;                                           // [LK] ky << 16 | kx
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n = blockIdx.y;
    const int tiles_x = (a.w + 7) / 8;
    const int qy0 = (blockIdx.x / tiles_x) * 8, qx0 = (blockIdx.x % tiles_x) * 8;
    const int64_t qframe = a.query_frames[n];

    // stage the query tile, channel-major
    for (int e = tid; e < LQ * (C / 4); e += 256) {
        const int q = e / (C / 4), c4 = e - q * (C / 4);
        const int qy = qy0 + (q >> 3), qx = qx0 + (q & 7);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qy < a.h && qx < a.w) v = __ldg(reinterpret_cast<const float4*>(a.feats + (qframe * hw + qy * a.w + qx) * C) + c4);
        Qs[(c4 * 4 + 0) * (LQ + LPAD) + q] = v.x;
        Qs[(c4 * 4 + 1) * (LQ + LPAD) + q] = v.y;
        Qs[(c4 * 4 + 2) * (LQ + LPAD) + q] = v.z;
        Qs[(c4 * 4 + 3) * (LQ + LPAD) + q] = v.w;
    }
    // epilogue ownership: thread (eq, part) scans keys part, part+4, ... of query eq
    const int eq = tid >> 2, part = tid & 3;
    const int eqy = qy0 + (eq >> 3), eqx = qx0 + (eq & 7);
    TopList<KCAP> top;
    top.init();
    __syncthreads();

    for (int slot = 0; slot < a.S; ++slot) {
        const int64_t kframe = a.key_frames[(int64_t)n * a.S + slot];
        const bool restricted = a.restricted && slot >= a.n_long;
        const bool dense = a.dense_mask != nullptr && slot >= a.n_long;
        int wy0 = 0, wx0 = 0, wh = a.h, ww = a.w;
        if (restricted) {
            wy0 = max(qy0 - a.R, 0);
            wx0 = max(qx0 - a.R, 0);
            wh = min(qy0 + 7 + a.R, a.h - 1) - wy0 + 1;
            ww = min(qx0 + 7 + a.R, a.w - 1) - wx0 + 1;
        }
        const int nkeys = wh * ww;
        const float* kbase = a.feats + kframe * (int64_t)hw * C;
        for (int k0 = 0; k0 < nkeys; k0 += LK) {
            if (tid < LK) {
                const int j = k0 + tid;
                int pos = -1, yx = 0;
                if (j < nkeys) {
                    const int wy = j / ww, wx = j - wy * ww;
                    pos = (wy0 + wy) * a.w + wx0 + wx;
                    yx = ((wy0 + wy) << 16) | (wx0 + wx);
                }
                kpos[tid] = pos;
                kyx[tid] = yx;
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            const int lkey = tid >> 2, lc4 = tid & 3;       // loader: key lkey, channels lc4*4..+3 of the chunk
            const int lpos = kpos[lkey];
            for (int c0 = 0; c0 < C; c0 += LC) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lpos >= 0 && c0 + lc4 * 4 < C) v = __ldg(reinterpret_cast<const float4*>(kbase + (int64_t)lpos * C + c0) + lc4);
                __syncthreads();                            // previous chunk fully consumed
                Ks[(lc4 * 4 + 0) * (LK + LPAD) + lkey] = v.x;
                Ks[(lc4 * 4 + 1) * (LK + LPAD) + lkey] = v.y;
                Ks[(lc4 * 4 + 2) * (LK + LPAD) + lkey] = v.z;
                Ks[(lc4 * 4 + 3) * (LK + LPAD) + lkey] = v.w;
                __syncthreads();
                const int cmax = min(LC, C - c0);
#pragma unroll 4
                for (int c = 0; c < cmax; ++c) {
                    const float4 qv = *reinterpret_cast<const float4*>(Qs + (size_t)(c0 + c) * (LQ + LPAD) + ty * 4);
                    const float4 kv = *reinterpret_cast<const float4*>(Ks + c * (LK + LPAD) + tx * 4);
                    const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
                    const float ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qa[i], ka[j], acc[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) Ss[(ty * 4 + i) * (LK + 1) + tx * 4 + j] = acc[i][j];
            __syncthreads();
            if (eqy < a.h && eqx < a.w) {
#pragma unroll 4
                for (int jj = 0; jj < LK / 4; ++jj) {
                    const int key = part + 4 * jj;
                    const int pos = kpos[key];
                    if (pos < 0) continue;
                    if (restricted) {
                        const int yx = kyx[key];
                        const int dy = (yx >> 16) - eqy, dx = (yx & 0xffff) - eqx;
                        if (!((float)(dy * dy + dx * dx) < a.r2)) continue;
                    }
                    float sc = Ss[eq * (LK + 1) + key];
                    if (dense) sc += __ldg(a.dense_mask + (int64_t)pos * hw + eqy * a.w + eqx);
                    top.push(sc / a.tau, slot * hw + pos);
                }
            }
            __syncthreads();
        }
    }

    // merge the 4 partial lists of each query (reuse the query staging area)
    float* mv = Qs;                                                  // [LQ][4][KCAP]
    int* mi = reinterpret_cast<int*>(Qs + LQ * 4 * KCAP);
#pragma unroll
    for (int i = 0; i < KCAP; ++i) { mv[(eq * 4 + part) * KCAP + i] = top.v[i]; mi[(eq * 4 + part) * KCAP + i] = top.idx[i]; }
    __syncthreads();
    if (tid < LQ) {
        const int q = tid;
        const int qy = qy0 + (q >> 3), qx = qx0 + (q & 7);
        if (qy < a.h && qx < a.w) {
            int head[4] = {0, 0, 0, 0};
            float vals[KCAP];
            float mxv = 0.f;
            const int64_t obase = (int64_t)n * a.k * hw + qy * a.w + qx;
            for (int r = 0; r < a.k; ++r) {
                float bv = -INFINITY;
                int bi = 0x7fffffff, bp = 0;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    if (head[pp] >= KCAP) continue;
                    const float v = mv[(q * 4 + pp) * KCAP + head[pp]];
                    const int id = mi[(q * 4 + pp) * KCAP + head[pp]];
                    if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; bp = pp; }
                }
                head[bp]++;
                if (r == 0) mxv = bv;
                vals[r < KCAP ? r : KCAP - 1] = bv;
                a.Is[obase + (int64_t)r * hw] = bi == 0x7fffffff ? 0 : bi;
            }
            float den = 0.f;
            for (int r = 0; r < a.k; ++r) { vals[r] = expf(vals[r] - mxv); den += vals[r]; }
            for (int r = 0; r < a.k; ++r) a.Ws[obase + (int64_t)r * hw] = vals[r] / den;
        }
    }
}

// ---- layout change + optional L2 normalisation: (C, Nf, hw) -> (Nf, hw, C) ------------------------------------------
__global__ void __launch_bounds__(256) lp_prepare_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int Nf,
                                                         int hw, int normalize) {
    CRW_DYN_SMEM(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw);               // [C][33]
    __shared__ float inv[32];
    const int f = blockIdx.y, p0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = warp; c < C; c += 8) {
        const int p = p0 + lane;
        tile[c * 33 + lane] = p < hw ? __ldg(src + ((int64_t)c * Nf + f) * hw + p) : 0.f;
    }
    __syncthreads();
    for (int pl = warp; pl < 32; pl += 8) {
        float ss = 0.f;
        for (int c = lane; c < C; c += 32) { const float v = tile[c * 33 + pl]; ss += v * v; }
        ss = warp_sum(ss);
        if (lane == 0) inv[pl] = normalize ? fmaxf(sqrtf(ss), kEpsNorm) : 1.0f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * C; e += 256) {
        const int pl = e / C, c = e - pl * C;
        if (p0 + pl < hw) dst[((int64_t)f * hw + p0 + pl) * C + c] = tile[c * 33 + pl] / inv[pl];
    }
}

// ---- one propagation step (test.py:147-157) ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) lp_gather_kernel(float* lbls, const int64_t* __restrict__ key_frames_n,
                                                        const float* __restrict__ Ws, const int64_t* __restrict__ Is, int hw,
                                                        int L, int k, int64_t out_frame) {
    const int64_t total = (int64_t)hw * L;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(e / L), l = (int)(e - (int64_t)q * L);
        float s = 0.f;
        for (int r = 0; r < k; ++r) {
            const int64_t id = Is[(int64_t)r * hw + q];
            const int64_t slot = id / hw, pos = id - slot * hw;
            const float v = lbls[(key_frames_n[slot] * hw + pos) * L + l];
            s += v * Ws[(int64_t)r * hw + q];
        }
        lbls[(out_frame * hw + q) * L + l] = s;
    }
}

template <int KCAP>
static int launch_lp(const LpArgs& a, crw_stream_t stream) {
    const size_t qbytes = (size_t)a.C * (LQ + LPAD) * 4;
    const size_t mbytes = (size_t)LQ * 4 * KCAP * 8;
    const size_t smem = (qbytes > mbytes ? qbytes : mbytes) + (size_t)LC * (LK + LPAD) * 4 + (size_t)LQ * (LK + 1) * 4 + 2 * LK * 4;
    if (smem > 227 * 1024) { set_error("lp_topk: C=%d too large for shared memory", a.C); return CRW_ERR_UNSUPPORTED; }
    auto k = lp_topk_kernel<KCAP>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(((a.w + 7) / 8) * ((a.h + 7) / 8), a.Nt);
    CRW_LAUNCH(k, grid, 256, smem, stream, a);
    return check_launch("lp_topk");
}

}  // namespace crw

using namespace crw;

static int radius_R(float radius) {          // largest integer offset R with R*R < radius^2
    int R = 0;
    while ((float)((R + 1) * (R + 1)) < radius * radius) ++R;
    return R;
}

extern "C" int crw_lp_topk_uses_tensor_cores(int C, int k, float radius, int has_dense_mask) {
    return lp_tc_supported(C, k, radius, radius > 0.f ? radius_R(radius) : 0, has_dense_mask != 0) ? 1 : 0;
}

extern "C" size_t crw_lp_topk_workspace_bytes(int Nf, int Nt, int S, int h, int w, int C, int k) {
    (void)S;
    size_t b = 256;                          // word 0: device-side error flag
    if (lp_tc_supported(C, k, 1.f, 0, false)) b = lp_tc_workspace_bytes(Nf, Nt, h, w, C);
    return b;
}

extern "C" int crw_lp_topk(const float* feats, int Nf, const int64_t* key_frames, const int64_t* query_frames, int Nt, int S,
                           int n_long, int h, int w, int C, float radius, const float* dense_mask, float temperature, int k,
                           unsigned flags, float* Ws, int64_t* Is, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    if (Nt < 0 || S <= 0 || h <= 0 || w <= 0 || C <= 0 || k <= 0 || n_long < 0 || n_long > S || !(temperature > 0.f)) {
        set_error("lp_topk: bad arguments"); return CRW_ERR_SHAPE;
    }
    if (C % 4 != 0 || (((uintptr_t)feats) & 15) != 0) { set_error("lp_topk: C must be a multiple of 4 and feats 16-byte aligned"); return CRW_ERR_UNSUPPORTED; }
    if (k > 32) { set_error("lp_topk: k <= 32 supported"); return CRW_ERR_UNSUPPORTED; }
    if ((int64_t)S * h * w >= 0x7fffffff || h >= 32768 || w >= 32768) { set_error("lp_topk: index range"); return CRW_ERR_UNSUPPORTED; }
    if (Nt == 0) return CRW_OK;
    if (!workspace || workspace_bytes < 256) { set_error("lp_topk: workspace missing"); return CRW_ERR_SHAPE; }
    const int Rr = radius > 0.f ? radius_R(radius) : 0;
    if (!(flags & CRW_LP_FORCE_SIMT) && lp_tc_supported(C, k, radius, Rr, dense_mask != nullptr) &&
        workspace_bytes >= lp_tc_workspace_bytes(Nf, Nt, h, w, C)) {
        LpTcArgs t{};
        t.key_frames = key_frames; t.query_frames = query_frames;
        t.Nt = Nt; t.S = S; t.n_long = n_long; t.h = h; t.w = w; t.C = C; t.k = k; t.R = Rr;
        t.restricted = radius > 0.f ? 1 : 0;
        // d2 < radius^2 for integer d2  <=>  d2 <= r2i
        int r2i = 0;
        while ((float)(r2i + 1) < radius * radius) ++r2i;
        t.r2i = radius > 0.f ? r2i : 0;
        t.tau = temperature; t.Ws = Ws; t.Is = Is; t.flags = flags;
        return launch_lp_tc(feats, Nf, t, workspace, workspace_bytes, stream);
    }
    cudaMemsetAsync(workspace, 0, 4, (cudaStream_t)stream);
    LpArgs a{};
    a.feats = feats; a.key_frames = key_frames; a.query_frames = query_frames;
    a.Nt = Nt; a.S = S; a.n_long = n_long; a.h = h; a.w = w; a.C = C; a.k = k;
    a.restricted = (radius > 0.f && !dense_mask) ? 1 : 0;
    a.dense_mask = dense_mask;
    a.r2 = radius * radius;
    a.R = Rr;
    a.tau = temperature;
    a.Ws = Ws; a.Is = Is;
    if (k <= 4) return launch_lp<4>(a, stream);
    if (k <= 8) return launch_lp<8>(a, stream);
    if (k <= 16) return launch_lp<16>(a, stream);
    return launch_lp<32>(a, stream);
}

extern "C" int crw_lp_prepare(const float* feats_cf, int C, int Nf, int hw, int normalize, float* feats_cl, crw_stream_t stream) {
    if (C <= 0 || Nf < 0 || hw <= 0) { set_error("lp_prepare: bad shape"); return CRW_ERR_SHAPE; }
    if ((size_t)C * 33 * 4 > 200 * 1024) { set_error("lp_prepare: C too large"); return CRW_ERR_UNSUPPORTED; }
    if (Nf == 0) return CRW_OK;
    const size_t smem = (size_t)C * 33 * 4;
    auto k = lp_prepare_kernel;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((hw + 31) / 32, Nf);
    CRW_LAUNCH(k, grid, 256, smem, stream, feats_cf, feats_cl, C, Nf, hw, normalize);
    return check_launch("lp_prepare");
}

// The whole propagation loop of a video (test.py:145-157) behind one call: the recurrence over target frames is sequential, one
// frame is tiny (hw x L outputs, k gathers each) but wants the whole GPU for a few microseconds, so it stays one launch per frame,
// issued back to back from here (36 launches cost ~0.35 ms of device time; from Python they cost more host time than that).
// (Measured alternative: the recurrence inside ONE 8-CTA cluster with a cluster barrier per frame - 36 us per frame on 8 SMs
// against 9.6 us per launch on the whole GPU.)
extern "C" int crw_lp_gather_all(float* lbls, const int64_t* key_frames, const float* Ws, const int64_t* Is, int Nt, int S, int hw, int L,
                                 int k, int first_target, int64_t out_frame0, crw_stream_t stream) {
    if (Nt < 0 || S <= 0 || hw <= 0 || L <= 0 || k <= 0 || first_target < 0 || out_frame0 < 0) { set_error("lp_gather_all: bad arguments"); return CRW_ERR_SHAPE; }
    for (int t = first_target; t < Nt; ++t) {
        int e = crw_lp_gather(lbls, key_frames + (int64_t)t * S, Ws + (int64_t)t * k * hw, Is + (int64_t)t * k * hw, hw, L, k, out_frame0 + t, stream);
        if (e != CRW_OK) return e;
    }
    return CRW_OK;
}

extern "C" int crw_lp_gather(float* lbls, const int64_t* key_frames_n, const float* Ws_n, const int64_t* Is_n,
                             int hw, int L, int k, int64_t out_frame, crw_stream_t stream) {
    if (hw <= 0 || L <= 0 || k <= 0 || out_frame < 0) { set_error("lp_gather: bad arguments"); return CRW_ERR_SHAPE; }
    const int64_t total = (int64_t)hw * L;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    CRW_LAUNCH(lp_gather_kernel, grid, 256, 0, stream, lbls, key_frames_n, Ws_n, Is_n, hw, L, k, out_frame);
    return check_launch("lp_gather");
}
