// lp_post.cu - label-map post-processing of the evaluator (reference: code/utils/test_utils.py:85-123 dump_predictions,
// code/test.py:162-164 --norm_mask): the propagated soft label maps (h, w, L) of every target frame are upsampled to the
// image size with OpenCV's bilinear resize (cv2.resize default, the float path of resizeGeneric), the hard label is the
// arg-max over L (first maximum wins, numpy.argmax) and is mapped through the colour table lbl_set (L, 3).
//
// One thread per output pixel and frame: the four source neighbours x L channels are read straight from the low-resolution
// map (L contiguous, the whole map of a frame sits in L1/L2), interpolated horizontally then vertically with the same
// float weights OpenCV computes ((dx + 0.5) * scale - 0.5 in double, rounded to float, clamped at the borders), and only
// the class index (1 byte) and the palette colour (3 bytes) are written: the (H, W, L) float tensor is never materialised.
// HBM: 4 bytes written per output pixel against 4 L read + 4 L written + 4 read by the reference's resize + argmax.
#include "common.cuh"

namespace crw {

struct LpPostArgs {
    const float* pred;          // (n, h, w, L)
    const unsigned char* pal;   // (L, 3) or null
    unsigned char* cls;         // (n, H, W) or null
    unsigned char* rgb;         // (n, H, W, 3) or null
    int n, h, w, L, H, W, norm_mask;
    double scale_x, scale_y;    // source / destination extent
};

// OpenCV's source coordinate and weight for destination index d (imgproc/resize.cpp, INTER_LINEAR, float images)
__device__ __forceinline__ void cv_linear_coord(int d, double scale, int ssize, int& s0, int& s1, float& w0, float& w1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    s0 = s;
    s1 = s + 1 < ssize ? s + 1 : ssize - 1;
    w0 = 1.f - f;
    w1 = f;
}

// arg-max class of output pixel e (flat index over (frame, Y, X))
__device__ __forceinline__ int lp_post_pixel(const LpPostArgs& a, int64_t e) {
    const int X = (int)(e % a.W);
    const int64_t r = e / a.W;
    const int Y = (int)(r % a.H), f = (int)(r / a.H);
    int x0, x1, y0, y1;
    float ax0, ax1, by0, by1;
    cv_linear_coord(X, a.scale_x, a.w, x0, x1, ax0, ax1);
    cv_linear_coord(Y, a.scale_y, a.h, y0, y1, by0, by1);
    const float* base = a.pred + (int64_t)f * a.h * a.w * a.L;
    const float* p00 = base + ((int64_t)y0 * a.w + x0) * a.L;
    const float* p01 = base + ((int64_t)y0 * a.w + x1) * a.L;
    const float* p10 = base + ((int64_t)y1 * a.w + x0) * a.L;
    const float* p11 = base + ((int64_t)y1 * a.w + x1) * a.L;
    // --norm_mask (test.py:162-164): per source pixel, subtract the min over L, then divide by the (new) max
    float mn[4] = {0.f, 0.f, 0.f, 0.f}, mx[4] = {1.f, 1.f, 1.f, 1.f};
    if (a.norm_mask) {
        const float* ps[4] = {p00, p01, p10, p11};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float lo = ps[q][0], hi = ps[q][0];
            for (int l = 1; l < a.L; ++l) { lo = fminf(lo, ps[q][l]); hi = fmaxf(hi, ps[q][l]); }
            mn[q] = lo;
            mx[q] = hi - lo;
        }
    }
    float best = -INFINITY;
    int bi = 0;
    for (int l = 0; l < a.L; ++l) {
        float v00 = p00[l], v01 = p01[l], v10 = p10[l], v11 = p11[l];
        if (a.norm_mask) {
            v00 = (v00 - mn[0]) / mx[0]; v01 = (v01 - mn[1]) / mx[1];
            v10 = (v10 - mn[2]) / mx[2]; v11 = (v11 - mn[3]) / mx[3];
        }
        // horizontal pass, then vertical pass (separate multiplies and adds, as the scalar OpenCV code)
        const float r0 = __fadd_rn(__fmul_rn(v00, ax0), __fmul_rn(v01, ax1));
        const float r1 = __fadd_rn(__fmul_rn(v10, ax0), __fmul_rn(v11, ax1));
        const float v = __fadd_rn(__fmul_rn(r0, by0), __fmul_rn(r1, by1));
        if (v > best || (l == 0 && !(v <= best))) { best = v; bi = l; }      // first maximum wins; a NaN at l = 0 stays (numpy)
    }
    return bi;
}

// one thread per output pixel: the work is the gather + interpolation (~100 instructions per pixel), not the 4 bytes written
// (four pixels per thread with packed 32-bit stores measured 25 % slower: fewer, longer dependent chains)
__global__ void __launch_bounds__(256) lp_post_kernel(LpPostArgs a) {
    const int64_t total = (int64_t)a.n * a.H * a.W;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int bi = lp_post_pixel(a, e);
        if (a.cls) a.cls[e] = (unsigned char)bi;
        if (a.rgb) {
            unsigned char c0 = (unsigned char)bi, c1 = (unsigned char)bi, c2 = (unsigned char)bi;
            if (a.pal) { c0 = a.pal[bi * 3]; c1 = a.pal[bi * 3 + 1]; c2 = a.pal[bi * 3 + 2]; }
            a.rgb[e * 3] = c0; a.rgb[e * 3 + 1] = c1; a.rgb[e * 3 + 2] = c2;
        }
    }
}

}  // namespace crw

using namespace crw;

// ---- JHMDB key points (utils/test_utils.py:60-84, process_pose) ----------------------------------------------------------
// Per frame and key-point channel c = 1..L-1: the top-k (k = min(hw, topk) <= 4) positions of the soft map, their values
// normalised to sum 1, and the value-weighted mean of their (x, y) - or (-1, -1) for a channel that is zero everywhere.
// One CTA per (frame, channel): every thread keeps the best k of its strided share, a shared-memory tree merges the sorted
// lists (larger value first, then the smaller position), thread 0 does the reference's float arithmetic in its order.
constexpr int kPoseK = 4;
struct PoseTop {
    float v[kPoseK];
    int p[kPoseK];
};
__device__ __forceinline__ bool pose_before(float va, int pa, float vb, int pb) { return va > vb || (va == vb && pa < pb); }
__device__ __forceinline__ void pose_insert(PoseTop& t, int k, float v, int p) {
    if (!pose_before(v, p, t.v[k - 1], t.p[k - 1])) return;
    int i = k - 1;
    while (i > 0 && pose_before(v, p, t.v[i - 1], t.p[i - 1])) { t.v[i] = t.v[i - 1]; t.p[i] = t.p[i - 1]; --i; }
    t.v[i] = v;
    t.p[i] = p;
}

__global__ void __launch_bounds__(256) lp_pose_kernel(const float* __restrict__ pred, int hw, int w, int L, int k, float* __restrict__ coords) {
    __shared__ float sv[256][kPoseK];
    __shared__ int sp[256][kPoseK];
    __shared__ int snz[256];
    const int tid = threadIdx.x;
    const int c = blockIdx.x % (L - 1), f = blockIdx.x / (L - 1);
    const float* src = pred + (int64_t)f * hw * L + (c + 1);
    PoseTop t;
#pragma unroll
    for (int i = 0; i < kPoseK; ++i) { t.v[i] = -INFINITY; t.p[i] = 0x7fffffff; }
    int nz = 0;
    for (int p = tid; p < hw; p += 256) {
        const float v = src[(int64_t)p * L];
        nz |= v != 0.f;
        pose_insert(t, k, v, p);
    }
#pragma unroll
    for (int i = 0; i < kPoseK; ++i) { sv[tid][i] = t.v[i]; sp[tid][i] = t.p[i]; }
    snz[tid] = nz;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) {
            for (int i = 0; i < k; ++i) pose_insert(t, k, sv[tid + s][i], sp[tid + s][i]);
            nz |= snz[tid + s];
        }
        __syncthreads();
        if (tid < s) {
#pragma unroll
            for (int i = 0; i < kPoseK; ++i) { sv[tid][i] = t.v[i]; sp[tid][i] = t.p[i]; }
            snz[tid] = nz;
        }
        __syncthreads();
    }
    if (tid == 0) {
        float x = -1.f, y = -1.f;
        if (nz) {                                    // flatlbls.sum(0) == 0 -> -1 (non-negative soft labels: the sum is 0 iff all are)
            float tot = 0.f;
            for (int i = 0; i < k; ++i) tot = __fadd_rn(tot, t.v[i]);
            x = 0.f;
            y = 0.f;
            for (int i = 0; i < k; ++i) {
                const float wgt = __fdiv_rn(t.v[i], tot);
                x = __fadd_rn(x, __fmul_rn((float)(t.p[i] % w), wgt));
                y = __fadd_rn(y, __fmul_rn((float)(t.p[i] / w), wgt));
            }
        }
        float* o = coords + (int64_t)f * 2 * (L - 1);
        o[c] = x;
        o[(L - 1) + c] = y;
    }
}

// test.py:162-164 (--norm_mask): pred -= min_L; pred /= max_L per position, in place (same two fp32 operations as the reference)
__global__ void __launch_bounds__(256) lp_minmax_kernel(float* maps, int64_t rows, int L) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float* p = maps + r * L;
        float mn = p[0];
        for (int l = 1; l < L; ++l) mn = fminf(mn, p[l]);
        float mx = p[0] - mn;
        for (int l = 1; l < L; ++l) mx = fmaxf(mx, p[l] - mn);
        for (int l = 0; l < L; ++l) p[l] = (p[l] - mn) / mx;
    }
}

extern "C" int crw_lp_minmax_normalize(float* maps, int64_t rows, int L, crw_stream_t stream) {
    if (rows < 0 || L <= 0) { set_error("lp_minmax_normalize: bad shape"); return CRW_ERR_SHAPE; }
    if (rows == 0) return CRW_OK;
    const int grid = (int)((rows + 255) / 256 < 148 * 8 ? (rows + 255) / 256 : 148 * 8);
    CRW_LAUNCH(lp_minmax_kernel, grid, 256, 0, stream, maps, rows, L);
    return check_launch("lp_minmax_normalize");
}

extern "C" int crw_lp_pose_coords(const float* pred, int n, int h, int w, int L, int topk, float* coords, crw_stream_t stream) {
    if (n < 0 || h <= 0 || w <= 0 || L < 1 || topk < 1) { set_error("lp_pose_coords: bad shape"); return CRW_ERR_SHAPE; }
    const int64_t hw = (int64_t)h * w;
    const int k = hw < topk ? (int)hw : topk;
    if (k > kPoseK || hw > 0x7ffffffe) { set_error("lp_pose_coords: top-k up to %d, got %d", kPoseK, k); return CRW_ERR_UNSUPPORTED; }
    if (n == 0 || L == 1) return CRW_OK;
    CRW_LAUNCH(lp_pose_kernel, n * (L - 1), 256, 0, stream, pred, (int)hw, w, L, k, coords);
    return check_launch("lp_pose_coords");
}

extern "C" int crw_lp_upsample_argmax(const float* pred, int n, int h, int w, int L, int H, int W, int norm_mask,
                                      const unsigned char* palette, unsigned char* cls, unsigned char* rgb, crw_stream_t stream) {
    if (n < 0 || h <= 0 || w <= 0 || L <= 0 || H <= 0 || W <= 0) { set_error("lp_upsample_argmax: bad shape"); return CRW_ERR_SHAPE; }
    if (L > 255) { set_error("lp_upsample_argmax: at most 255 labels (byte class map), got %d", L); return CRW_ERR_UNSUPPORTED; }
    if (n == 0 || (!cls && !rgb)) return CRW_OK;
    LpPostArgs a{};
    a.pred = pred; a.pal = palette; a.cls = cls; a.rgb = rgb;
    a.n = n; a.h = h; a.w = w; a.L = L; a.H = H; a.W = W; a.norm_mask = norm_mask;
    a.scale_x = 1.0 / ((double)W / (double)w);          // OpenCV: inv_scale = dsize / ssize, scale = 1 / inv_scale
    a.scale_y = 1.0 / ((double)H / (double)h);
    const int64_t total = (int64_t)n * H * W;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    CRW_LAUNCH(lp_post_kernel, grid, 256, 0, stream, a);
    return check_launch("lp_upsample_argmax");
}
