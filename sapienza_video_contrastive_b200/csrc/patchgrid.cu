// patchgrid.cu - the patch-grid producer of the training pipeline: code/utils/augs.py:59-82 (patch_grid:
// skimage view_as_windows -> per window RandomResizedCrop(64, scale (0.7, 0.9)) -> ToTensor -> Normalize), i.e. the CPU
// data-loader work that builds the (B, T, 49*3, 64, 64) tensor CRW.forward consumes (model.py:348-349).
//
// The random crop boxes are index math on the host (drawn there with the reference's own generator calls, see augs.py in the
// package); this kernel does the pixel work of all windows of all frames in one launch: crop -> Pillow's BILINEAR resize
// (Pillow 12.2, src/libImaging/Resample.c: two passes, horizontal first, 8-bit intermediate, fixed-point coefficients with
// PRECISION_BITS = 22, triangle filter, support max(scale, 1)) -> x / 255 -> (x - mean) / std, written channel-first.
// One CTA = one window of one frame; the crop (<= 64 x 64 x 3 bytes) and the 8-bit intermediate live in shared memory.
// Bit-exact with PIL on the 8-bit image, hence bit-exact floats.
#include "common.cuh"

namespace crw {

constexpr int PG_MAX = 64;          // window / output side supported (the reference's 64)
constexpr int PG_TAPS = 3;          // crop side <= output side: support 1, at most 3 taps

struct PgArgs {
    const unsigned char* frames;    // (F, H, W, 3) uint8
    const int* boxes;               // (F, P, 4): top, left, height, width of the crop inside its window
    float* out;                     // (F, P * 3, ps, ps)
    int F, H, W, win, stride, nwx, P, ps;
    float mean[3], stdv[3];
};

// Pillow precompute_coeffs for one output index (bilinear): bounds and normalised fixed-point taps
__device__ __forceinline__ void pg_coeffs(int in_size, int out_size, int xx, int& xmin, int& n, int* kk) {
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    hi -= lo;
    double w[PG_TAPS + 1], ww = 0.0;
    for (int x = 0; x < hi && x <= PG_TAPS; ++x) {
        double v = (x + lo - center + 0.5) * ss;
        v = v < 0.0 ? -v : v;
        w[x] = v < 1.0 ? 1.0 - v : 0.0;
        ww += w[x];
    }
    xmin = lo;
    n = hi > PG_TAPS ? PG_TAPS : hi;
    for (int x = 0; x < PG_TAPS; ++x) {
        double c = x < n ? (ww != 0.0 ? w[x] / ww : w[x]) : 0.0;
        kk[x] = c < 0.0 ? (int)(-0.5 + c * (double)(1 << 22)) : (int)(0.5 + c * (double)(1 << 22));
    }
}

__device__ __forceinline__ unsigned char pg_clip8(int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void __launch_bounds__(256) patch_grid_kernel(PgArgs a) {
    __shared__ unsigned char crop[PG_MAX * PG_MAX * 3];
    __shared__ unsigned char tmp[PG_MAX * PG_MAX * 3];
    __shared__ int kx[PG_MAX][PG_TAPS + 1], ky[PG_MAX][PG_TAPS + 1];      // [.][3] = first tap
    const int p = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int* b = a.boxes + ((int64_t)f * a.P + p) * 4;
    const int top = b[0], left = b[1], ch = b[2], cw = b[3];
    const int wy = (p / a.nwx) * a.stride, wx = (p % a.nwx) * a.stride;
    const int ps = a.ps;
    if (tid < ps) {
        int lo, n, k[PG_TAPS];
        pg_coeffs(cw, ps, tid, lo, n, k);
        kx[tid][0] = k[0]; kx[tid][1] = k[1]; kx[tid][2] = k[2]; kx[tid][3] = lo;
    } else if (tid >= 64 && tid < 64 + ps) {
        int lo, n, k[PG_TAPS];
        pg_coeffs(ch, ps, tid - 64, lo, n, k);
        ky[tid - 64][0] = k[0]; ky[tid - 64][1] = k[1]; ky[tid - 64][2] = k[2]; ky[tid - 64][3] = lo;
    }
    const unsigned char* src = a.frames + (((int64_t)f * a.H + wy + top) * a.W + wx + left) * 3;
    for (int e = tid; e < ch * cw * 3; e += blockDim.x) {
        const int r = e / (cw * 3), c = e - r * (cw * 3);
        crop[e] = src[(int64_t)r * a.W * 3 + c];
    }
    __syncthreads();
    // horizontal pass: (ch rows) x (ps columns) x 3, 8-bit result
    for (int e = tid; e < ch * ps * 3; e += blockDim.x) {
        const int r = e / (ps * 3), rem = e - r * (ps * 3), xx = rem / 3, c = rem - xx * 3;
        const int lo = kx[xx][3];
        int acc = 1 << 21;
#pragma unroll
        for (int t = 0; t < PG_TAPS; ++t) {
            const int x = lo + t < cw ? lo + t : cw - 1;                    // (taps past the edge carry weight 0)
            acc += (int)crop[(r * cw + x) * 3 + c] * kx[xx][t];
        }
        tmp[e] = pg_clip8(acc >> 22);
    }
    __syncthreads();
    // vertical pass + ToTensor + Normalize, channel-first output
    float* dst = a.out + ((int64_t)f * a.P + p) * 3 * ps * ps;
    for (int e = tid; e < 3 * ps * ps; e += blockDim.x) {
        const int c = e / (ps * ps), rem = e - c * (ps * ps), yy = rem / ps, xx = rem - yy * ps;
        const int lo = ky[yy][3];
        int acc = 1 << 21;
#pragma unroll
        for (int t = 0; t < PG_TAPS; ++t) {
            const int y = lo + t < ch ? lo + t : ch - 1;
            acc += (int)tmp[(y * ps + xx) * 3 + c] * ky[yy][t];
        }
        const float v = (float)pg_clip8(acc >> 22) / 255.0f;
        dst[e] = (v - a.mean[c]) / a.stdv[c];
    }
}

}  // namespace crw

using namespace crw;

extern "C" int crw_patch_grid(const unsigned char* frames, const int* boxes, int F, int H, int W, int win, int stride, int out_size,
                              const float* mean3, const float* std3, float* out, crw_stream_t stream) {
    if (F < 0 || H <= 0 || W <= 0 || win <= 0 || stride <= 0 || out_size <= 0 || !mean3 || !std3) { set_error("patch_grid: bad arguments"); return CRW_ERR_SHAPE; }
    if (win > PG_MAX || out_size > PG_MAX || out_size < win || H < win || W < win) {
        set_error("patch_grid: windows up to %d px, output side >= window side (crops are enlarged, as in the reference), got win=%d out=%d", PG_MAX, win, out_size);
        return CRW_ERR_UNSUPPORTED;
    }
    if (F == 0) return CRW_OK;
    PgArgs a{};
    a.frames = frames; a.boxes = boxes; a.out = out; a.F = F; a.H = H; a.W = W; a.win = win; a.stride = stride; a.ps = out_size;
    a.nwx = (W - win) / stride + 1;
    a.P = a.nwx * ((H - win) / stride + 1);
    for (int c = 0; c < 3; ++c) { a.mean[c] = mean3[c]; a.stdv[c] = std3[c]; }
    dim3 grid(a.P, F);
    CRW_LAUNCH(patch_grid_kernel, grid, 256, 0, stream, a);
    return check_launch("patch_grid");
}
