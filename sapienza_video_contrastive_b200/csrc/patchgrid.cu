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

constexpr int PG_THREADS = 2 * PG_MAX * 3;         // two groups of (column, channel) threads, alternating rows

__global__ void __launch_bounds__(PG_THREADS) patch_grid_kernel(PgArgs a) {
    __shared__ unsigned char crop[PG_MAX * PG_MAX * 3];
    __shared__ unsigned char tmp[PG_MAX * PG_MAX * 3];
    __shared__ __align__(16) int kx[PG_MAX][PG_TAPS + 1], ky[PG_MAX][PG_TAPS + 1];      // [.][3] = first tap
    __shared__ float lut[3][256];                       // ToTensor + Normalize of the 256 possible bytes per channel
    const int p = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int* b = a.boxes + ((int64_t)f * a.P + p) * 4;
    const int top = b[0], left = b[1], ch = b[2], cw = b[3];
    const int wy = (p / a.nwx) * a.stride, wx = (p % a.nwx) * a.stride;
    const int ps = a.ps;
    // the two thread groups take alternate rows; inside a group a thread owns one byte column (pixel column x channel)
    const int grp = tid >= PG_MAX * 3 ? 1 : 0, col = tid - grp * (PG_MAX * 3);
    const int rowb = cw * 3, outb = ps * 3;
    const unsigned char* src = a.frames + (((int64_t)f * a.H + wy + top) * a.W + wx + left) * 3;
    if (col < rowb) {
#pragma unroll 8
        for (int r = grp; r < ch; r += 2) crop[r * rowb + col] = src[(int64_t)r * a.W * 3 + col];     // (independent loads, in flight together)
    }
    // the resampling coefficients (double arithmetic, Pillow's) and the value table are computed while the crop travels
    if (tid < ps) {
        int lo, n, k[PG_TAPS];
        pg_coeffs(cw, ps, tid, lo, n, k);
        kx[tid][0] = k[0]; kx[tid][1] = k[1]; kx[tid][2] = k[2]; kx[tid][3] = lo;
    } else if (tid >= 64 && tid < 64 + ps) {
        int lo, n, k[PG_TAPS];
        pg_coeffs(ch, ps, tid - 64, lo, n, k);
        ky[tid - 64][0] = k[0]; ky[tid - 64][1] = k[1]; ky[tid - 64][2] = k[2]; ky[tid - 64][3] = lo;
    }
    for (int e = tid; e < 768; e += PG_THREADS) {
        const int c = e >> 8;
        lut[c][e & 255] = ((float)(e & 255) / 255.0f - a.mean[c]) / a.stdv[c];
    }
    __syncthreads();
    // horizontal pass: (ch rows) x (ps columns) x 3, 8-bit result
    if (col < outb) {
        const int xx = col / 3, c = col - xx * 3;
        const int lo = kx[xx][3], k0 = kx[xx][0], k1 = kx[xx][1], k2 = kx[xx][2];
        // (taps past the edge carry weight 0)
        const int o0 = (lo < cw ? lo : cw - 1) * 3 + c, o1 = (lo + 1 < cw ? lo + 1 : cw - 1) * 3 + c, o2 = (lo + 2 < cw ? lo + 2 : cw - 1) * 3 + c;
        for (int r = grp; r < ch; r += 2) {
            const unsigned char* q = crop + r * rowb;
            const int acc = (1 << 21) + (int)q[o0] * k0 + (int)q[o1] * k1 + (int)q[o2] * k2;
            tmp[r * outb + col] = pg_clip8(acc >> 22);
        }
    }
    __syncthreads();
    // vertical pass + ToTensor + Normalize, channel-first output: a thread owns (channel, column), a warp stores 32 neighbouring columns
    if (col < outb) {
        const int c = col / ps, xx = col - c * ps;
        const unsigned char* q = tmp + xx * 3 + c;
        float* dst = a.out + (((int64_t)f * a.P + p) * 3 + c) * ps * ps + xx;
        const float* lc = lut[c];
        for (int yy = grp; yy < ps; yy += 2) {
            const int4 k = *reinterpret_cast<const int4*>(ky[yy]);
            const int lo = k.w;
            const int y0 = lo < ch ? lo : ch - 1, y1 = lo + 1 < ch ? lo + 1 : ch - 1, y2 = lo + 2 < ch ? lo + 2 : ch - 1;
            const int acc = (1 << 21) + (int)q[y0 * outb] * k.x + (int)q[y1 * outb] * k.y + (int)q[y2 * outb] * k.z;
            dst[yy * ps] = lc[pg_clip8(acc >> 22)];
        }
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
    CRW_LAUNCH(patch_grid_kernel, grid, PG_THREADS, 0, stream, a);
    return check_launch("patch_grid");
}
