// slic.cu - the superpixel label-map producer of the training pipeline (second half):
// code/data/superpixels.py:9-16 (compute_sp_slic: cv2.normalize(0..255, MINMAX, 8U) -> skimage.segmentation.slic) for every
// frame of a clip (compute_mask, superpixels.py:24-63), i.e. the CPU data-loader work that builds the label maps CRW's
// superpixel pooling consumes (model.py:136-200).
//
// scikit-image is a dependency that is not in this image: the kernels follow the PUBLISHED algorithm (scikit-image >= 0.19
// slic_superpixels.py / _slic.pyx / _regular_grid.py; Achanta et al., TPAMI 2012) (the tests hold a CPU restatement of it), with the
// deviations DESIGN.md section 2 lists (features quantised to 2^-24 so that the centre update is an exact integer sum; a fixed Newton cube
// root; empty segments stay empty).  All floating-point work is IEEE double with explicit, unfused operations in a fixed order,
// so that the label maps are that restatement's bit for bit.
//
//   slic_minmax    global min / max of each frame (ordered-integer atomics)
//   slic_features  min-max normalise to 8 bit (OpenCV's single-rounding float multiply-add), sRGB -> linear by table, XYZ, Lab,
//                  scale by 1 / compactness, quantise                               -> 3 double planes per frame
//   slic_assign    one launch per iteration: every CTA rebuilds the K centres (from the grid, or from the previous iteration's
//                  integer sums) in shared memory, each thread finds the nearest centre of its pixels among those whose
//                  2-step window holds the pixel (centres in increasing order, strict comparison = the sequential scatter of
//                  _slic.pyx), and the pixel is added to that centre's integer sums (warp reduction per segment, then shared-memory
//                  and global 64-bit atomics)
//   slic_connect   enforce_connectivity: the scan-order breadth-first relabelling depends on the queue order, which one warp per
//                  frame reproduces exactly while popping 32 queue entries at a time (frames in parallel; queue and the per-batch
//                  pixel table in shared memory)
#include <math.h>

#include "common.cuh"

namespace crw {

constexpr int SLIC_CHUNK = 32;          // frames per launch (their grid parameters travel as kernel arguments)
constexpr int SLIC_MAX_K = 2048;        // centres per frame (shared-memory budget of slic_assign)
constexpr int SLIC_QBITS = 24;
constexpr int SLIC_CBRT_ITERS = 12;
constexpr int SLIC_THREADS = 256;
constexpr int SLIC_PIX = 4;             // pixels per thread in slic_assign

struct SlicGrid { int y0, sy, x0, sx, ny, nx; };
struct SlicFrames {
    SlicGrid g[SLIC_CHUNK];
    double sw[SLIC_CHUNK];              // 1 / step^2
};

#ifdef CRW_SIM
__device__ __forceinline__ double dmul(double a, double b) { volatile double r = a * b; return r; }
__device__ __forceinline__ double dadd(double a, double b) { volatile double r = a + b; return r; }
__device__ __forceinline__ double ddiv(double a, double b) { volatile double r = a / b; return r; }
__device__ __forceinline__ long long d2ll_rn(double a) { return llrint(a); }
__device__ __forceinline__ int f2i_rn(float a) { return (int)lrintf(a); }
#else
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ long long d2ll_rn(double a) { return __double2ll_rn(a); }
__device__ __forceinline__ int f2i_rn(float a) { return __float2int_rn(a); }
#endif
__device__ __forceinline__ double dsub(double a, double b) { return dadd(a, -b); }

__device__ __forceinline__ unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// order-preserving map float -> uint32 (and back)
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// mm[f] = {max of ~ord (i.e. the minimum), max of ord}, both zero-initialised
__global__ void __launch_bounds__(SLIC_THREADS) slic_minmax_kernel(const float* __restrict__ video, int64_t n_per_frame, unsigned* __restrict__ mm) {
    const int f = blockIdx.y;
    const float* v = video + (int64_t)f * n_per_frame;
    unsigned lo = 0u, hi = 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_frame; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned o = f2ord(v[i]);
        lo = umax(lo, ~o);
        hi = umax(hi, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = umax(lo, __shfl_xor_sync(kFull, lo, o));
        hi = umax(hi, __shfl_xor_sync(kFull, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(mm + 2 * f, lo);
        atomicMax(mm + 2 * f + 1, hi);
    }
}

__device__ __forceinline__ double slic_cbrt(double x) {
    double y = ddiv(dadd(x, 2.0), 3.0);
#pragma unroll 1
    for (int i = 0; i < SLIC_CBRT_ITERS; ++i) y = ddiv(dadd(dmul(2.0, y), ddiv(x, dmul(y, y))), 3.0);
    return y;
}
__device__ __forceinline__ double slic_labf(double v) {
    return v > 0.008856 ? slic_cbrt(v) : dadd(dmul(7.787, v), 16.0 / 116.0);
}

// video (F, 3, H, W) fp32 -> feat (F, 3, H*W) double, quantised to 2^-24
__global__ void __launch_bounds__(SLIC_THREADS) slic_features_kernel(const float* __restrict__ video, int hw, const unsigned* __restrict__ mm,
                                                                       const double* __restrict__ lin_table, double inv_compactness,
                                                                       double* __restrict__ feat) {
    __shared__ double lin[256];
    const int f = blockIdx.y;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lin[i] = lin_table[i];
    __syncthreads();
    const double mn = (double)ord2f(~mm[2 * f]), mx = (double)ord2f(mm[2 * f + 1]);
    const double range = dsub(mx, mn);
    const double scale = range > 2.220446049250313e-16 ? dmul(255.0, ddiv(1.0, range)) : 0.0;   // OpenCV normalize(): MINMAX
    const double shift = dsub(0.0, dmul(mn, scale));
    const float fs = (float)scale, fb = (float)shift;
    const float* v = video + (int64_t)f * 3 * hw;
    double* o = feat + (int64_t)f * 3 * hw;
    const double Qs = (double)(1 << SLIC_QBITS), Qi = 1.0 / (double)(1 << SLIC_QBITS);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
        double l[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int q = f2i_rn(__fmaf_rn(v[(int64_t)c * hw + p], fs, fb));
            q = q < 0 ? 0 : (q > 255 ? 255 : q);
            l[c] = lin[q];
        }
        const double x = ddiv(dadd(dadd(dmul(l[0], 0.412453), dmul(l[1], 0.357580)), dmul(l[2], 0.180423)), 0.95047);
        const double y = ddiv(dadd(dadd(dmul(l[0], 0.212671), dmul(l[1], 0.715160)), dmul(l[2], 0.072169)), 1.0);
        const double z = ddiv(dadd(dadd(dmul(l[0], 0.019334), dmul(l[1], 0.119193)), dmul(l[2], 0.950227)), 1.08883);
        const double fx = slic_labf(x), fy = slic_labf(y), fz = slic_labf(z);
        const double lab[3] = {dsub(dmul(116.0, fy), 16.0), dmul(500.0, dsub(fx, fy)), dmul(200.0, dsub(fy, fz))};
#pragma unroll
        for (int c = 0; c < 3; ++c)
            o[(int64_t)c * hw + p] = dmul((double)d2ll_rn(dmul(dmul(lab[c], inv_compactness), Qs)), Qi);
    }
}

// sums layout per frame and iteration: [K][6] u64 = {n, sum y, sum x, sum f0, sum f1, sum f2} (two's complement)
__global__ void __launch_bounds__(SLIC_THREADS) slic_assign_kernel(SlicFrames fr, int frame0, int H, int W, int it, int Kmax,
                                                                     const double* __restrict__ feat, const unsigned long long* __restrict__ prev,
                                                                     unsigned long long* __restrict__ next, int* __restrict__ nearest,
                                                                     int* __restrict__ kinfo) {
    CRW_DYN_SMEM(smem);
    const int fl = blockIdx.y, f = frame0 + fl;
    const SlicGrid g = fr.g[fl];
    const int K = g.ny * g.nx, hw = H * W;
    if (it == 0 && blockIdx.x == 0 && threadIdx.x == 0) kinfo[f] = K;          // for the connectivity pass
    int* win = reinterpret_cast<int*>(smem);                                         // [K][4]: ya, yb, xa, xb (16-byte rows)
    double* cen = reinterpret_cast<double*>(win + (size_t)K * 4);                    // [K][5]: y, x, c0, c1, c2
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(cen + (size_t)K * 5);   // [K][6]
    const unsigned long long* pv = prev + (int64_t)f * Kmax * 6;
    const double Qi = 1.0 / (double)(1 << SLIC_QBITS);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double cy, cx, c0 = 0.0, c1 = 0.0, c2 = 0.0;
        bool alive = true;
        if (it == 0) {
            cy = (double)(g.y0 + (k / g.nx) * g.sy);
            cx = (double)(g.x0 + (k % g.nx) * g.sx);
        } else {
            const long long n = (long long)pv[k * 6];
            alive = n > 0;
            const double nn = (double)(n > 0 ? n : 1);
            cy = ddiv((double)(long long)pv[k * 6 + 1], nn);
            cx = ddiv((double)(long long)pv[k * 6 + 2], nn);
            c0 = dmul(ddiv((double)(long long)pv[k * 6 + 3], nn), Qi);
            c1 = dmul(ddiv((double)(long long)pv[k * 6 + 4], nn), Qi);
            c2 = dmul(ddiv((double)(long long)pv[k * 6 + 5], nn), Qi);
        }
        cen[k * 5] = cy; cen[k * 5 + 1] = cx; cen[k * 5 + 2] = c0; cen[k * 5 + 3] = c1; cen[k * 5 + 4] = c2;
        int ya = 0, yb = 0, xa = 0, xb = 0;
        if (alive) {
            ya = (int)fmax(dsub(cy, (double)(2 * g.sy)), 0.0);
            yb = (int)fmin(dadd(dadd(cy, (double)(2 * g.sy)), 1.0), (double)H);
            xa = (int)fmax(dsub(cx, (double)(2 * g.sx)), 0.0);
            xb = (int)fmin(dadd(dadd(cx, (double)(2 * g.sx)), 1.0), (double)W);
        }
        win[k * 4] = ya; win[k * 4 + 1] = yb; win[k * 4 + 2] = xa; win[k * 4 + 3] = xb;
    }
    for (int i = threadIdx.x; i < K * 6; i += blockDim.x) acc[i] = 0ull;
    __syncthreads();
    const double sw = fr.sw[fl];
    const double* ft = feat + (int64_t)f * 3 * hw;
    int* nr = nearest + (int64_t)f * hw;
    const int base = blockIdx.x * (SLIC_THREADS * SLIC_PIX);
#pragma unroll 1
    for (int j = 0; j < SLIC_PIX; ++j) {
        const int p = base + j * SLIC_THREADS + threadIdx.x;
        const bool valid = p < hw;                                 // (no early exit: the warp reduces together below)
        int y = 0, x = 0, bk = -1;
        long long q0 = 0, q1 = 0, q2 = 0;
        if (valid) {
            y = p / W; x = p - y * W;
            const double f0 = ft[p], f1 = ft[(int64_t)hw + p], f2 = ft[(int64_t)2 * hw + p];
            double best = 1.7976931348623157e308;
            for (int k = 0; k < K; ++k) {
                const int4 w4 = *reinterpret_cast<const int4*>(win + k * 4);
                if (y < w4.x || y >= w4.y || x < w4.z || x >= w4.w) continue;
                const double* c = cen + k * 5;
                const double dy = dsub(c[0], (double)y), dx = dsub(c[1], (double)x);
                double d = dmul(dadd(dmul(dy, dy), dmul(dx, dx)), sw);
                const double e0 = dsub(f0, c[2]), e1 = dsub(f1, c[3]), e2 = dsub(f2, c[4]);
                d = dadd(d, dadd(dadd(dmul(e0, e0), dmul(e1, e1)), dmul(e2, e2)));
                if (best > d) { best = d; bk = k; }
            }
            if (bk < 0) bk = nr[p];                  // no window holds the pixel: it keeps its segment
            else nr[p] = bk;
            q0 = d2ll_rn(dmul(f0, (double)(1 << SLIC_QBITS)));
            q1 = d2ll_rn(dmul(f1, (double)(1 << SLIC_QBITS)));
            q2 = d2ll_rn(dmul(f2, (double)(1 << SLIC_QBITS)));
        }
        // a warp is 32 neighbouring pixels: one segment, sometimes two or three.  Per segment present the warp adds up first and
        // one lane issues the six shared-memory atomics (64-bit shared atomics are compare-and-swap loops: 32 lanes on one
        // address serialise completely); the sums are integers, so the order does not matter.
        unsigned todo = __ballot_sync(kFull, valid);
        while (todo) {
            const int leader = __ffs((int)todo) - 1;
            const int bkl = __shfl_sync(kFull, bk, leader);
            const bool mine = valid && bk == bkl;
            const unsigned m = __ballot_sync(kFull, mine);
            long long sy = mine ? y : 0, sx = mine ? x : 0, s0 = mine ? q0 : 0, s1 = mine ? q1 : 0, s2 = mine ? q2 : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sy += __shfl_xor_sync(kFull, sy, o); sx += __shfl_xor_sync(kFull, sx, o);
                s0 += __shfl_xor_sync(kFull, s0, o); s1 += __shfl_xor_sync(kFull, s1, o); s2 += __shfl_xor_sync(kFull, s2, o);
            }
            if ((int)(threadIdx.x & 31) == leader) {
                unsigned long long* a = acc + bkl * 6;
                atomicAdd(a, (unsigned long long)__popc(m));
                atomicAdd(a + 1, (unsigned long long)sy);
                atomicAdd(a + 2, (unsigned long long)sx);
                atomicAdd(a + 3, (unsigned long long)s0);
                atomicAdd(a + 4, (unsigned long long)s1);
                atomicAdd(a + 5, (unsigned long long)s2);
            }
            todo &= ~m;
        }
    }
    __syncthreads();
    unsigned long long* nx = next + (int64_t)f * Kmax * 6;
    for (int i = threadIdx.x; i < K * 6; i += blockDim.x)
        if (acc[i] != 0ull) atomicAdd(nx + i, acc[i]);
}

__global__ void slic_offset_kernel(const int* __restrict__ nearest, int64_t n, int* __restrict__ labels) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) labels[i] = nearest[i] + 1;
}

__device__ __forceinline__ int ld_cg_i(const int* p) {
#ifdef CRW_SIM
    return *p;
#else
    return __ldcg(p);
#endif
}

// _enforce_label_connectivity_cython, one WARP per frame.  con (= the output labels) starts at -1.
// The reference algorithm is a scan-order breadth-first search whose result depends on the queue order (the max_size cut, the
// "last labelled neighbour met" merge rule), so the queue order is reproduced exactly: the warp pops up to 32 queue entries at
// once (lane = pop order), every lane looks at its 4 neighbours (x+1, x-1, y+1, y-1), and a pixel reachable from several pops is
// pushed by the earliest (lane, neighbour) - a small shared-memory table keyed by pixel decides -, at the queue position a prefix sum over that
// order assigns.  Pushes only turn con from -1 into the current label, which the merge rule ignores, so looking at the 32 pops
// side by side reads the same values as one after the other.
constexpr int SLIC_QSMEM = 11264;       // queue entries kept in shared memory (a larger max_size uses the workspace in global memory)
constexpr int SLIC_HASH = 256;          // slots of the per-batch pixel table (at most 128 candidates)

__global__ void __launch_bounds__(32) slic_connect_kernel(const int* __restrict__ kinfo, int min_size, int max_size, int H, int W,
                                                          const int* __restrict__ nearest, int* __restrict__ queue, int* __restrict__ labels) {
    __shared__ int qs[SLIC_QSMEM];
    __shared__ int hkey[SLIC_HASH], hord[SLIC_HASH];
    const int f = blockIdx.x, lane = threadIdx.x, hw = H * W;
    if (kinfo) {                                                  // slic(): 0.5 and 3 times the mean segment size
        const double seg_size = (double)hw / (double)kinfo[f];
        min_size = (int)(0.5 * seg_size);
        max_size = (int)(3.0 * seg_size);
    }
    const int* seg = nearest + (int64_t)f * hw;
    int* con = labels + (int64_t)f * hw;
    int* q = max_size <= SLIC_QSMEM ? qs : queue + (int64_t)f * hw;      // the search never holds more than max_size pixels
    for (int i = lane; i < SLIC_HASH; i += 32) { hkey[i] = -1; hord[i] = 0x7fffffff; }
    __syncwarp();
    int new_label = 1, scan = 0;
    while (true) {
        int p0 = -1;                                              // next pixel in scan order without a label
        while (scan < hw) {
            const int p = scan + lane;
            const unsigned m = __ballot_sync(kFull, p < hw && con[p] < 0);
            if (m) { p0 = scan + __ffs((int)m) - 1; break; }
            scan += 32;
        }
        if (p0 < 0) break;
        scan = p0 + 1;
        const int label = seg[p0];
        if (lane == 0) { con[p0] = new_label; q[0] = p0; }
        __syncwarp();
        int size = 1, visited = 0, adjacent = 0;
        while (visited < size && size < max_size) {
            const int nb = size - visited < 32 ? size - visited : 32;
            int cand[4] = {-1, -1, -1, -1}, slot[4] = {0, 0, 0, 0};
            int adj = -1;
            if (lane < nb) {
                const int p = q[visited + lane];
                const int y = p / W, x = p - y * W;
                int pps[4], cs[4], ss[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {                     // all eight loads first: one round trip to L2, not two
                    const int xx = x + (i == 0 ? 1 : (i == 1 ? -1 : 0)), yy = y + (i == 2 ? 1 : (i == 3 ? -1 : 0));
                    const bool in = xx >= 0 && xx < W && yy >= 0 && yy < H;
                    pps[i] = in ? yy * W + xx : -1;
                    cs[i] = in ? con[pps[i]] : 0;
                    ss[i] = in ? seg[pps[i]] : -1;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (pps[i] < 0) continue;
                    const int pp = pps[i], c = cs[i];
                    if (c == -1) {
                        if (ss[i] == label) {
                            // a pixel several pops can reach is pushed by the earliest (pop, neighbour): find or claim the pixel's
                            // slot in the batch table, then the smallest order wins (shared-memory atomics: no trip to L2)
                            cand[i] = pp;
                            int h = (int)(((unsigned)pp * 2654435761u) >> 24);
                            while (true) {
                                const int old = atomicCAS(hkey + h, -1, pp);
                                if (old == -1 || old == pp) break;
                                h = (h + 1) & (SLIC_HASH - 1);
                            }
                            slot[i] = h;
                            atomicMin(hord + h, lane * 4 + i);
                        }
                    } else if (c != new_label) {
                        adj = c;
                    }
                }
            }
            __syncwarp();
            bool win[4];
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                win[i] = cand[i] >= 0 && hord[slot[i]] == lane * 4 + i;
                cnt += win[i] ? 1 : 0;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (cand[i] >= 0) { hkey[slot[i]] = -1; hord[slot[i]] = 0x7fffffff; }      // the table is empty again for the next batch
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(kFull, incl, 31);
            int pos = size + incl - cnt;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!win[i]) continue;
                if (pos < max_size) { con[cand[i]] = new_label; q[pos] = cand[i]; }      // beyond max_size: the pixel stays free for a later segment
                ++pos;
            }
            const unsigned am = __ballot_sync(kFull, adj >= 0);
            if (am) adjacent = __shfl_sync(kFull, adj, 31 - __clz((int)am));
            size = size + total < max_size ? size + total : max_size;
            visited += nb;
            __syncwarp();
        }
        if (size < min_size) {                                    // too small: merged into the last labelled neighbour met
            for (int i = lane; i < size; i += 32) con[q[i]] = adjacent;
        } else {
            ++new_label;
        }
        __syncwarp();
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------

// skimage.util.regular_grid((1, H, W), n) -> slices along y and x
static void slic_regular_grid(int H, int W, int n, SlicGrid& g, double& sw) {
    double sd[3] = {1.0, (double)(H < W ? H : W), (double)(H < W ? W : H)};
    const bool y_first = H <= W;                                        // stable sort: y before x when equal
    double steps[3];
    int starts[3], isteps[3];
    double space = sd[0] * sd[1] * sd[2];
    if (space <= (double)n) {
        starts[0] = starts[1] = starts[2] = 0;
        isteps[0] = isteps[1] = isteps[2] = 1;
    } else {
        const double s = pow(space / n, 1.0 / 3);
        steps[0] = steps[1] = steps[2] = s;
        if (sd[0] < s || sd[1] < s || sd[2] < s) {
            for (int d = 0; d < 2; ++d) {
                steps[d] = sd[d];
                space = 1.0;
                for (int e = d + 1; e < 3; ++e) space *= sd[e];
                const double t = pow(space / n, 1.0 / (3 - d - 1));
                for (int e = d + 1; e < 3; ++e) steps[e] = t;
                if (sd[0] >= steps[0] && sd[1] >= steps[1] && sd[2] >= steps[2]) break;
            }
        }
        for (int d = 0; d < 3; ++d) {
            starts[d] = (int)floor(steps[d] / 2.0);
            isteps[d] = (int)nearbyint(steps[d]);
        }
    }
    const int iy = y_first ? 1 : 2, ix = y_first ? 2 : 1;
    g.y0 = starts[iy]; g.sy = isteps[iy]; g.x0 = starts[ix]; g.sx = isteps[ix];
    g.ny = g.y0 < H ? (H - g.y0 + g.sy - 1) / g.sy : 0;
    g.nx = g.x0 < W ? (W - g.x0 + g.sx - 1) / g.sx : 0;
    int step = g.sy > g.sx ? g.sy : g.sx;
    if (step < 1) step = 1;
    sw = 1.0 / ((double)step * (double)step);
}

// skimage.color.rgb2xyz's gamma expansion of the 256 possible inputs (libm pow, as the CPU restatement)
struct SlicLinTable {
    double v[256];
    SlicLinTable() {
        for (int i = 0; i < 256; ++i) {
            const double x = i / 255.0;
            v[i] = x > 0.04045 ? pow((x + 0.055) / 1.055, 2.4) : x / 12.92;
        }
    }
};
static const double* slic_lin_table() {
    static const SlicLinTable t;
    return t.v;
}

struct SlicLayout { size_t mm, kinfo, table, feat, nearest, queue, sums, total; int Kmax; };

static int slic_layout(int F, int H, int W, const int* n_segments, int n_iter, SlicLayout& L) {
    int Kmax = 1;
    for (int f = 0; f < F; ++f) {
        if (n_segments[f] < 1) return 1;
        SlicGrid g; double sw;
        slic_regular_grid(H, W, n_segments[f], g, sw);
        const int K = g.ny * g.nx;
        if (K < 1 || K > SLIC_MAX_K) return 1;
        if (K > Kmax) Kmax = K;
    }
    const size_t hw = (size_t)H * W;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    L.mm = off; off += up((size_t)F * 2 * sizeof(unsigned));
    L.kinfo = off; off += up((size_t)F * sizeof(int));
    L.table = off; off += up(256 * sizeof(double));
    L.feat = off; off += up((size_t)F * 3 * hw * sizeof(double));
    L.nearest = off; off += up((size_t)F * hw * sizeof(int));
    L.queue = off; off += up((size_t)F * hw * sizeof(int));
    L.sums = off; off += up((size_t)n_iter * F * Kmax * 6 * sizeof(unsigned long long));
    L.total = off;
    L.Kmax = Kmax;
    return 0;
}

}  // namespace crw

using namespace crw;

extern "C" size_t crw_slic_workspace_bytes(int F, int H, int W, const int* n_segments, int n_iter) {
    SlicLayout L;
    if (F < 1 || H < 1 || W < 1 || n_iter < 1 || !n_segments || slic_layout(F, H, W, n_segments, n_iter, L)) return 0;
    return L.total;
}

extern "C" int crw_slic(const float* video, int F, int H, int W, const int* n_segments, double compactness, int n_iter, int connectivity,
                        int* labels, void* ws, size_t ws_bytes, void* stream) {
    SlicLayout L;
    if (F < 1 || H < 1 || W < 1 || n_iter < 1 || !n_segments || !(compactness > 0.0) || (int64_t)H * W > (1 << 30)) {
        set_error("crw_slic: bad arguments (F=%d H=%d W=%d n_iter=%d compactness=%g)", F, H, W, n_iter, compactness);
        return 1;
    }
    if (slic_layout(F, H, W, n_segments, n_iter, L)) {
        set_error("crw_slic: every frame needs 1 <= n_segments and at most %d grid centres", SLIC_MAX_K);
        return 1;
    }
    if (ws_bytes < L.total) {
        set_error("crw_slic: workspace too small (%zu < %zu bytes)", ws_bytes, L.total);
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = static_cast<unsigned char*>(ws);
    unsigned* mm = reinterpret_cast<unsigned*>(w + L.mm);
    double* table = reinterpret_cast<double*>(w + L.table);
    double* feat = reinterpret_cast<double*>(w + L.feat);
    int* nearest = reinterpret_cast<int*>(w + L.nearest);
    int* queue = reinterpret_cast<int*>(w + L.queue);
    int* kinfo = reinterpret_cast<int*>(w + L.kinfo);
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(w + L.sums);
    const int hw = H * W;
    const size_t sums_per_iter = (size_t)F * L.Kmax * 6;
    cudaMemsetAsync(mm, 0, (size_t)F * 2 * sizeof(unsigned), st);
    cudaMemsetAsync(nearest, 0, (size_t)F * hw * sizeof(int), st);
    cudaMemsetAsync(sums, 0, (size_t)n_iter * sums_per_iter * sizeof(unsigned long long), st);
    if (connectivity) {
        cudaMemsetAsync(labels, 0xff, (size_t)F * hw * sizeof(int), st);
    }
    cudaMemcpyAsync(table, slic_lin_table(), 256 * sizeof(double), cudaMemcpyHostToDevice, st);
    int blocks = (hw * 3 + SLIC_THREADS * 8 - 1) / (SLIC_THREADS * 8);
    if (blocks > 64) blocks = 64;
    CRW_LAUNCH(slic_minmax_kernel, dim3(blocks, F), SLIC_THREADS, 0, st, video, (int64_t)3 * hw, mm);
    blocks = (hw + SLIC_THREADS - 1) / SLIC_THREADS;
    if (blocks > 256) blocks = 256;
    CRW_LAUNCH(slic_features_kernel, dim3(blocks, F), SLIC_THREADS, 0, st, video, hw, mm, table, 1.0 / compactness, feat);
    const size_t smem = (size_t)L.Kmax * (5 * sizeof(double) + 4 * sizeof(int) + 6 * sizeof(unsigned long long)) + 16;
#ifndef CRW_SIM
    if (smem > 48 * 1024) cudaFuncSetAttribute(slic_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    const int tiles = (hw + SLIC_THREADS * SLIC_PIX - 1) / (SLIC_THREADS * SLIC_PIX);
    for (int f0 = 0; f0 < F; f0 += SLIC_CHUNK) {
        const int nf = F - f0 < SLIC_CHUNK ? F - f0 : SLIC_CHUNK;
        SlicFrames fr;
        for (int i = 0; i < nf; ++i) slic_regular_grid(H, W, n_segments[f0 + i], fr.g[i], fr.sw[i]);
        for (int it = 0; it < n_iter; ++it) {
            const unsigned long long* prev = sums + (size_t)(it > 0 ? it - 1 : 0) * sums_per_iter;
            CRW_LAUNCH(slic_assign_kernel, dim3(tiles, nf), SLIC_THREADS, smem, st, fr, f0, H, W, it, L.Kmax, feat, prev,
                       sums + (size_t)it * sums_per_iter, nearest, kinfo);
        }
    }
    if (connectivity) CRW_LAUNCH(slic_connect_kernel, dim3(F), 32, 0, st, kinfo, 0, 0, H, W, nearest, queue, labels);
    else CRW_LAUNCH(slic_offset_kernel, dim3(296), 256, 0, st, nearest, (int64_t)F * hw, labels);
    return check_launch("crw_slic");
}

extern "C" size_t crw_label_connectivity_workspace_bytes(int F, int H, int W) {
    if (F < 1 || H < 1 || W < 1) return 0;
    return (size_t)F * H * W * sizeof(int);
}

extern "C" int crw_label_connectivity(const int* segments, int F, int H, int W, int min_size, int max_size, int* labels, void* ws,
                                      size_t ws_bytes, void* stream) {
    if (F < 1 || H < 1 || W < 1 || (int64_t)H * W > (1 << 30)) {
        set_error("crw_label_connectivity: bad arguments (F=%d H=%d W=%d)", F, H, W);
        return 1;
    }
    if (ws_bytes < crw_label_connectivity_workspace_bytes(F, H, W)) {
        set_error("crw_label_connectivity: workspace too small (%zu bytes)", ws_bytes);
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)F * H * W;
    int* queue = static_cast<int*>(ws);
    cudaMemsetAsync(labels, 0xff, n * sizeof(int), st);
    CRW_LAUNCH(slic_connect_kernel, dim3(F), 32, 0, st, (const int*)nullptr, min_size, max_size, H, W, segments, queue, labels);
    return check_launch("crw_label_connectivity");
}
