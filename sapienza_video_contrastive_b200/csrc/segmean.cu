// segmean.cu - superpixel node pooling (reference: code/model.py:296-325, CRW.image_to_nodes, which builds a
// (B,T,SP,h,w) one-hot, counts it per 8x8 window with utils.view_as_windows (utils/__init__.py:433-584) and
// multiplies a (B,T,H,W,C,SP) broadcast).  Mathematically that is a SEGMENT MEAN of the nearest-neighbour
// upsampled feature map over each superpixel label; here it is computed as a scatter-reduce
// straight from the label map:
//   1. segmean_count: one warp per feature cell turns its sy*sx labels into a short (label,count) list
//      (<= 64 entries, typically 1-4) and adds the counts to the per-label sizes;
//   2. segmean_csr: one CTA per (clip, frame) turns the per-cell lists into a CSR sorted by (256-cell chunk, label, cell)
//      with one offset per (chunk, label);
//   3. segmean_accum: one CTA per (clip, frame, 32-channel tile) streams the feature tile through a double-buffered
//      cp.async pipeline and reduces it per label with the accumulators in registers; the epilogue divides by the sizes;
//   4. segmean_bwd mirrors 3 as a gather over the per-cell lists (deterministic).
// Neither the one-hot nor the broadcast product is ever materialised; forward and backward are deterministic.
#include <stdlib.h>

#include "tc_common.cuh"

namespace crw {

#ifdef CRW_SIM
struct SegMap { int unused; };
#define CRW_GRID_CONSTANT
#else
typedef CUtensorMap SegMap;
#define CRW_GRID_CONSTANT __grid_constant__
#endif

constexpr int kSegCT = 32;       // channels per CTA
constexpr int kSegMaxEnt = 64;   // max distinct labels per feature cell (= max sy*sx)

constexpr int kSegChunk = 256;   // cells per pipeline stage of the forward
constexpr int kSegCTF = 64;      // channels per CTA in the forward (two per lane)
constexpr int kSegLDC = kSegChunk + 1;  // staged tile is [channel][cell], row stride 257 floats: 32 channels of one cell sit in 32 banks

struct SegWs {
    int* size;               // (B*T, SP)
    unsigned char* nent;     // (B*T, cells)
    unsigned* ent;           // (B*T, cap, cells): label << 8 | count
    uint2* csr;              // (B*T, cap*cells): {count as float bits, cell % kSegChunk}, sorted by (chunk, label, cell)
    int* cstart;             // (B*T, nchunks*SP + 1): CSR offset of (chunk k, label s) at [k*SP + s]; the last entry is the total
    int cap, nchunks;
    size_t bytes;
};

__host__ __device__ inline SegWs seg_ws(void* base, int B, int T, int cells, int SP, int cap) {
    SegWs w;
    size_t o = 0;
    w.size = (int*)((char*)base + o);
    o += ((size_t)B * T * SP * sizeof(int) + 255) / 256 * 256;
    w.nent = (unsigned char*)base + o;
    o += ((size_t)B * T * cells + 255) / 256 * 256;
    w.ent = (unsigned*)((char*)base + o);
    o += ((size_t)B * T * cap * cells * sizeof(unsigned) + 255) / 256 * 256;
    w.csr = (uint2*)((char*)base + o);
    o += ((size_t)B * T * cap * cells * sizeof(uint2) + 255) / 256 * 256;
    w.nchunks = (cells + kSegChunk - 1) / kSegChunk;
    w.cstart = (int*)((char*)base + o);
    o += ((size_t)B * T * ((size_t)w.nchunks * SP + 1) * sizeof(int) + 255) / 256 * 256;
    w.cap = cap;
    w.bytes = o;
    return w;
}

__global__ void __launch_bounds__(256) segmean_count_kernel(const int64_t* __restrict__ labels, int64_t ls_b, int64_t ls_t,
                                                            int64_t ls_y, int64_t ls_x, int T, int Hm, int Wm, int sy, int sx,
                                                            int SP, SegWs ws, int64_t total_cells) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int cells = Hm * Wm, npix = sy * sx;
    const bool pow2 = (sx & (sx - 1)) == 0;
    const int sxs = 31 - __clz(sx);
    for (int64_t gc = warp; gc < total_cells; gc += nw) {
        const int64_t bt = gc / cells;
        const int cell = (int)(gc - bt * cells);
        const int b = (int)(bt / T), t = (int)(bt - (int64_t)b * T);
        const int cy = cell / Wm, cx = cell - cy * Wm;
        const int64_t* base = labels + b * ls_b + t * ls_t + (int64_t)(cy * sy) * ls_y + (int64_t)(cx * sx) * ls_x;
        // labels outside [0,SP) never match a one-hot plane (model.py:299-301): they are dropped here, and everything
        // below works on 32-bit labels
        int lab[2];
        bool pend[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pix = lane + 32 * h;
            lab[h] = -1;
            if (pix < npix) {
                const int py = pow2 ? pix >> sxs : pix / sx, px = pix - py * sx;
                const int64_t L = base[(int64_t)py * ls_y + (int64_t)px * ls_x];
                if (L >= 0 && L < SP) lab[h] = (int)L;
            }
            pend[h] = lab[h] >= 0;
        }
        int n = 0;
        for (;;) {
            const unsigned m0 = __ballot_sync(kFull, pend[0]);
            const unsigned m1 = __ballot_sync(kFull, pend[1]);
            if (!(m0 | m1)) break;
            // leader: lowest pending pixel
            const int L = m0 ? __shfl_sync(kFull, lab[0], __ffs((int)m0) - 1) : __shfl_sync(kFull, lab[1], __ffs((int)m1) - 1);
            const bool h0 = lab[0] == L, h1 = lab[1] == L;         // pending by construction: equal labels retire together
            const int cnt = __popc(__ballot_sync(kFull, h0)) + __popc(__ballot_sync(kFull, h1));
            if (h0) pend[0] = false;
            if (h1) pend[1] = false;
            if (lane == 0) {
                ws.ent[((int64_t)bt * ws.cap + n) * cells + cell] = ((unsigned)L << 8) | (unsigned)cnt;
                atomicAdd(ws.size + bt * SP + L, cnt);
            }
            ++n;
        }
        if (lane == 0) ws.nent[bt * cells + cell] = (unsigned char)n;
    }
}

// Thread-per-cell variant for unit-stride label rows (the common case: sp_mask[:, :, 0] of a contiguous mask): a warp's lanes are
// 32 neighbouring cells, so at every (row, pixel) step the warp reads one 8-byte label per lane out of the same few cache lines,
// and the cell's list is built in registers (run-length shortcut, four slots, an array in local memory beyond that).  Same
// lists in the same order as segmean_count_kernel (entries by first occurrence in pixel order).
__global__ void __launch_bounds__(256) segmean_count_cell_kernel(const int64_t* __restrict__ labels, int64_t ls_b, int64_t ls_t, int64_t ls_y,
                                                                 int T, int Hm, int Wm, int sy, int sx, int SP, SegWs ws, int64_t total_cells) {
    const int cells = Hm * Wm;
    for (int64_t gc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gc < total_cells; gc += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bt = gc / cells;
        const int cell = (int)(gc - bt * cells);
        const int b = (int)(bt / T), t = (int)(bt - (int64_t)b * T);
        const int cy = cell / Wm, cx = cell - cy * Wm;
        const int64_t* base = labels + b * ls_b + t * ls_t + (int64_t)(cy * sy) * ls_y + (int64_t)(cx * sx);
        int l0 = -1, l1 = -1, l2 = -1, l3 = -1, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        int xl[kSegMaxEnt], xc[kSegMaxEnt];            // entries beyond the four slots (rare: a cell on a junction of many segments)
        int nx = 0;
        auto take = [&](int64_t L64) {
            if (L64 < 0 || L64 >= SP) return;          // labels outside [0,SP) never match a one-hot plane (model.py:299-301)
            const int L = (int)L64;
            if (L == l0) ++c0;
            else if (L == l1) ++c1;
            else if (L == l2) ++c2;
            else if (L == l3) ++c3;
            else if (l0 < 0) { l0 = L; c0 = 1; }
            else if (l1 < 0) { l1 = L; c1 = 1; }
            else if (l2 < 0) { l2 = L; c2 = 1; }
            else if (l3 < 0) { l3 = L; c3 = 1; }
            else {
                int j = 0;
                while (j < nx && xl[j] != L) ++j;
                if (j == nx) { xl[nx] = L; xc[nx] = 0; ++nx; }
                ++xc[j];
            }
        };
        if (sx <= 8) {
            // a row of the cell is at most 8 labels: all of them are requested before any is looked at, and the next row is on its
            // way while this one is counted (the loads are what this kernel waits for when the label map comes from DRAM)
            int64_t cur[8], nxt[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
#pragma unroll
            for (int px = 0; px < 8; ++px) cur[px] = px < sx ? base[px] : -1;
            for (int py = 0; py < sy; ++py) {
                if (py + 1 < sy) {
                    const int64_t* row = base + (int64_t)(py + 1) * ls_y;
#pragma unroll
                    for (int px = 0; px < 8; ++px) nxt[px] = px < sx ? row[px] : -1;
                }
#pragma unroll
                for (int px = 0; px < 8; ++px) take(cur[px]);
#pragma unroll
                for (int px = 0; px < 8; ++px) cur[px] = nxt[px];
            }
        } else {
            for (int py = 0; py < sy; ++py) {
                const int64_t* row = base + (int64_t)py * ls_y;
                for (int px = 0; px < sx; ++px) take(row[px]);
            }
        }
        unsigned* e = ws.ent + (int64_t)bt * ws.cap * cells + cell;
        int* size = ws.size + bt * SP;
        int n = 0;
        if (l0 >= 0) { e[(int64_t)n++ * cells] = ((unsigned)l0 << 8) | (unsigned)c0; atomicAdd(size + l0, c0); }
        if (l1 >= 0) { e[(int64_t)n++ * cells] = ((unsigned)l1 << 8) | (unsigned)c1; atomicAdd(size + l1, c1); }
        if (l2 >= 0) { e[(int64_t)n++ * cells] = ((unsigned)l2 << 8) | (unsigned)c2; atomicAdd(size + l2, c2); }
        if (l3 >= 0) { e[(int64_t)n++ * cells] = ((unsigned)l3 << 8) | (unsigned)c3; atomicAdd(size + l3, c3); }
        for (int j = 0; j < nx; ++j) { e[(int64_t)n++ * cells] = ((unsigned)xl[j] << 8) | (unsigned)xc[j]; atomicAdd(size + xl[j], xc[j]); }
        ws.nent[bt * cells + cell] = (unsigned char)n;
    }
}

// ---- dilated superpixel masks (model.py:303-309, utils/__init__.py:590-608) -------------------------------
// The reference dilates every label's one-hot mask with a 51..55-pixel structuring element (a depthwise 55 x 55 convolution
// over T*SP channels, thresholded at > 0), so masks overlap and a pixel carries a SET of labels.  Here the element is
// described by its half-width per vertical offset, w(dy) (diamond: R - |dy|, disc: floor(sqrt(R^2 - dy^2)), cross: R at
// dy = 0 and 0 elsewhere), and a label s reaches pixel (y, x) iff some row y' = y + dy holds a run of s whose x-range,
// widened by w(dy), covers x.  Pass 1 turns every label row into its runs; pass 2 (a warp per feature cell) sweeps the rows
// the cell's pixels can see, builds the pixels' label sets as bitsets in shared memory and emits the cell's (label, count)
// list - the same lists the undilated path produces, so the CSR / accumulate / backward kernels are shared.
constexpr int kDilMaxR = 63;            // structuring element up to 127 x 127
constexpr unsigned kDilBadLabel = 0x3ffu;

struct DilArgs {
    const int64_t* labels;
    int64_t ls_b, ls_t, ls_y, ls_x;
    int T, Hm, Wm, sy, sx, h, w, SP, R;
    unsigned* runs;              // (B*T, h, w): label << 16 | x0 of every run, in x order (x1 = next x0 - 1)
    int* nruns;                  // (B*T, h)
    unsigned char hw[kDilMaxR + 1];
    int NW, cpc, nxc;            // pass 2 strips: 32-column words and cells per strip, strips per row of cells
};

// pass 1: one warp per label row
__global__ void __launch_bounds__(256) segdil_runs_kernel(DilArgs a, int64_t total_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < total_rows; row += nw) {
        const int64_t bt = row / a.h;
        const int y = (int)(row - bt * a.h);
        const int b = (int)(bt / a.T), t = (int)(bt - (int64_t)b * a.T);
        const int64_t* src = a.labels + b * a.ls_b + t * a.ls_t + (int64_t)y * a.ls_y;
        unsigned* out = a.runs + row * a.w;
        int n = 0;
        unsigned prev_last = 0xffffffffu;                   // label of the last pixel of the previous 32-pixel chunk
        for (int x0 = 0; x0 < a.w; x0 += 32) {
            const int x = x0 + lane;
            unsigned lab = 0xfffffffeu;
            if (x < a.w) {
                const int64_t L = src[(int64_t)x * a.ls_x];
                lab = (L >= 0 && L < a.SP) ? (unsigned)L : kDilBadLabel;
            }
            unsigned prev = __shfl_up_sync(kFull, lab, 1);
            if (lane == 0) prev = prev_last;
            const bool start = x < a.w && lab != prev;
            const unsigned m = __ballot_sync(kFull, start);
            if (start) out[n + __popc(m & ((1u << lane) - 1u))] = (lab << 16) | (unsigned)x;
            n += __popc(m);
            prev_last = __shfl_sync(kFull, lab, 31);
        }
        if (lane == 0) a.nruns[row] = n;
    }
}

// pass 2: one CTA per (clip, frame, row of cells, chunk of <= 256 pixel columns).  The dilated masks of that strip are built
// as BITMAPS in shared memory, cov[pixel row][32-column word][label]: a half-warp takes one label row the strip can see,
// each lane one run of it, and ORs the run's widened interval into the bitmap of every pixel row it reaches (the work is
// shared by all the cells of the strip instead of being repeated per cell).  A warp then turns one cell into its
// (label, count) list: lanes own labels, a count is the popcount of the cell's columns over its pixel rows, and ballots
// compact the list in label order - the same lists the undilated kernel writes.
__global__ void __launch_bounds__(256) segdil_count_kernel(DilArgs a, SegWs ws) {
    CRW_DYN_SMEM(smem_raw);
    unsigned* cov = reinterpret_cast<unsigned*>(smem_raw);              // [sy][NW][SP]
    __shared__ int shw[kDilMaxR + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int blk = blockIdx.x;
    const int xc = blk % a.nxc;
    blk /= a.nxc;
    const int cy = blk % a.Hm, bt = blk / a.Hm;
    const int cells = a.Hm * a.Wm;
    const int cx_lo = xc * a.cpc, ncell = min(a.cpc, a.Wm - cx_lo);
    const int X0 = cx_lo * a.sx, X1 = X0 + ncell * a.sx - 1, py0 = cy * a.sy;
    for (int i = tid; i < a.sy * a.NW * a.SP; i += 256) cov[i] = 0u;
    if (tid <= kDilMaxR) shw[tid] = a.hw[tid];
    __syncthreads();
    const int ylo = max(0, py0 - a.R), yhi = min(a.h - 1, py0 + a.sy - 1 + a.R);
    for (int yy = ylo + (tid >> 4); yy <= yhi; yy += 16) {
        const int wmax = shw[max(0, max(py0 - yy, yy - (py0 + a.sy - 1)))];     // the widest reach of this row into the strip
        const int64_t row = (int64_t)bt * a.h + yy;
        const unsigned* rr = a.runs + row * a.w;
        const int nr = __ldg(a.nruns + row);
        for (int i = tid & 15; i < nr; i += 16) {
            const unsigned rec = __ldg(rr + i);
            const unsigned lab = rec >> 16;
            const int x0 = (int)(rec & 0xffffu);
            const int x1 = i + 1 < nr ? (int)(__ldg(rr + i + 1) & 0xffffu) - 1 : a.w - 1;
            if (lab == kDilBadLabel || x1 + wmax < X0 || x0 - wmax > X1) continue;
            for (int ly = 0; ly < a.sy; ++ly) {
                const int dy = abs(yy - (py0 + ly));
                if (dy > a.R) continue;
                const int wv = shw[dy];
                const int lo = max(x0 - wv, X0) - X0, hi = min(x1 + wv, X1) - X0;
                unsigned* base = cov + (size_t)ly * a.NW * a.SP + lab;
                for (int wd = lo >> 5; wd <= hi >> 5; ++wd) {          // empty when lo > hi
                    const int b0 = max(lo, wd * 32) - wd * 32, b1 = min(hi, wd * 32 + 31) - wd * 32;
                    if (b0 <= b1) atomicOr(base + (size_t)wd * a.SP, (0xffffffffu >> (31 - (b1 - b0))) << b0);
                }
            }
        }
    }
    __syncthreads();
    const unsigned colmask = a.sx == 32 ? 0xffffffffu : (1u << a.sx) - 1u;
    for (int c = warp; c < ncell; c += 8) {
        const int cell = cy * a.Wm + cx_lo + c;
        const int bit = c * a.sx, wd = bit >> 5, off = bit & 31;
        const bool two = off + a.sx > 32;                               // the cell's columns straddle two words
        int n = 0;
        for (int s0 = 0; s0 < a.SP; s0 += 32) {
            const int lab = s0 + lane;
            int cnt = 0;
            if (lab < a.SP)
                for (int ly = 0; ly < a.sy; ++ly) {
                    const unsigned* q = cov + ((size_t)ly * a.NW + wd) * a.SP + lab;
                    cnt += __popc(__funnelshift_r(q[0], two ? q[a.SP] : 0u, off) & colmask);
                }
            const unsigned have = __ballot_sync(kFull, cnt > 0);
            if (cnt > 0) {
                const int pos = n + __popc(have & ((1u << lane) - 1u));
                ws.ent[((int64_t)bt * ws.cap + pos) * cells + cell] = ((unsigned)lab << 8) | (unsigned)cnt;
                atomicAdd(ws.size + bt * a.SP + lab, cnt);
            }
            n += __popc(have);
        }
        if (lane == 0) ws.nent[(int64_t)bt * cells + cell] = (unsigned char)n;
    }
}

// CSR of one (clip, frame), one CTA: entries ordered by (chunk of kSegChunk cells, label, cell).  A per-label bitmask over
// the cells (shared memory) gives every (cell, label) entry its rank inside its (chunk, label) group by popcount, so the
// order needs no sort and no atomics on the positions: deterministic.  Needs cells <= 1024 (32 mask words per label).
__global__ void __launch_bounds__(1024) segmean_csr_kernel(SegWs ws, int cells, int SP) {
    CRW_DYN_SMEM(smem_raw);
    constexpr int wpc = kSegChunk / 32;                               // mask words per chunk
    unsigned* mask = reinterpret_cast<unsigned*>(smem_raw);          // [SP][32]
    int* P = reinterpret_cast<int*>(mask + (size_t)SP * 32);         // [nchunks * SP + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bt = blockIdx.x;
    const int L = ws.nchunks * SP;
    const unsigned char* gnent = ws.nent + (int64_t)bt * cells;
    const unsigned* gent = ws.ent + (int64_t)bt * ws.cap * cells;
    for (int e = tid; e < SP * 32; e += 1024) mask[e] = 0u;
    __syncthreads();
    for (int cell = tid; cell < cells; cell += 1024) {
        const int ne = gnent[cell];
        for (int sl = 0; sl < ne; ++sl) {
            const unsigned e = gent[(int64_t)sl * cells + cell];
            atomicOr(mask + (size_t)(e >> 8) * 32 + (cell >> 5), 1u << (cell & 31));
        }
    }
    __syncthreads();
    // group sizes
    for (int g = tid; g < L; g += 1024) {
        const int k = g / SP, sidx = g - k * SP;
        int n = 0;
#pragma unroll
        for (int i = 0; i < wpc; ++i) n += __popc(mask[(size_t)sidx * 32 + k * wpc + i]);
        P[g + 1] = n;
    }
    if (tid == 0) P[0] = 0;
    __syncthreads();
    if (warp == 0) {                                                   // inclusive scan, 32 groups at a time
        int carry = 0;
        for (int g0 = 0; g0 < L; g0 += 32) {
            int v = g0 + lane < L ? P[g0 + lane + 1] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(kFull, v, d);
                if (lane >= d) v += u;
            }
            if (g0 + lane < L) P[g0 + lane + 1] = v + carry;
            carry += __shfl_sync(kFull, v, 31);
        }
    }
    __syncthreads();
    int* gP = ws.cstart + (int64_t)bt * (L + 1);
    for (int g = tid; g <= L; g += 1024) gP[g] = P[g];
    // scatter the entries to their ranks
    uint2* csr = ws.csr + (int64_t)bt * ws.cap * cells;
    for (int cell = tid; cell < cells; cell += 1024) {
        const int ne = gnent[cell];
        const int w = cell >> 5, k = cell / kSegChunk;
        for (int sl = 0; sl < ne; ++sl) {
            const unsigned e = gent[(int64_t)sl * cells + cell];
            const unsigned* m = mask + (size_t)(e >> 8) * 32;
            int pos = P[k * SP + (int)(e >> 8)] + __popc(m[w] & ((1u << (cell & 31)) - 1u));
            for (int i = k * wpc; i < w; ++i) pos += __popc(m[i]);
            csr[pos] = make_uint2(__float_as_uint((float)(e & 255u)), (unsigned)(cell - k * kSegChunk));
        }
    }
}

__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
#ifdef CRW_SIM
    *dst = *src;
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef CRW_SIM
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef CRW_SIM
    asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
#endif
}

// Forward accumulation: persistent CTAs (one per SM, 32 warps) walk contiguous runs of (clip, frame, 64-channel tile) items.
//   * a lane owns channels l and l + 32 of the tile; warp w owns labels w, w + 32, ... and keeps their running sums in
//     registers (LMAX float2 per lane), so nothing is accumulated in shared memory and there are no atomics: deterministic;
//   * the features stream in 256-cell chunks: TMA bulk copies (one 1 KB row per channel, two chunks = 128 KB in flight per
//     SM, completing on mbarriers) land in dense staging buffers; the CTA re-lays a landed chunk out as [channel][257]
//     (128-bit reads, transposing 4-byte stores, both conflict-free) so that the 32 channels of one cell sit in 32
//     different banks; every feature is read from HBM exactly once;
//   * per chunk a warp walks, for each of its labels, the label's CSR entries (cell order): one broadcast 8-byte load of
//     the entry, two conflict-free loads of the features and one packed FMA per (cell, label) pair and 64 channels.
// Shapes the bulk copies cannot take (cells % 256, C % 64, unaligned base) use 4-byte cp.async copies instead.
constexpr int kSegEntCap = 1024;     // CSR entries of a chunk staged in shared memory (a fuller chunk is read from L2)
constexpr int kSegTileWords = (kSegCTF * kSegLDC + 3) & ~3;

__host__ __device__ inline size_t seg_meta_words(int SP) { return 2 * kSegEntCap + (size_t)((SP + 1 + 3) & ~3); }

// (chunk, tile, frame, clip) cursor advanced incrementally: no integer division inside the pipeline
struct SegCursor {
    int k, tile, b, t;
};
__device__ __forceinline__ void seg_advance(SegCursor& c, int nchunks, int ntiles, int T) {
    if (++c.k == nchunks) {
        c.k = 0;
        if (++c.tile == ntiles) {
            c.tile = 0;
            if (++c.t == T) { c.t = 0; ++c.b; }
        }
    }
}

// CSR entries + per-label offsets of one chunk -> shared memory (4-byte cp.async, one commit group)
__device__ __forceinline__ int2 seg_entry_range(const SegWs& ws, const SegCursor& q, int T, int SP) {      // (first entry, count) of a chunk
    const int* pk = ws.cstart + (int64_t)(q.b * T + q.t) * ((int64_t)ws.nchunks * SP + 1) + (int64_t)q.k * SP;
    const int eb = __ldg(pk);
    return make_int2(eb, __ldg(pk + SP) - eb);
}
template <int NT>
__device__ __forceinline__ void seg_issue_meta(const SegWs& ws, const SegCursor& q, int2 range, int T, int cells, int SP, uint2* est, int* pst, int tid) {
    const int bt = q.b * T + q.t;
    const int* pk = ws.cstart + (int64_t)bt * ((int64_t)ws.nchunks * SP + 1) + (int64_t)q.k * SP;
    for (int i = tid; i <= SP; i += NT) cp_async4(reinterpret_cast<float*>(pst + i), reinterpret_cast<const float*>(pk + i));
    const int eb = range.x, ne = range.y;
    if (ne <= kSegEntCap) {
        const float* esrc = reinterpret_cast<const float*>(ws.csr + (int64_t)bt * ws.cap * cells + eb);
        for (int i = tid; i < 2 * ne; i += NT) cp_async4(reinterpret_cast<float*>(est) + i, esrc + i);
    }
    cp_async_commit();
}

// A chunk with more CSR entries than the staging buffer holds (dilated masks: a cell carries ten and more labels).  The
// entries stay in L2; a warp fetches 32 of a label with one coalesced load - the next label's first batch while the current
// label is reduced - and hands them round by shuffles.  Same summation order as the staged path.
template <int LMAX>
__device__ __forceinline__ void seg_reduce_chunk_l2(float2 (&acc)[LMAX], const float* tl0, const float* tl1, const int* pst,
                                                    const uint2* csr_bt, int SP, int warp, int lane) {
    uint2 nxt = make_uint2(0u, 0u);
    if (warp < SP && pst[warp] + lane < pst[warp + 1]) nxt = csr_bt[pst[warp] + lane];
#pragma unroll
    for (int i = 0; i < LMAX; ++i) {
        const int s = warp + 32 * i;
        if (s < SP) {
            const int e0 = pst[s], e1 = pst[s + 1];
            uint2 cur = nxt;
            if (i + 1 < LMAX && s + 32 < SP) {
                const int f0 = pst[s + 32] + lane;
                nxt = f0 < pst[s + 33] ? csr_bt[f0] : make_uint2(0u, 0u);
            }
            float2 a = acc[i];
            for (int base = e0; base < e1; base += 32) {
                if (base != e0) cur = base + lane < e1 ? csr_bt[base + lane] : make_uint2(0u, 0u);
                const int n = min(32, e1 - base);
                for (int j = 0; j < n; ++j) {
                    const float wgt = __uint_as_float(__shfl_sync(kFull, cur.x, j));
                    const unsigned cl = __shfl_sync(kFull, cur.y, j);
                    a = ffma2(wgt, make_float2(tl0[cl], tl1[cl]), a);
                }
            }
            acc[i] = a;
        }
    }
}

template <int LMAX>
__device__ __forceinline__ void seg_reduce_chunk(float2 (&acc)[LMAX], const float* tile, const uint2* est, const int* pst,
                                                 const uint2* csr_bt, int SP, int warp, int lane) {
    const float* tl0 = tile + lane * kSegLDC;
    const float* tl1 = tl0 + 32 * kSegLDC;
    const int eb = pst[0];
    if (pst[SP] - eb > kSegEntCap) {
        seg_reduce_chunk_l2<LMAX>(acc, tl0, tl1, pst, csr_bt, SP, warp, lane);
        return;
    }
    const uint2* ents = est - eb;                                                   // indexed by the CSR position
#pragma unroll
    for (int i = 0; i < LMAX; ++i) {
        const int s = warp + 32 * i;
        if (s < SP) {
            const int e0 = pst[s], e1 = pst[s + 1];
            float2 a = acc[i];
#pragma unroll 8
            for (int e = e0; e < e1; ++e) {
                const uint2 en = ents[e];
                a = ffma2(__uint_as_float(en.x), make_float2(tl0[en.y], tl1[en.y]), a);
            }
            acc[i] = a;
        }
    }
}

template <int LMAX>
__device__ __forceinline__ void seg_load_den(float (&den)[LMAX], const SegWs& ws, const SegCursor& q, int T, int SP, int warp) {
    const int* size = ws.size + (int64_t)(q.b * T + q.t) * SP;
#pragma unroll
    for (int i = 0; i < LMAX; ++i) den[i] = warp + 32 * i < SP ? (float)__ldg(size + warp + 32 * i) + kEpsLog : 1.f;
}

// x / d for a positive normal d through one reciprocal per label and one FMA correction per value: the quotient the
// division operator gives (its fast path is the same sequence) at a fraction of the instructions - the epilogue of
// 32 warps is issue-bound otherwise
__device__ __forceinline__ float seg_div(float x, float d, float r) {
    const float q = x * r;
    return fmaf(fmaf(-q, d, x), r, q);
}

template <int LMAX>
__device__ __forceinline__ void seg_write_out(const float2 (&acc)[LMAX], const float (&den)[LMAX], const SegCursor& q, int C, int T, int SP,
                                              float* __restrict__ out, int warp, int lane) {
    const int c = q.tile * kSegCTF + lane;
#pragma unroll
    for (int i = 0; i < LMAX; ++i) {
        const int s = warp + 32 * i;
        if (s < SP) {
            const float d = den[i];
            float r = 1.0f / d;                                    // d >= 1, or d = 1e-20 for an empty label whose sums are 0
            r = fmaf(fmaf(-d, r, 1.0f), r, r);
            float* o = out + (((int64_t)q.b * SP + s) * T + q.t) * C + c;
            if (c < C) o[0] = seg_div(acc[i].x, d, r);
            if (c + 32 < C) o[32] = seg_div(acc[i].y, d, r);
        }
    }
}

template <int LMAX>
__device__ __forceinline__ void seg_first_cursor(SegCursor& cur, int& total, int BT, int C, int T, int nchunks) {
    const int ntiles = (C + kSegCTF - 1) / kSegCTF;
    const int nitems = BT * ntiles;                              // item = bt * ntiles + tile: neighbours share their CSR in L2
    const int item_lo = (int)(((int64_t)nitems * blockIdx.x) / gridDim.x), item_hi = (int)(((int64_t)nitems * (blockIdx.x + 1)) / gridDim.x);
    total = (item_hi - item_lo) * nchunks;                      // this CTA's contiguous run of items, chunk by chunk
    const int bt0 = item_lo / ntiles;
    cur.k = 0; cur.tile = item_lo - bt0 * ntiles; cur.b = bt0 / T; cur.t = bt0 - cur.b * T;
}

// ---- TMA path: cells % 256 == 0, C % 64 == 0, 16-byte aligned maps ---------------------------------------------------
// shared memory: staging[2][64][256] | tile[64][257] | meta[2] (entries, offsets) | 2 mbarriers
template <int LMAX>
__global__ void __launch_bounds__(1024, 1) segmean_accum_tma_kernel(const CRW_GRID_CONSTANT SegMap fmap, const float* __restrict__ maps, SegWs ws,
                                                                    int C, int T, int BT, int cells, int SP, float* __restrict__ out) {
    CRW_DYN_SMEM(smem_raw);
    constexpr int NT = 1024, kStageWords = kSegCTF * kSegChunk;
    float* staging = reinterpret_cast<float*>(smem_raw);
    float* tile = staging + 2 * kStageWords;
    float* meta = tile + kSegTileWords;
    const size_t meta_words = seg_meta_words(SP);
    uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 2 * meta_words);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = ws.nchunks, ntiles = C / kSegCTF;
    [[maybe_unused]] const int64_t cstride = (int64_t)T * cells;      // the host simulator's stand-in for the tensor copy
    SegCursor cur;
    int total;
    seg_first_cursor<LMAX>(cur, total, BT, C, T, nchunks);
    SegCursor nxt = cur;            // next chunk whose features get issued
    SegCursor mnx = cur;            // next chunk whose entries / offsets get issued

    bulk_bar_init(bars, tid);
    bulk_bar_init(bars + 1, tid);

    // one TMA tensor copy per chunk: box (256 cells, 1 frame, 64 channels) of the (cells, T, B*C) view of the maps, issued by
    // one thread (every thread copies in the host simulator)
    auto issue_rows = [&](const SegCursor& q, int s) {
        float* dst = staging + (size_t)s * kStageWords;
#ifdef CRW_SIM
        const float* src = maps + (((int64_t)q.b * C + q.tile * kSegCTF) * T + q.t) * cells + q.k * kSegChunk;
        for (int i = tid; i < kStageWords; i += NT) dst[i] = src[(int64_t)(i / kSegChunk) * cstride + (i % kSegChunk)];
#else
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was just read through the generic proxy
            mbar_expect_tx(bars + s, (unsigned)(kStageWords * 4));
            tma_load_3d(dst, &fmap, q.k * kSegChunk, q.t, q.b * C + q.tile * kSegCTF, bars + s);
        }
#endif
    };
    auto meta_est = [&](int s) { return reinterpret_cast<uint2*>(meta + (size_t)s * meta_words); };
    auto meta_pst = [&](int s) { return reinterpret_cast<int*>(meta_est(s) + kSegEntCap); };

    float2 acc[LMAX];
    float den[LMAX];
    int issued = 0;
    if (total > 0) {
        issue_rows(nxt, 0);
        issued = 1;
        seg_issue_meta<NT>(ws, mnx, seg_entry_range(ws, mnx, T, SP), T, cells, SP, meta_est(0), meta_pst(0), tid);
        if (total > 1) {
            seg_advance(nxt, nchunks, ntiles, T);
            issue_rows(nxt, 1);
            issued = 2;
        }
    }
    int2 range1 = make_int2(0, 0);                                   // entry range of chunk g + 1, fetched one iteration early
    if (total > 1) { seg_advance(mnx, nchunks, ntiles, T); range1 = seg_entry_range(ws, mnx, T, SP); }
    for (int g = 0; g < total; ++g, seg_advance(cur, nchunks, ntiles, T)) {
        const int s = g & 1;
        // entries / offsets of chunk g + 1 (their buffer was last read in iteration g - 1)
        if (g + 1 < total) {
            seg_issue_meta<NT>(ws, mnx, range1, T, cells, SP, meta_est(s ^ 1), meta_pst(s ^ 1), tid);
            if (g + 2 < total) { seg_advance(mnx, nchunks, ntiles, T); range1 = seg_entry_range(ws, mnx, T, SP); }
            cp_async_wait<1>();
        } else cp_async_wait<0>();
        bulk_wait(bars + s, (unsigned)((g >> 1) & 1));
        // dense [64][256] -> padded [64][257]: a warp instruction covers 4 channels x 32 cells
        {
            const float* src = staging + (size_t)s * kStageWords;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = warp + 32 * u, ch = (p >> 3) * 4 + (lane >> 3), j = (p & 7) * 32 + 4 * (lane & 7);
                const float4 v = *reinterpret_cast<const float4*>(src + ch * kSegChunk + j);
                float* d = tile + ch * kSegLDC + j;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        }
        __syncthreads();
        if (issued < total) {                                        // staging[s] is free again: fetch chunk g + 2 into it
            seg_advance(nxt, nchunks, ntiles, T);
            issue_rows(nxt, s);
            ++issued;
        }
        if (cur.k == 0) {
#pragma unroll
            for (int i = 0; i < LMAX; ++i) acc[i] = make_float2(0.f, 0.f);
            seg_load_den<LMAX>(den, ws, cur, T, SP, warp);           // the divisors arrive while the item is reduced
        }
        seg_reduce_chunk<LMAX>(acc, tile, meta_est(s), meta_pst(s), ws.csr + (int64_t)(cur.b * T + cur.t) * ws.cap * cells, SP, warp, lane);
        if (cur.k == nchunks - 1) seg_write_out<LMAX>(acc, den, cur, C, T, SP, out, warp, lane);
        __syncthreads();
    }
}

// ---- generic path: two-stage pipeline of 4-byte cp.async copies -----------------------------------------------------
// shared memory: 2 x (tile[64][257] | meta)
template <int LMAX>
__global__ void __launch_bounds__(1024, 1) segmean_accum_kernel(const float* __restrict__ maps, SegWs ws, int C, int T, int BT, int cells,
                                                                int SP, float* __restrict__ out) {
    CRW_DYN_SMEM(smem_raw);
    constexpr int NT = 1024, CSTEP = NT / kSegChunk;
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    const size_t stage_words = kSegTileWords + seg_meta_words(SP);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = ws.nchunks, ntiles = (C + kSegCTF - 1) / kSegCTF;
    const int64_t cstride = (int64_t)T * cells;
    SegCursor cur;
    int total;
    seg_first_cursor<LMAX>(cur, total, BT, C, T, nchunks);
    SegCursor nxt = cur;

    auto issue = [&](const SegCursor& q, int buf) {
        float* tile = stage0 + (size_t)buf * stage_words;
        uint2* est = reinterpret_cast<uint2*>(tile + kSegTileWords);
        const int c0 = q.tile * kSegCTF;
        {   // thread = one cell of the chunk, stepping over the channels: coalesced 128-byte reads along the cells
            const int j = tid & (kSegChunk - 1), cell = q.k * kSegChunk + j, cl0 = tid / kSegChunk;
            float* d = tile + cl0 * kSegLDC + j;
            const float* gsrc = maps + (((int64_t)q.b * C + c0 + cl0) * T + q.t) * cells + cell;
            for (int cl = cl0; cl < kSegCTF; cl += CSTEP, d += CSTEP * kSegLDC, gsrc += CSTEP * cstride) {
                if (cell < cells && c0 + cl < C) cp_async4(d, gsrc);
                else *d = 0.f;
            }
        }
        seg_issue_meta<NT>(ws, q, seg_entry_range(ws, q, T, SP), T, cells, SP, est, reinterpret_cast<int*>(est + kSegEntCap), tid);     // commits the group
    };

    float2 acc[LMAX];
    float den[LMAX];
    if (total > 0) issue(cur, 0);
    for (int g = 0; g < total; ++g, seg_advance(cur, nchunks, ntiles, T)) {
        if (g + 1 < total) {
            seg_advance(nxt, nchunks, ntiles, T);
            issue(nxt, (g + 1) & 1);
            cp_async_wait<1>();
        } else cp_async_wait<0>();
        __syncthreads();
        if (cur.k == 0) {
#pragma unroll
            for (int i = 0; i < LMAX; ++i) acc[i] = make_float2(0.f, 0.f);
            seg_load_den<LMAX>(den, ws, cur, T, SP, warp);
        }
        const float* tile = stage0 + (size_t)(g & 1) * stage_words;
        const uint2* est = reinterpret_cast<const uint2*>(tile + kSegTileWords);
        seg_reduce_chunk<LMAX>(acc, tile, est, reinterpret_cast<const int*>(est + kSegEntCap),
                               ws.csr + (int64_t)(cur.b * T + cur.t) * ws.cap * cells, SP, warp, lane);
        if (cur.k == nchunks - 1) seg_write_out<LMAX>(acc, den, cur, C, T, SP, out, warp, lane);
        __syncthreads();
    }
}

// Backward, one CTA per (clip, frame, 32-channel tile): gout / size staged in shared memory as [label][36]; a thread owns
// a cell, keeps the cell's first four (label, count) entries in registers and produces the cell's 32 channels with 128-bit
// shared-memory reads (equal labels broadcast, different labels mostly sit in different bank groups), so a warp writes 128
// contiguous bytes per channel.  A gather: deterministic.
constexpr int kSegLDB = kSegCT + 4;

// kCrowded (dilated masks, ten and more labels per cell): the cell's 32 channels are summed in registers over all its
// entries and written once, instead of one read-modify-write round of the output per four entries.
template <bool kCrowded>
__global__ void __launch_bounds__(256) segmean_bwd_kernel(const float* __restrict__ gout, SegWs ws, int C, int T, int cells,
                                                          int SP, float* __restrict__ gmaps) {
    CRW_DYN_SMEM(smem_raw);
    float* wg = reinterpret_cast<float*>(smem_raw);             // [SP][kSegLDB] = gout / size
    float* dn = wg + (size_t)SP * kSegLDB;                      // [SP] size + eps
    float* rc = dn + SP;                                        // [SP] its reciprocal
    const int tid = threadIdx.x;
    const int bt = blockIdx.y, b = bt / T, t = bt - b * T;
    const int c0 = blockIdx.x * kSegCT;
    const int* size = ws.size + (int64_t)bt * SP;
    for (int s = tid; s < SP; s += 256) {
        const float d = (float)size[s] + kEpsLog;
        float r = 1.0f / d;
        r = fmaf(fmaf(-d, r, 1.0f), r, r);
        dn[s] = d;
        rc[s] = r;
    }
    __syncthreads();
    const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(gout) & 15) == 0 && c0 + kSegCT <= C;
    if (vec) {      // 8 threads x 16 bytes cover one label's 32 channels
        for (int e = tid; e < SP * (kSegCT / 4); e += 256) {
            const int s = e >> 3, c4 = e & 7;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gout + (((int64_t)b * SP + s) * T + t) * C + c0) + c4);
            const float d = dn[s], r = rc[s];
            *reinterpret_cast<float4*>(wg + s * kSegLDB + 4 * c4) = make_float4(seg_div(g.x, d, r), seg_div(g.y, d, r), seg_div(g.z, d, r), seg_div(g.w, d, r));
        }
    } else {
        for (int e = tid; e < SP * kSegCT; e += 256) {
            const int s = e / kSegCT, cl = e - s * kSegCT;
            wg[s * kSegLDB + cl] = (c0 + cl < C) ? seg_div(__ldg(gout + (((int64_t)b * SP + s) * T + t) * C + c0 + cl), dn[s], rc[s]) : 0.f;
        }
    }
    __syncthreads();
    const unsigned char* nent = ws.nent + (int64_t)bt * cells;
    const unsigned* ent = ws.ent + (int64_t)bt * ws.cap * cells;
    const int nc = min(kSegCT, C - c0);
    float* dst0 = gmaps + (((int64_t)b * C + c0) * T + t) * cells;
    const int64_t cstride = (int64_t)T * cells;
    for (int cell0 = 0; cell0 < cells; cell0 += 256) {                        // warp-uniform trip count
        const int cell = cell0 + tid;
        const bool live = cell < cells;
        const int ne = live ? nent[cell] : 0;
        int nmax = ne;                                                        // slots are taken four at a time, for as long as any lane of the warp has some
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, o));
        if constexpr (kCrowded) {
            float acc[kSegCT];
#pragma unroll
            for (int c = 0; c < kSegCT; ++c) acc[c] = 0.f;
            for (int base = 0; base < nmax; base += 4) {
                float cnt[4];
                const float4* row[4];
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) {
                    const unsigned e = base + sl < ne ? __ldg(ent + (int64_t)(base + sl) * cells + cell) : 0u;
                    cnt[sl] = (float)(e & 255u);
                    row[sl] = reinterpret_cast<const float4*>(wg + (e >> 8) * kSegLDB);
                }
#pragma unroll
                for (int c4 = 0; c4 < kSegCT / 4; ++c4) {
                    const float4 v0 = row[0][c4], v1 = row[1][c4], v2 = row[2][c4], v3 = row[3][c4];
                    acc[4 * c4 + 0] = fmaf(cnt[3], v3.x, fmaf(cnt[2], v2.x, fmaf(cnt[1], v1.x, fmaf(cnt[0], v0.x, acc[4 * c4 + 0]))));
                    acc[4 * c4 + 1] = fmaf(cnt[3], v3.y, fmaf(cnt[2], v2.y, fmaf(cnt[1], v1.y, fmaf(cnt[0], v0.y, acc[4 * c4 + 1]))));
                    acc[4 * c4 + 2] = fmaf(cnt[3], v3.z, fmaf(cnt[2], v2.z, fmaf(cnt[1], v1.z, fmaf(cnt[0], v0.z, acc[4 * c4 + 2]))));
                    acc[4 * c4 + 3] = fmaf(cnt[3], v3.w, fmaf(cnt[2], v2.w, fmaf(cnt[1], v1.w, fmaf(cnt[0], v0.w, acc[4 * c4 + 3]))));
                }
            }
            if (live) {
                float* d = dst0 + cell;
#pragma unroll
                for (int c = 0; c < kSegCT; ++c)
                    if (c < nc) d[c * cstride] = acc[c];
            }
        } else
        for (int base = 0; base < nmax || base == 0; base += 4) {
            float cnt[4];
            const float4* row[4];
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
                const unsigned e = base + sl < ne ? __ldg(ent + (int64_t)(base + sl) * cells + cell) : 0u;
                cnt[sl] = (float)(e & 255u);                      // 0 for the unused slots: they add wg[0][c] * 0
                row[sl] = reinterpret_cast<const float4*>(wg + (e >> 8) * kSegLDB);
            }
            if (!live || (base > 0 && base >= ne)) continue;
            float* d = dst0 + cell;
#pragma unroll
            for (int c4 = 0; c4 < kSegCT / 4; ++c4) {
                const float4 v0 = row[0][c4], v1 = row[1][c4], v2 = row[2][c4], v3 = row[3][c4];
                float g[4] = {fmaf(cnt[3], v3.x, fmaf(cnt[2], v2.x, fmaf(cnt[1], v1.x, cnt[0] * v0.x))),
                              fmaf(cnt[3], v3.y, fmaf(cnt[2], v2.y, fmaf(cnt[1], v1.y, cnt[0] * v0.y))),
                              fmaf(cnt[3], v3.z, fmaf(cnt[2], v2.z, fmaf(cnt[1], v1.z, cnt[0] * v0.z))),
                              fmaf(cnt[3], v3.w, fmaf(cnt[2], v2.w, fmaf(cnt[1], v1.w, cnt[0] * v0.w)))};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (4 * c4 + u < nc) {
                        float* o = d + (4 * c4 + u) * cstride;
                        *o = base == 0 ? g[u] : *o + g[u];        // cells with more than four labels (rare) take another round
                    }
                }
            }
        }
    }
}

static int seg_check(int B, int C, int T, int Hm, int Wm, int h, int w, int SP, int* cap) {
    if (B <= 0 || C <= 0 || T <= 0 || Hm <= 0 || Wm <= 0 || SP <= 0 || h % Hm || w % Wm) {
        set_error("segmean: bad shape B=%d C=%d T=%d maps=%dx%d labels=%dx%d SP=%d", B, C, T, Hm, Wm, h, w, SP);
        return CRW_ERR_SHAPE;
    }
    const int npix = (h / Hm) * (w / Wm);
    if (npix > kSegMaxEnt || SP >= (1 << 24) || (size_t)SP * (kSegCT + 6) * 4 > 227 * 1024) {
        set_error("segmean: unsupported scale %dx%d or SP=%d", h / Hm, w / Wm, SP);
        return CRW_ERR_UNSUPPORTED;
    }
    *cap = npix < SP ? npix : SP;
    return CRW_OK;
}

}  // namespace crw

using namespace crw;

extern "C" size_t crw_segmean_workspace_bytes(int B, int T, int Hm, int Wm, int h, int w, int SP) {
    int cap = 0;
    if (seg_check(B, 1, T, Hm, Wm, h, w, SP, &cap) != CRW_OK) return 0;
    return seg_ws(nullptr, B, T, Hm * Wm, SP, cap).bytes;
}

// everything after the per-cell (label, count) lists: CSR, then the accumulation
static int seg_fwd_tail(const float* maps, const SegWs& ws, int B, int C, int T, int cells, int SP, float* out, crw_stream_t stream);

extern "C" int crw_segmean_fwd(const float* maps, const int64_t* labels, int64_t ls_b, int64_t ls_t, int64_t ls_y, int64_t ls_x,
                               int B, int C, int T, int Hm, int Wm, int h, int w, int SP,
                               float* out, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    int cap = 0;
    int e = seg_check(B, C, T, Hm, Wm, h, w, SP, &cap);
    if (e != CRW_OK) return e;
    const int cells = Hm * Wm;
    SegWs ws = seg_ws(workspace, B, T, cells, SP, cap);
    if (!workspace || workspace_bytes < ws.bytes) { set_error("segmean_fwd: workspace too small"); return CRW_ERR_SHAPE; }
    cudaMemsetAsync(ws.size, 0, sizeof(int) * (size_t)B * T * SP, (cudaStream_t)stream);
    const int64_t total = (int64_t)B * T * cells;
    const int grid1 = (int)((total + 7) / 8 < 148 * 8 ? (total + 7) / 8 : 148 * 8);
    if (ls_x == 1) {
        const int gridc = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        CRW_LAUNCH(segmean_count_cell_kernel, gridc, 256, 0, stream, labels, ls_b, ls_t, ls_y, T, Hm, Wm, h / Hm, w / Wm, SP, ws, total);
    } else {
        CRW_LAUNCH(segmean_count_kernel, grid1, 256, 0, stream, labels, ls_b, ls_t, ls_y, ls_x, T, Hm, Wm, h / Hm, w / Wm, SP, ws, total);
    }
    e = check_launch("segmean_count");
    if (e != CRW_OK) return e;
    return seg_fwd_tail(maps, ws, B, C, T, cells, SP, out, stream);
}

#include "segmean_tc.cuh"

static int seg_fwd_tail(const float* maps, const SegWs& ws, int B, int C, int T, int cells, int SP, float* out, crw_stream_t stream) {
    int e;
#ifndef CRW_SIM
    static const bool force_simt = getenv("CRW_SEGMEAN_SIMT") != nullptr;     // A/B switch for tests and profiles
    if (!force_simt && seg_mma_eligible(maps, out, C, cells, SP))          // the tensor-core forward consumes the per-cell lists directly
        return C % 256 == 0 ? seg_mma_launch<256>(maps, ws, B, C, T, cells, SP, out, stream)
                            : seg_mma_launch<128>(maps, ws, B, C, T, cells, SP, out, stream);
#endif
    if (cells > 1024 || SP > 1024 || (size_t)SP * 33 * 4 > 200 * 1024) {
        set_error("segmean_fwd: unsupported size (cells=%d, SP=%d)", cells, SP);
        return CRW_ERR_UNSUPPORTED;
    }
    {
        const size_t smem = sizeof(unsigned) * (size_t)SP * 32 + sizeof(int) * ((size_t)ws.nchunks * SP + 1);
        auto k = segmean_csr_kernel;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CRW_LAUNCH(k, B * T, 1024, smem, stream, ws, cells, SP);
        e = check_launch("segmean_csr");
        if (e != CRW_OK) return e;
    }
    const bool tma = cells % kSegChunk == 0 && C % kSegCTF == 0 && (reinterpret_cast<uintptr_t>(maps) & 15) == 0;
    const size_t smem = tma ? sizeof(float) * (2 * (size_t)kSegCTF * kSegChunk + kSegTileWords + 2 * seg_meta_words(SP)) + 16
                            : sizeof(float) * 2 * (kSegTileWords + seg_meta_words(SP));
    SegMap fmap{};
#ifndef CRW_SIM
    if (tma) {   // the maps as a (cells, T, B*C) tensor; one box = 256 cells of one frame of 64 consecutive channels
        EncodeTiledFn enc = get_encode();
        cuuint64_t dims[3] = {(cuuint64_t)cells, (cuuint64_t)T, (cuuint64_t)B * C};
        cuuint64_t strides[2] = {(cuuint64_t)cells * 4, (cuuint64_t)T * cells * 4};
        cuuint32_t box[3] = {(cuuint32_t)kSegChunk, 1, (cuuint32_t)kSegCTF};
        cuuint32_t estr[3] = {1, 1, 1};
        if (!enc || enc(&fmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(maps), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            set_error("segmean_fwd: cuTensorMapEncodeTiled failed");
            return CRW_ERR_CUDA;
        }
    }
#endif
    const int nitems = B * T * ((C + kSegCTF - 1) / kSegCTF);
    const int grid = nitems < 148 ? nitems : 148;                 // persistent: one CTA per SM
#define CRW_SEG_LAUNCH(LMAX)                                                                              \
    do {                                                                                                  \
        if (tma) {                                                                                        \
            auto k = segmean_accum_tma_kernel<LMAX>;                                                      \
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
            CRW_LAUNCH(k, grid, 1024, smem, stream, fmap, maps, ws, C, T, B * T, cells, SP, out);         \
        } else {                                                                                          \
            auto k = segmean_accum_kernel<LMAX>;                                                          \
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
            CRW_LAUNCH(k, grid, 1024, smem, stream, maps, ws, C, T, B * T, cells, SP, out);               \
        }                                                                                                 \
    } while (0)
    if (SP <= 4 * 32) CRW_SEG_LAUNCH(4);
    else if (SP <= 7 * 32) CRW_SEG_LAUNCH(7);
    else if (SP <= 8 * 32) CRW_SEG_LAUNCH(8);
    else if (SP <= 16 * 32) CRW_SEG_LAUNCH(16);
    else CRW_SEG_LAUNCH(32);
#undef CRW_SEG_LAUNCH
    return check_launch("segmean_accum");
}

static int seg_bwd_run(const float* grad_out, const SegWs& ws, int B, int C, int T, int cells, int SP, float* grad_maps, bool crowded,
                       crw_stream_t stream);

extern "C" int crw_segmean_bwd(const float* grad_out, const void* workspace, size_t workspace_bytes,
                               int B, int C, int T, int Hm, int Wm, int h, int w, int SP,
                               float* grad_maps, crw_stream_t stream) {
    int cap = 0;
    int e = seg_check(B, C, T, Hm, Wm, h, w, SP, &cap);
    if (e != CRW_OK) return e;
    const int cells = Hm * Wm;
    SegWs ws = seg_ws(const_cast<void*>(workspace), B, T, cells, SP, cap);
    if (!workspace || workspace_bytes < ws.bytes) { set_error("segmean_bwd: workspace too small"); return CRW_ERR_SHAPE; }
    return seg_bwd_run(grad_out, ws, B, C, T, cells, SP, grad_maps, false, stream);
}

static int seg_bwd_run(const float* grad_out, const SegWs& ws, int B, int C, int T, int cells, int SP, float* grad_maps, bool crowded,
                       crw_stream_t stream) {
    const size_t smem = ((size_t)SP * kSegLDB + 2 * (size_t)SP) * sizeof(float);
    dim3 grid((C + kSegCT - 1) / kSegCT, B * T);
    if (crowded) {
        auto k = segmean_bwd_kernel<true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CRW_LAUNCH(k, grid, 256, smem, stream, grad_out, ws, C, T, cells, SP, grad_maps);
    } else {
        auto k = segmean_bwd_kernel<false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CRW_LAUNCH(k, grid, 256, smem, stream, grad_out, ws, C, T, cells, SP, grad_maps);
    }
    return check_launch("segmean_bwd");
}

// ---- dilated masks: C ABI ----------------------------------------------------------------------------------------------
// workspace = the plain layout with cap = SP (a cell can see every label), then the run records and per-row run counts
static int segdil_check(int B, int C, int T, int Hm, int Wm, int h, int w, int SP, int ksize, int shape) {
    int cap = 0;
    int e = seg_check(B, C, T, Hm, Wm, h, w, SP, &cap);
    if (e != CRW_OK) return e;
    if (ksize < 1 || !(ksize & 1) || shape < CRW_DILATE_L1 || shape > CRW_DILATE_CROSS) {
        set_error("segmean_dilated: kernel size %d must be odd and positive, shape %d one of CRW_DILATE_*", ksize, shape);
        return CRW_ERR_SHAPE;
    }
    if (SP > 255 || w > 65535 || ksize / 2 > kDilMaxR || w / Wm > 32) {
        set_error("segmean_dilated: unsupported SP=%d (<= 255), width=%d, cell width %d (<= 32) or kernel size %d (<= %d)", SP, w, w / Wm, ksize,
                  2 * kDilMaxR + 1);
        return CRW_ERR_UNSUPPORTED;
    }
    return CRW_OK;
}

static SegWs segdil_ws(void* base, int B, int T, int cells, int h, int w, int SP, unsigned** runs, int** nruns) {
    SegWs ws = seg_ws(base, B, T, cells, SP, SP);
    *runs = (unsigned*)((char*)base + ws.bytes);
    ws.bytes += ((size_t)B * T * h * w * sizeof(unsigned) + 255) / 256 * 256;
    *nruns = (int*)((char*)base + ws.bytes);
    ws.bytes += ((size_t)B * T * h * sizeof(int) + 255) / 256 * 256;
    return ws;
}

extern "C" size_t crw_segmean_dilated_workspace_bytes(int B, int T, int Hm, int Wm, int h, int w, int SP) {
    if (segdil_check(B, 1, T, Hm, Wm, h, w, SP, 1, CRW_DILATE_L1) != CRW_OK) return 0;
    unsigned* runs;
    int* nruns;
    return segdil_ws(nullptr, B, T, Hm * Wm, h, w, SP, &runs, &nruns).bytes;
}

extern "C" int crw_segmean_dilated_fwd(const float* maps, const int64_t* labels, int64_t ls_b, int64_t ls_t, int64_t ls_y, int64_t ls_x,
                                       int B, int C, int T, int Hm, int Wm, int h, int w, int SP, int ksize, int shape,
                                       float* out, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    int e = segdil_check(B, C, T, Hm, Wm, h, w, SP, ksize, shape);
    if (e != CRW_OK) return e;
    const int cells = Hm * Wm;
    DilArgs a{};
    SegWs ws = segdil_ws(workspace, B, T, cells, h, w, SP, &a.runs, &a.nruns);
    if (!workspace || workspace_bytes < ws.bytes) { set_error("segmean_dilated_fwd: workspace too small"); return CRW_ERR_SHAPE; }
    a.labels = labels; a.ls_b = ls_b; a.ls_t = ls_t; a.ls_y = ls_y; a.ls_x = ls_x;
    a.T = T; a.Hm = Hm; a.Wm = Wm; a.sy = h / Hm; a.sx = w / Wm; a.h = h; a.w = w; a.SP = SP; a.R = ksize / 2;
    for (int dy = 0; dy <= a.R; ++dy) {                                 // the structuring elements of utils/__init__.py:590-608
        int wv;
        if (shape == CRW_DILATE_L1) wv = a.R - dy;
        else if (shape == CRW_DILATE_CROSS) wv = dy == 0 ? a.R : 0;
        else for (wv = 0; (wv + 1) * (wv + 1) + dy * dy <= a.R * a.R; ++wv) {}
        a.hw[dy] = (unsigned char)wv;
    }
    cudaMemsetAsync(ws.size, 0, sizeof(int) * (size_t)B * T * SP, (cudaStream_t)stream);
    const int64_t rows = (int64_t)B * T * h;
    const int grid0 = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
    CRW_LAUNCH(segdil_runs_kernel, grid0, 256, 0, stream, a, rows);
    e = check_launch("segdil_runs");
    if (e != CRW_OK) return e;
    {   // strips of whole cells, at most 256 columns, bitmap at most 160 KB
        const size_t col_bytes = (size_t)a.sy * SP * sizeof(unsigned);
        int nw = (Wm * a.sx + 31) / 32;
        if (nw > 8) nw = 8;
        if ((size_t)nw * col_bytes > 160 * 1024) nw = (int)((160 * 1024) / col_bytes);
        a.NW = nw < 1 ? 1 : nw;
        a.cpc = a.NW * 32 / a.sx;
        if (a.cpc > Wm) a.cpc = Wm;
        a.nxc = (Wm + a.cpc - 1) / a.cpc;
        const size_t smem = (size_t)a.NW * col_bytes;
        auto k = segdil_count_kernel;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CRW_LAUNCH(k, B * T * Hm * a.nxc, 256, smem, stream, a, ws);
    }
    e = check_launch("segdil_count");
    if (e != CRW_OK) return e;
    return seg_fwd_tail(maps, ws, B, C, T, cells, SP, out, stream);
}

extern "C" int crw_segmean_dilated_bwd(const float* grad_out, const void* workspace, size_t workspace_bytes,
                                       int B, int C, int T, int Hm, int Wm, int h, int w, int SP,
                                       float* grad_maps, crw_stream_t stream) {
    int e = segdil_check(B, C, T, Hm, Wm, h, w, SP, 1, CRW_DILATE_L1);
    if (e != CRW_OK) return e;
    unsigned* runs;
    int* nruns;
    SegWs ws = segdil_ws(const_cast<void*>(workspace), B, T, Hm * Wm, h, w, SP, &runs, &nruns);
    if (!workspace || workspace_bytes < ws.bytes) { set_error("segmean_dilated_bwd: workspace too small"); return CRW_ERR_SHAPE; }
    return seg_bwd_run(grad_out, ws, B, C, T, Hm * Wm, SP, grad_maps, true, stream);
}
