// segmean.cu - superpixel node pooling (reference: code/model.py:296-325, CRW.image_to_nodes, which builds a
// (B,T,SP,h,w) one-hot, counts it per 8x8 window with utils.view_as_windows (utils/__init__.py:433-584) and
// multiplies a (B,T,H,W,C,SP) broadcast).  Mathematically that is a SEGMENT MEAN of the nearest-neighbour
// upsampled feature map over each superpixel label; here it is computed as a scatter-reduce
// straight from the label map:
//   1. segmean_count: one warp per feature cell turns its sy*sx labels into a short (label,count) list
//      (<= 64 entries, typically 1-4) and adds the counts to the per-label sizes;
//   2. segmean_accum: one CTA per (clip, frame, 32-channel tile) stages the feature tile transposed in shared memory and
//      reduces it per label through per-label cell bitmasks (see the kernel); the epilogue divides by the sizes;
//   3. segmean_bwd mirrors 2 as a gather (deterministic).
// Neither the one-hot nor the broadcast product is ever materialised; forward and backward are deterministic.
#include "common.cuh"

namespace crw {

constexpr int kSegCT = 32;       // channels per CTA
constexpr int kSegMaxEnt = 64;   // max distinct labels per feature cell (= max sy*sx)

struct SegWs {
    int* size;               // (B*T, SP)
    unsigned char* nent;     // (B*T, cells)
    unsigned* ent;           // (B*T, cap, cells): label << 8 | count
    int cap;
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
    o += (size_t)B * T * cap * cells * sizeof(unsigned);
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
    for (int64_t gc = warp; gc < total_cells; gc += nw) {
        const int64_t bt = gc / cells;
        const int cell = (int)(gc - bt * cells);
        const int b = (int)(bt / T), t = (int)(bt - (int64_t)b * T);
        const int cy = cell / Wm, cx = cell - cy * Wm;
        int64_t lab[2];
        bool pend[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pix = lane + 32 * h;
            pend[h] = pix < npix;
            lab[h] = -1;
            if (pend[h]) {
                const int py = pix / sx, px = pix - py * sx;
                lab[h] = labels[b * ls_b + t * ls_t + (int64_t)(cy * sy + py) * ls_y + (int64_t)(cx * sx + px) * ls_x];
            }
        }
        int n = 0;
        for (;;) {
            const unsigned m0 = __ballot_sync(kFull, pend[0]);
            const unsigned m1 = __ballot_sync(kFull, pend[1]);
            if (!(m0 | m1)) break;
            // leader: lowest pending pixel
            int64_t L;
            if (m0) L = __shfl_sync(kFull, lab[0], __ffs((int)m0) - 1);
            else    L = __shfl_sync(kFull, lab[1], __ffs((int)m1) - 1);
            const bool h0 = pend[0] && lab[0] == L, h1 = pend[1] && lab[1] == L;
            const int cnt = __popc(__ballot_sync(kFull, h0)) + __popc(__ballot_sync(kFull, h1));
            if (h0) pend[0] = false;
            if (h1) pend[1] = false;
            if (L >= 0 && L < SP) {              // labels outside [0,SP) never match a one-hot plane (model.py:299-301)
                if (lane == 0) {
                    ws.ent[((int64_t)bt * ws.cap + n) * cells + cell] = ((unsigned)L << 8) | (unsigned)cnt;
                    atomicAdd(ws.size + bt * SP + (int)L, cnt);
                }
                ++n;
            }
        }
        if (lane == 0) ws.nent[bt * cells + cell] = (unsigned char)n;
    }
}

// Forward accumulation, one CTA per (clip, frame, 32-channel tile):
//   * the feature tile is read once, coalesced along the cells, and stored TRANSPOSED in shared memory ([cell][32+1]) so
//     that the 32 lanes of a warp = 32 channels read one cell's features without bank conflicts;
//   * a per-label bitmask of the cells containing the label (plus a 32-bit summary of its non-empty words) is built in
//     shared memory from the per-cell lists;
//   * a warp owns one label at a time and walks the set bits of its mask in increasing cell order: every lane does one
//     useful FMA per (cell, label) entry.  No floating-point atomics: deterministic, each feature read from HBM once.
constexpr int kSegCTF = 32;          // channels per CTA in the forward
constexpr int kSegSlots = 4;         // per-cell list entries cached in shared memory (rest read from the workspace)
constexpr int kSegFwdThreads = 1024; // one CTA per SM (the transposed tile fills shared memory): many warps hide the label walks' latency

__global__ void __launch_bounds__(kSegFwdThreads, 1) segmean_accum_kernel(const float* __restrict__ maps, SegWs ws, int C, int T, int cells,
                                                            int SP, float* __restrict__ out) {
    CRW_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bt = blockIdx.y, b = bt / T, t = bt - b * T;
    const int c0 = blockIdx.x * kSegCTF;
    const int nwords = (cells + 31) / 32;
    constexpr int LD = kSegCTF + 1;
    float* tile = reinterpret_cast<float*>(smem_raw);                              // [cells][33]
    unsigned* mask = reinterpret_cast<unsigned*>(tile + (size_t)cells * LD);        // [SP][nwords]
    unsigned* occ = mask + (size_t)SP * nwords;                                     // [SP] non-empty words (nwords <= 32)
    unsigned* ents = occ + SP;                                                      // [kSegSlots][cells]
    unsigned char* nes = reinterpret_cast<unsigned char*>(ents + kSegSlots * cells);  // [cells]

    for (int e = tid; e < SP * nwords; e += kSegFwdThreads) mask[e] = 0u;
    for (int e = tid; e < SP; e += kSegFwdThreads) occ[e] = 0u;
    // transposing tile load: warp w streams channels w, w+8, ... (coalesced along the cells, 8 x 128-bit loads in flight per lane)
    const bool vec = (cells & 127) == 0 && ((reinterpret_cast<uintptr_t>(maps) & 15) == 0);
    for (int cl = warp; cl < kSegCTF; cl += kSegFwdThreads / 32) {
        const int c = c0 + cl;
        const float* src = maps + (((int64_t)b * C + min(c, C - 1)) * T + t) * cells;
        if (vec) {
            for (int base = 0; base < cells; base += 1024) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int cell = base + (u * 32 + lane) * 4;
                    v[u] = cell < cells ? __ldg(reinterpret_cast<const float4*>(src + cell)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int cell = base + (u * 32 + lane) * 4;
                    if (cell < cells) {
                        const float m = c < C ? 1.f : 0.f;
                        tile[(cell + 0) * LD + cl] = v[u].x * m;
                        tile[(cell + 1) * LD + cl] = v[u].y * m;
                        tile[(cell + 2) * LD + cl] = v[u].z * m;
                        tile[(cell + 3) * LD + cl] = v[u].w * m;
                    }
                }
            }
        } else {
            for (int cell = lane; cell < cells; cell += 32) tile[cell * LD + cl] = c < C ? __ldg(src + cell) : 0.f;
        }
    }
    __syncthreads();
    const unsigned char* gnent = ws.nent + (int64_t)bt * cells;
    const unsigned* gent = ws.ent + (int64_t)bt * ws.cap * cells;
    for (int cell = tid; cell < cells; cell += kSegFwdThreads) {
        const int ne = gnent[cell];
        nes[cell] = (unsigned char)ne;
        for (int sl = 0; sl < ne; ++sl) {
            const unsigned e = gent[(int64_t)sl * cells + cell];
            if (sl < kSegSlots) ents[sl * cells + cell] = e;
            atomicOr(mask + (size_t)(e >> 8) * nwords + (cell >> 5), 1u << (cell & 31));
            atomicOr(occ + (e >> 8), 1u << (cell >> 5));
        }
    }
    __syncthreads();
    const int* size = ws.size + (int64_t)bt * SP;
    for (int s = warp; s < SP; s += kSegFwdThreads / 32) {
        float acc = 0.f;
        unsigned ow = occ[s];
        while (ow) {                                                    // warp-uniform control flow throughout
            const int w = __ffs((int)ow) - 1;
            ow &= ow - 1;
            unsigned bits = mask[(size_t)s * nwords + w];
            while (bits) {
                const int cell = w * 32 + __ffs((int)bits) - 1;
                bits &= bits - 1;
                const int ne = nes[cell];
                float cnt = 0.f;
                for (int sl = 0; sl < ne; ++sl) {
                    const unsigned e = sl < kSegSlots ? ents[sl * cells + cell] : gent[(int64_t)sl * cells + cell];
                    if ((int)(e >> 8) == s) { cnt = (float)(e & 255u); break; }
                }
                acc = fmaf(cnt, tile[cell * LD + lane], acc);
            }
        }
        if (c0 + lane < C) out[(((int64_t)b * SP + s) * T + t) * C + c0 + lane] = acc / ((float)size[s] + kEpsLog);
    }
}

__global__ void __launch_bounds__(256) segmean_bwd_kernel(const float* __restrict__ gout, SegWs ws, int C, int T, int cells,
                                                          int SP, float* __restrict__ gmaps) {
    CRW_DYN_SMEM(smem_raw);
    float* wg = reinterpret_cast<float*>(smem_raw);             // [SP][kSegCT + 1] = gout / size
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bt = blockIdx.y, b = bt / T, t = bt - b * T;
    const int c0 = blockIdx.x * kSegCT;
    constexpr int LD = kSegCT + 1;
    const int* size = ws.size + (int64_t)bt * SP;
    for (int e = tid; e < SP * kSegCT; e += 256) {
        const int s = e / kSegCT, cl = e - s * kSegCT;
        wg[s * LD + cl] = (c0 + cl < C) ? __ldg(gout + (((int64_t)b * SP + s) * T + t) * C + c0 + cl) / ((float)size[s] + kEpsLog) : 0.f;
    }
    __syncthreads();
    const unsigned char* nent = ws.nent + (int64_t)bt * cells;
    const unsigned* ent = ws.ent + (int64_t)bt * ws.cap * cells;
    for (int cell = lane; cell < cells; cell += 32) {
        const int ne = nent[cell];
        for (int cl = warp; cl < kSegCT; cl += 8) {
            const int c = c0 + cl;
            if (c >= C) break;
            float g = 0.f;
            for (int s = 0; s < ne; ++s) {
                const unsigned e = __ldg(ent + (int64_t)s * cells + cell);
                g = fmaf((float)(e & 255u), wg[(e >> 8) * LD + cl], g);
            }
            gmaps[(((int64_t)b * C + c) * T + t) * cells + cell] = g;
        }
    }
}

static int seg_check(int B, int C, int T, int Hm, int Wm, int h, int w, int SP, int* cap) {
    if (B <= 0 || C <= 0 || T <= 0 || Hm <= 0 || Wm <= 0 || SP <= 0 || h % Hm || w % Wm) {
        set_error("segmean: bad shape B=%d C=%d T=%d maps=%dx%d labels=%dx%d SP=%d", B, C, T, Hm, Wm, h, w, SP);
        return CRW_ERR_SHAPE;
    }
    const int npix = (h / Hm) * (w / Wm);
    if (npix > kSegMaxEnt || SP >= (1 << 24) || (size_t)SP * (kSegCT + 1) * 4 > 227 * 1024) {
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
    CRW_LAUNCH(segmean_count_kernel, grid1, 256, 0, stream, labels, ls_b, ls_t, ls_y, ls_x, T, Hm, Wm, h / Hm, w / Wm, SP, ws, total);
    e = check_launch("segmean_count");
    if (e != CRW_OK) return e;
    const int nwords = (cells + 31) / 32;
    const size_t smem = sizeof(float) * (size_t)cells * (kSegCTF + 1) + sizeof(unsigned) * ((size_t)SP * nwords + SP + (size_t)kSegSlots * cells) +
                        (size_t)((cells + 15) & ~15) + 16;
    if (smem > 227 * 1024 || nwords > 32) {
        set_error("segmean_fwd: unsupported size (cells=%d, SP=%d)", cells, SP);
        return CRW_ERR_UNSUPPORTED;
    }
    auto k = segmean_accum_kernel;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((C + kSegCTF - 1) / kSegCTF, B * T);
    CRW_LAUNCH(k, grid, kSegFwdThreads, smem, stream, maps, ws, C, T, cells, SP, out);
    return check_launch("segmean_accum");
}

extern "C" int crw_segmean_bwd(const float* grad_out, const void* workspace, size_t workspace_bytes,
                               int B, int C, int T, int Hm, int Wm, int h, int w, int SP,
                               float* grad_maps, crw_stream_t stream) {
    int cap = 0;
    int e = seg_check(B, C, T, Hm, Wm, h, w, SP, &cap);
    if (e != CRW_OK) return e;
    const int cells = Hm * Wm;
    SegWs ws = seg_ws(const_cast<void*>(workspace), B, T, cells, SP, cap);
    if (!workspace || workspace_bytes < ws.bytes) { set_error("segmean_bwd: workspace too small"); return CRW_ERR_SHAPE; }
    const size_t smem = (size_t)SP * (kSegCT + 1) * sizeof(float);
    auto k = segmean_bwd_kernel;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((C + kSegCT - 1) / kSegCT, B * T);
    CRW_LAUNCH(k, grid, 256, smem, stream, grad_out, ws, C, T, cells, SP, grad_maps);
    return check_launch("segmean_bwd");
}
