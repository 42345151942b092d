// walk.cuh - shared declarations of the walk kernels (walk_fused.cu, walk_general.cu, walk_api.cu).
#pragma once
#include "common.cuh"

namespace crw {

struct WalkParams {
    const float* feats;        // (B,N,T,D) pre-normalisation node vectors
    float* q;                  // (B,N,T,D) unit-norm
    float* xent;               // (T-2)
    float* acc;                // (T-2)
    float* grad;               // (B,N,T,D) or nullptr
    const float* u12;          // (T-1,B,N,N) or nullptr
    const float* u21p;         // (T-1,B,N,N) physical layout, or nullptr
    uint64_t seed, offset;     // torch Philox state for the in-kernel replay
    uint32_t pthreads, pinc;   // torch rand launch threads; offset increment per draw
    uint64_t* dev_state;       // optional device {seed, offset}, advanced after the launch
    int B, N, T, D;
    float tau, rate;
    unsigned flags;
    // workspace carve-up
    float* ws_araw;            // (B,T-1,N,N) raw affinities
    unsigned char* ws_codes;   // (B,T-1,N,N) dropout codes: bit0 forward draw, bit1 backward draw
    float* ws_partial;         // (B,T-2,2) per-clip loss / accuracy sums
    unsigned* ws_counter;      // 1 ticket counter (zero between launches)
    float* ws_mats;            // general path: transition / chain / gradient matrices
    float* ws_stat;            // general path: row denominators, norms
};

struct FusedLayout {
    int NP, MS, DP;
    int off_F, off_G, off_R, off_stat, off_codes;   // in floats
    size_t bytes;
};

__host__ __device__ __forceinline__ FusedLayout fused_layout(int N, int T, int D) {
    FusedLayout L;
    int np = (N + 3) & ~3;
    if (((np >> 2) & 1) == 0) np += 4;           // NP/4 odd: conflict-free 128-bit row accesses
    L.NP = np;
    L.MS = N * np;
    L.DP = D + 4;
    L.off_F = 0;
    L.off_G = (T - 1) * L.MS;
    L.off_R = 2 * (T - 1) * L.MS;
    const int chain = (T >= 3 ? (2 * (T - 2) + 3) : 0) * L.MS;
    const int stage = L.MS + 2 * N * L.DP;
    const int r = chain > stage ? chain : stage;
    L.off_stat = L.off_R + r;
    const int stat = 2 * (T - 1) * N + 2 * T * N + 64;
    L.off_codes = (L.off_stat + stat + 3) & ~3;
    L.bytes = (size_t)L.off_codes * 4 + (size_t)((N * N + 15) & ~15);
    return L;
}

constexpr int kFusedMaxT = 32;
constexpr size_t kMaxDynSmem = 232448;            // 227 KB opt-in limit per CTA on sm_100

inline bool fused_fits(int N, int T, int D) {
    return N <= 64 && T >= 2 && T <= kFusedMaxT && D % 4 == 0 && D <= 256 && fused_layout(N, T, D).bytes <= kMaxDynSmem;
}

int launch_walk_fused(const WalkParams& p, crw_stream_t stream);
int launch_walk_general(const WalkParams& p, crw_stream_t stream);
size_t walk_general_mats_floats(int B, int N, int T);
size_t walk_general_stat_floats(int B, int N, int T);

}  // namespace crw
