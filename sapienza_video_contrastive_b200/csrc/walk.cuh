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
    void* ws_tc;               // general path: tensor-core GEMM operand planes (nullptr = SIMT GEMM only)
    size_t ws_tc_bytes;
    // fused (small-N) path exchange buffers, all compact (stride N)
    float *ws_F, *ws_G, *ws_dF, *ws_dG;        // (B,T-1,N,N)
    float *ws_s12, *ws_s21;                    // (B,T-1,N) row denominators
    float *ws_invn, *ws_nrm;                   // (B,T,N)
    float *ws_dqa, *ws_dqb;                    // (B,T-1,N,D) per-pair contributions to dQ_i / dQ_{i+1}
    unsigned* ws_clipcnt;                      // (B,T) per-frame tickets of the pair-backward CTAs (zero between launches)
    // teacher-student walk (teacherstudent.py:472-580), general path only
    const float* ts_target;    // teacher chain products (B,T-2,N,N): adds the soft cross-entropy against them, or nullptr
    float ts_alpha;            // loss = alpha * walk loss + (1 - alpha) * teacher-student loss
    float* ts_xent;            // (T-2 + 1) per-walk soft cross-entropies and their mean
    float* chains_out;         // (B,T-2,N,N): the chain products of this call (a teacher's), or nullptr
};

struct FusedLayout {
    int NP, MS, DP;
    size_t pair_bytes;     // walk_pairs_fwd: two staged frames + raw affinity + codes
    size_t chain_bytes;    // walk_chain: F, G, P, S stacks + 3 scratch matrices + reduction scratch
    size_t pairb_bytes;    // walk_pairs_bwd: two staged frames + dA + raw affinity + F, dF, G, dG + codes + flag
};

__host__ __device__ __forceinline__ FusedLayout fused_layout(int N, int T, int D) {
    FusedLayout L;
    int np = (N + 3) & ~3;
    if (((np >> 2) & 1) == 0) np += 4;           // NP/4 odd: conflict-free 128-bit accesses to neighbouring rows
    L.NP = np;
    L.MS = N * np;
    L.DP = D + 4;
    const size_t codes = (size_t)((N * N + 15) & ~15);
    L.pair_bytes = sizeof(float) * ((size_t)2 * N * L.DP + L.MS) + codes;
    const int nm = 2 * (T - 1) + (T >= 3 ? 2 * (T - 2) + 3 : 0);
    L.chain_bytes = sizeof(float) * ((size_t)nm * L.MS + 64);     // + reduction scratch (32) + mbarrier
    L.pairb_bytes = sizeof(float) * ((size_t)2 * N * L.DP + 6 * L.MS) + codes + 16;
    return L;
}

constexpr int kFusedMaxT = 32;
constexpr size_t kMaxDynSmem = 232448;            // 227 KB opt-in limit per CTA on sm_100

inline bool fused_fits(int N, int T, int D) {
    if (!(N <= 64 && T >= 1 && T <= kFusedMaxT && D % 4 == 0 && D <= 256)) return false;
    const FusedLayout L = fused_layout(N, T, D);
    return L.pair_bytes <= kMaxDynSmem && L.chain_bytes <= kMaxDynSmem && L.pairb_bytes <= kMaxDynSmem;
}

constexpr int kChainCluster = 4;                   // CTAs (SMs) per clip in the clustered chain kernel
bool chain_cluster_fits(int N, int T);
int launch_walk_chain_cluster(const WalkParams& p, size_t smem_bytes, crw_stream_t stream);
int launch_walk_fused(const WalkParams& p, crw_stream_t stream);
int launch_walk_general(const WalkParams& p, crw_stream_t stream);
size_t walk_general_mats_floats(int B, int N, int T);
size_t walk_general_stat_floats(int B, int N, int T);

}  // namespace crw
