// lp_tc.cuh - host interface of the tcgen05 label-propagation top-k kernel (lp_tc.cu), called by the dispatcher in lp_simt.cu.
#pragma once
#include "common.cuh"

namespace crw {

struct LpTcArgs {
    const int64_t* key_frames;   // (Nt, S)
    const int64_t* query_frames; // (Nt)
    int Nt, S, n_long, h, w, C, k, R, r2i;   // r2i: largest integer d2 admitted (d2 <= r2i  <=>  d2 < radius^2)
    int restricted;
    float tau;
    float* Ws;
    int64_t* Is;
    unsigned flags;              // CRW_LP_* flags of the call
    unsigned* err;               // device error flag (barrier timeout)
    // filled by launch_lp_tc:
    float* short_v;              // (Nt*hw, 16) shortlist scores, descending
    int* short_i;                // (Nt*hw, 16) shortlist ids slot*hw + pos
    const int* list;             // work list of tiles (n * tiles + tile) for the fp32-faithful pass, or null = every tile
    const unsigned* count;       // its length (device)
};

size_t lp_tc_workspace_bytes(int Nf, int Nt, int h, int w, int C);
bool lp_tc_supported(int C, int k, float radius, int R, bool dense);
int launch_lp_tc(const float* feats, int Nf, const LpTcArgs& a, void* workspace, size_t workspace_bytes, crw_stream_t stream);

}  // namespace crw
