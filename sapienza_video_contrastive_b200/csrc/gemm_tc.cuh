// gemm_tc.cuh - host interface of the batched fp32-faithful tcgen05 GEMM (gemm_tc.cu), used by the large-graph walk.
#pragma once
#include "common.cuh"

namespace crw {

struct TcOperand {           // element (r, k) of matrix z: p[(z / nj) * sb + (z % nj) * sj + r * rs + k * cs]
    const float* p;
    int64_t sb, sj, rs, cs;
};

// One launch computes up to two independent batched products of identical shape ("groups", e.g. P_j = P_{j-1} X_j and
// S_j = Y_j S_{j-1}), each the sum of up to two terms:
//   C_g[z] (M x N, fp32, row stride ldc, matrix z at C + (z / nj) * csb + (z % nj) * csj)  (+)=  sum_t A_gt[z] (M x K_t) B_gt[z] (K_t x N)
// The terms of a product are concatenated along K into one pair of operand planes (common per-row exponents), so a
// two-term product is still a single pass of the tensor-core pipeline.
struct TcGroup {
    TcOperand A[2], B[2];        // B given as (k, c): element (k, c) of B_t = p[... + k * rs + c * cs]
    float* C;
    int64_t csb, csj, ldc;
    int accumulate;
};
struct TcGemmCall {
    TcGroup grp[2];
    int ngroups;
    int K[2];
    int nterms;
    int M, N, nb, nj;
};

// Kcat = sum over terms of K_t rounded up to 8; Z = ngroups * nb * nj
size_t gemm_tc_workspace_bytes(int M, int N, int Kcat, int Z);
bool gemm_tc_eligible(int M, int N, int Kmin, int Kmax);
int gemm_tc_run(const TcGemmCall& c, void* workspace, size_t workspace_bytes, crw_stream_t stream);

// gemm_tf32.cu: the same call on tcgen05 kind::tf32 with the operands read in place by TMA and split inside the kernel (no
// workspace, no separate split launch); for products too small to amortise gemm_tc's operand pass
bool gemm_tf32_eligible(const TcGemmCall& c);
int gemm_tf32_run(const TcGemmCall& c, unsigned* err_word, crw_stream_t stream);          // reads the error word back (synchronises the stream)

}  // namespace crw
