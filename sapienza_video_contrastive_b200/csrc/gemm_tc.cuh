// gemm_tc.cuh - host interface of the batched fp32-faithful tcgen05 GEMM (gemm_tc.cu), used by the large-graph walk.
#pragma once
#include "common.cuh"

namespace crw {

struct TcOperand {           // element (r, k) of matrix z: p[(z / nj) * sb + (z % nj) * sj + r * rs + k * cs]
    const float* p;
    int64_t sb, sj, rs, cs;
};

// C[z] (M x N, fp32, row stride ldc, matrix z at C + (z / nj) * csb + (z % nj) * csj)  (+)=  sum_t A_t[z] (M x K_t) B_t[z] (K_t x N)
struct TcGemmCall {
    TcOperand A[2], B[2];        // B given as (k, c): element (k, c) of B_t = p[... + k * rs + c * cs]
    int K[2];
    int nterms;
    float* C;
    int64_t csb, csj, ldc;
    int M, N, nb, nj, accumulate;
};

size_t gemm_tc_workspace_bytes(int M, int N, int Kmax, int Z);
bool gemm_tc_eligible(int M, int N, int Kmin, int Kmax);
int gemm_tc_run(const TcGemmCall& c, void* workspace, size_t workspace_bytes, crw_stream_t stream);
int gemm_tc_check(void* workspace, crw_stream_t stream);          // reads the error word back (synchronises the stream)

}  // namespace crw
