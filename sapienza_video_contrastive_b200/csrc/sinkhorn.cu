// sinkhorn.cu - Sinkhorn-Knopp normalisation (utils/__init__.py:615-641) behind CRW.stoch_mat(do_sinkhorn=True)
// (model.py:83-87).  The reference iterates "L1-normalise the columns, L1-normalise the rows" until the standard deviation
// of ALL column sums of the batch drops under `tol` - a data-dependent stop that it evaluates on the host every sweep; so
// does this entry point (one 4-byte read-back per sweep).  One CTA per matrix; matrices stay in global memory (L2-resident
// at the walk's sizes), every pass reads and writes rows with consecutive threads on consecutive columns.
#include "common.cuh"

namespace crw {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

// A <- exp(A / tau) if asked (model.py:84), then A <- A / sum(A) per matrix (utils/__init__.py:618-621)
__global__ void __launch_bounds__(256) sk_init_kernel(float* __restrict__ A, int64_t MS, int apply_exp, float tau) {
    __shared__ float red[8];
    float* a = A + (int64_t)blockIdx.x * MS;
    float s = 0.f;
    for (int64_t e = threadIdx.x; e < MS; e += 256) {
        float v = a[e];
        if (apply_exp) { v = expf(__fdiv_rn(v, tau)); a[e] = v; }
        s += v;
    }
    const float tot = block_sum_256(s, red);
    for (int64_t e = threadIdx.x; e < MS; e += 256) a[e] = __fdiv_rn(a[e], tot);
}

// one sweep: columns (dim -2), then rows (dim -1), F.normalize(p=1) semantics: x / max(sum |x|, 1e-12); leaves the signed
// column sums of the result in colsum
__global__ void __launch_bounds__(256) sk_sweep_kernel(float* __restrict__ A, float* __restrict__ colnorm, float* __restrict__ colsum,
                                                       int N, int M) {
    float* a = A + (int64_t)blockIdx.x * N * M;
    float* cn = colnorm + (int64_t)blockIdx.x * M;
    float* cs = colsum + (int64_t)blockIdx.x * M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < M; c += 256) {
        float s = 0.f;
        for (int r = 0; r < N; ++r) s += fabsf(a[(int64_t)r * M + c]);
        cn[c] = fmaxf(s, kEpsNorm);
    }
    __syncthreads();
    for (int r = warp; r < N; r += 8) {
        float* row = a + (int64_t)r * M;
        float s = 0.f;
        for (int c = lane; c < M; c += 32) {
            const float v = __fdiv_rn(row[c], cn[c]);
            row[c] = v;
            s += fabsf(v);
        }
        s = fmaxf(warp_sum(s), kEpsNorm);
        for (int c = lane; c < M; c += 32) row[c] = __fdiv_rn(row[c], s);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < M; c += 256) {
        float s = 0.f;
        for (int r = 0; r < N; ++r) s += a[(int64_t)r * M + c];
        cs[c] = s;
    }
}

// unbiased standard deviation of n values (torch.std default), one CTA, fixed summation tree
__global__ void __launch_bounds__(256) sk_std_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    __shared__ float red[8];
    float s = 0.f;
    for (int64_t e = threadIdx.x; e < n; e += 256) s += x[e];
    const float mean = block_sum_256(s, red) / (float)n;
    float q = 0.f;
    for (int64_t e = threadIdx.x; e < n; e += 256) { const float d = x[e] - mean; q = fmaf(d, d, q); }
    const float ss = block_sum_256(q, red);
    if (threadIdx.x == 0) out[0] = sqrtf(ss / (float)(n - 1));           // n = 1: 0 / 0 = NaN, like torch
}

}  // namespace crw

using namespace crw;

extern "C" size_t crw_sinkhorn_workspace_bytes(int64_t R, int N, int M) {
    if (R <= 0 || N <= 0 || M <= 0) return 0;
    return ((size_t)R * M * 2 * sizeof(float) + 255) / 256 * 256 + 256;
}

extern "C" int crw_sinkhorn_knopp(float* A, int64_t R, int N, int M, int apply_exp, float temperature, float tol, int max_iter,
                                  int* iterations, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    if (R <= 0 || N <= 0 || M <= 0 || R > 0x7fffffff) { set_error("sinkhorn: bad shape R=%lld N=%d M=%d", (long long)R, N, M); return CRW_ERR_SHAPE; }
    if (apply_exp && !(temperature > 0.f)) { set_error("sinkhorn: temperature must be > 0"); return CRW_ERR_SHAPE; }
    if (!workspace || workspace_bytes < crw_sinkhorn_workspace_bytes(R, N, M)) { set_error("sinkhorn: workspace too small"); return CRW_ERR_SHAPE; }
    float* colnorm = (float*)workspace;
    float* colsum = colnorm + (size_t)R * M;
    float* dstd = (float*)((char*)workspace + ((size_t)R * M * 2 * sizeof(float) + 255) / 256 * 256);
    CRW_LAUNCH(sk_init_kernel, (int)R, 256, 0, stream, A, (int64_t)N * M, apply_exp, temperature);
    int e = check_launch("sinkhorn_init");
    if (e != CRW_OK) return e;
    int it = 0;
    float hstd = 0.f;
    do {     // utils/__init__.py:625: at least one sweep, then while the column sums still spread more than tol
        CRW_LAUNCH(sk_sweep_kernel, (int)R, 256, 0, stream, A, colnorm, colsum, N, M);
        CRW_LAUNCH(sk_std_kernel, 1, 256, 0, stream, (const float*)colsum, (int64_t)R * M, dstd);
        e = check_launch("sinkhorn_sweep");
        if (e != CRW_OK) return e;
        if (cudaMemcpyAsync(&hstd, dstd, sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
            cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
            set_error("sinkhorn: reading the stop criterion back failed");
            return CRW_ERR_CUDA;
        }
        ++it;
    } while (hstd > tol && it < max_iter);
    if (iterations) *iterations = it;
    return CRW_OK;
}
