// walk_general.cu - the walk of code/model.py:366-413 for clips that do not fit one CTA's shared memory
// (superpixel graphs, fine-stride grids, long clips): the same algebra as walk_fused.cu, but every stage is its
// own batched kernel over all clips and the matrices live in a caller-provided workspace (L2 / HBM).
//
// Stages (all batched over clips b and pairs / walks j):
//   norm -> A_i = Q_i Q_{i+1}^T and A_i^T (two GEMMs so both row passes stay coalesced) -> row-stochastic F_i, G_i
//   -> chain levels P_j, S_j -> W_j -> loss rows + dW_j -> reverse sweep gP_j, gS_j -> dX_j, dY_j
//   -> transition-matrix backward (in place over A_i / A_i^T) -> dQ -> normalisation backward.
// The GEMM here is a plain fp32 SIMT tile kernel (64x64x16, 4x4 per thread) that takes up to two product terms
// and arbitrary operand strides; it is exact-fp32 like the reference's cuBLAS sgemm.
#include "walk.cuh"
#include "gemm_tc.cuh"

namespace crw {

// element (r,c) of matrix (b,j) lives at p[b*sb + j*sj + r*rs + c*cs]
struct MatRef {
    const float* p;
    int64_t sb, sj, rs, cs;
};

struct GemmArgs {
    MatRef A[2], B[2];       // C = sum_t A[t] (M x K[t]) * B[t] (K[t] x N)
    int K[2];
    int nterms;
    float* C;
    int64_t csb, csj, ldc;
    int M, N, nj;
    int accumulate;
    int ktot;                // split-K: if > 0, batch j covers k in [j*K[0], min((j+1)*K[0], ktot))
};

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int z = blockIdx.z, b = z / g.nj, j = z - b * g.nj;
    const int r0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
    for (int t = 0; t < g.nterms; ++t) {
        const float* Ap = g.A[t].p + b * g.A[t].sb + j * g.A[t].sj;
        const float* Bp = g.B[t].p + b * g.B[t].sb + j * g.B[t].sj;
        const int64_t ars = g.A[t].rs, acs = g.A[t].cs, brs = g.B[t].rs, bcs = g.B[t].cs;
        int K = g.K[t];
        if (g.ktot > 0) K = min(K, g.ktot - j * g.K[t]);
        for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
                int r, k;
                if (acs == 1) { k = tid & 15; r = (tid >> 4) + 16 * pass; }      // k contiguous in memory
                else          { r = tid & 63; k = (tid >> 6) + 4 * pass; }       // r contiguous (or generic)
                float v = 0.f;
                if (r0 + r < g.M && k0 + k < K) v = __ldg(Ap + (int64_t)(r0 + r) * ars + (int64_t)(k0 + k) * acs);
                As[k][r] = v;
            }
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
                int c, k;
                if (brs == 1) { k = tid & 15; c = (tid >> 4) + 16 * pass; }      // k contiguous
                else          { c = tid & 63; k = (tid >> 6) + 4 * pass; }       // c contiguous (or generic)
                float v = 0.f;
                if (c0 + c < g.N && k0 + k < K) v = __ldg(Bp + (int64_t)(k0 + k) * brs + (int64_t)(c0 + c) * bcs);
                Bs[k][c] = v;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(av[i], bv[jj], acc[i][jj]);
            }
            __syncthreads();
        }
    }
    float* Cp = g.C + b * g.csb + j * g.csj;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r >= g.M) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int c = c0 + tx * 4 + jj;
            if (c >= g.N) continue;
            float* o = Cp + (int64_t)r * g.ldc + c;
            *o = g.accumulate ? (*o + acc[i][jj]) : acc[i][jj];
        }
    }
}

// tensor-core workspace of the walk in flight (nullptr: SIMT only)
struct TcCtx {
    void* ws;
    size_t bytes;
    bool no_tf32;
};
constexpr int kTf32MaxDim = 512;

static int run_gemm_simt(const GemmArgs& g, int nb, crw_stream_t stream) {
    if (g.M <= 0 || g.N <= 0 || nb * g.nj <= 0) return CRW_OK;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, nb * g.nj);
    CRW_LAUNCH(gemm_f32_kernel, grid, 256, 0, stream, g);
    return check_launch("gemm_f32");
}

static void tc_group_of(TcGroup& o, const GemmArgs& g) {
    for (int t = 0; t < g.nterms; ++t) {
        o.A[t] = TcOperand{g.A[t].p, g.A[t].sb, g.A[t].sj, g.A[t].rs, g.A[t].cs};
        o.B[t] = TcOperand{g.B[t].p, g.B[t].sb, g.B[t].sj, g.B[t].rs, g.B[t].cs};
    }
    o.C = g.C; o.csb = g.csb; o.csj = g.csj; o.ldc = g.ldc; o.accumulate = g.accumulate;
}

// One or two independent products of identical shape (same M, N, K's, term count and batch): a single tensor-core launch
// when the walk has a tensor-core workspace and the shape is worth it, the SIMT kernel otherwise.
static int run_gemms(const GemmArgs* gs, int ng, int nb, crw_stream_t stream, const TcCtx* tc) {
    const GemmArgs& g = gs[0];
    if (g.M <= 0 || g.N <= 0 || nb * g.nj <= 0) return CRW_OK;
    if (tc && tc->ws && g.ktot == 0 && g.nterms >= 1) {
        int kmin = g.K[0], kmax = g.K[0], kcat = 0;
        for (int t = 0; t < g.nterms; ++t) {
            kmin = g.K[t] < kmin ? g.K[t] : kmin;
            kmax = g.K[t] > kmax ? g.K[t] : kmax;
            kcat += (g.K[t] + 7) & ~7;
        }
        if (gemm_tc_eligible(g.M, g.N, kmin, kmax) && gemm_tc_workspace_bytes(g.M, g.N, kcat, ng * nb * g.nj) <= tc->bytes) {
            TcGemmCall c{};
            for (int i = 0; i < ng; ++i) tc_group_of(c.grp[i], gs[i]);
            c.ngroups = ng; c.nterms = g.nterms; c.K[0] = g.K[0]; c.K[1] = g.K[1];
            c.M = g.M; c.N = g.N; c.nb = nb; c.nj = g.nj;
            // mid-size products: operands read in place and split inside the kernel (gemm_tf32.cu); large ones amortise the
            // separate fp16 operand pass and run the faster kind::f16 pipeline (gemm_tc.cu)
            if (g.M < kTf32MaxDim && g.N < kTf32MaxDim && !tc->no_tf32 && gemm_tf32_eligible(c))
                return gemm_tf32_run(c, (unsigned*)tc->ws, stream);
            const int e = gemm_tc_run(c, tc->ws, tc->bytes, stream);
            if (e != CRW_ERR_UNSUPPORTED) return e;
        }
    }
    for (int i = 0; i < ng; ++i) {
        const int e = run_gemm_simt(gs[i], nb, stream);
        if (e != CRW_OK) return e;
    }
    return CRW_OK;
}

static int run_gemm(const GemmArgs& g, int nb, crw_stream_t stream, const TcCtx* tc = nullptr) { return run_gemms(&g, 1, nb, stream, tc); }

static MatRef mref(const float* p, int64_t sb, int64_t sj, int64_t rs, int64_t cs) { return MatRef{p, sb, sj, rs, cs}; }

// ---- normalisation ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gnorm_kernel(const float* __restrict__ f, float* __restrict__ q,
                                                    float* __restrict__ invn, float* __restrict__ nrm, int64_t rows, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nw) {
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = f[r * D + d]; ss += v * v; }
        ss = warp_sum(ss);
        const float nr = sqrtf(ss), den = fmaxf(nr, kEpsNorm);
        for (int d = lane; d < D; d += 32) q[r * D + d] = f[r * D + d] / den;
        if (lane == 0) { invn[r] = 1.0f / den; nrm[r] = nr; }
    }
}

__global__ void __launch_bounds__(256) gnorm_bwd_kernel(const float* __restrict__ q, float* __restrict__ g,
                                                        const float* __restrict__ invn, const float* __restrict__ nrm,
                                                        int64_t rows, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nw) {
        float dot = 0.f;
        for (int d = lane; d < D; d += 32) dot += q[r * D + d] * g[r * D + d];
        dot = warp_sum(dot);
        if (!(nrm[r] > kEpsNorm)) dot = 0.f;
        const float in = invn[r];
        for (int d = lane; d < D; d += 32) g[r * D + d] = (g[r * D + d] - q[r * D + d] * dot) * in;
    }
}

// ---- row-stochastic forward --------------------------------------------------------------------------------------
// One warp per row of a (R, N, M) stack.  mode 0: standalone stoch_mat (drop where u[row][m] < rate, written back in
// place as -1e20).  mode 1: walk forward rows (code bit0 from the forward draw).  mode 2: walk backward rows (input is
// A^T, code = union of both draws taken at the transposed element).
struct StochArgs {
    float* A;                  // (R rows) x M, row stride M
    float* out;
    float* denom;              // per row, may be null
    unsigned char* codes;      // per element, may be null
    const unsigned char* codes_in;   // mode 2: the codes the mode-1 pass drew, in the forward orientation (no second Philox pass)
    const float* u1;
    const float* u2;
    uint64_t seed, offset;
    uint32_t pthreads, pinc;
    const uint64_t* dev_state;
    int mode, softmax;
    int64_t rows;
    int N, M, T, B;
    float tau, rate;
};

__global__ void __launch_bounds__(256) stoch_rows_kernel(StochArgs s) {
    if (s.dev_state) { s.seed = ld_cg64(s.dev_state); s.offset = ld_cg64(s.dev_state + 1); }
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int M = s.M;
    for (int64_t row = warp; row < s.rows; row += nw) {
        float* a = s.A + row * M;
        int b = 0, i = 0, n = 0;
        if (s.mode) {                       // row = (b*(T-1) + i)*N + n
            n = (int)(row % s.N);
            const int64_t bi = row / s.N;
            i = (int)(bi % (s.T - 1));
            b = (int)(bi / (s.T - 1));
        }
        const int64_t numel = (int64_t)s.B * s.N * s.N;
        auto masked = [&](int m, float& av) -> bool {
            av = a[m];
            if (!(s.rate > 0.f)) return false;
            if (s.mode == 0) {
                const bool d = s.u1[row * M + m] < s.rate;
                if (d) a[m] = kNegDrop;
                return d;
            }
            unsigned code;
            if (s.mode == 2 && s.codes_in) {
                // entry m of row n of G is A[m][n]: its draws were made (and stored) by the forward-orientation pass
                code = s.codes_in[((int64_t)(b * (s.T - 1) + i) * s.N + m) * s.N + n];
            } else {
                // element of the (n_src, m_src) affinity this entry derives from
                const int64_t e = s.mode == 1 ? ((int64_t)b * s.N + n) * s.N + m : ((int64_t)b * s.N + m) * s.N + n;
                float u1, u2;
                if (s.u1) { u1 = s.u1[(int64_t)i * numel + e]; u2 = s.u2[(int64_t)i * numel + e]; }
                else {
                    u1 = torch_uniform(s.seed, s.offset + (uint64_t)s.pinc * i, s.pthreads, (uint64_t)e);
                    u2 = torch_uniform(s.seed, s.offset + (uint64_t)s.pinc * (s.T - 1 + i), s.pthreads, (uint64_t)e);
                }
                code = (u1 < s.rate ? 1u : 0u) | (u2 < s.rate ? 2u : 0u);
            }
            if (s.codes) s.codes[row * M + m] = (unsigned char)code;
            return s.mode == 1 ? (code & 1u) != 0 : code != 0;
        };
        // pass 1: statistics
        float mx = -INFINITY, sum = 0.f;
        if (s.softmax) {
            for (int m = lane; m < M; m += 32) { float av; const bool d = masked(m, av); mx = fmaxf(mx, (d ? kNegDrop : av) / s.tau); }
            mx = warp_max(mx);
        }
        // (re-evaluating masked() in pass 2 is idempotent: mode 0 has already overwritten a[m])
        for (int m = lane; m < M; m += 32) {
            float av;
            bool d;
            if (s.softmax && s.mode == 0) { av = a[m]; d = false; }
            else d = masked(m, av);
            const float x = (d ? kNegDrop : av) / s.tau;
            float e;
            if (s.softmax) e = expf(x - mx);
            else { const float E = expf(x) - 1.0f; e = E * E; }
            s.out[row * M + m] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        if (!s.softmax) sum += kEpsZs;
        __syncwarp();
        for (int m = lane; m < M; m += 32) s.out[row * M + m] = s.out[row * M + m] / sum;
        if (s.denom && lane == 0) s.denom[row] = sum;
    }
}

// ---- loss rows: W -> dW in place, per-row loss and accuracy -------------------------------------------------------
// With a teacher (Wt != nullptr, teacherstudent.py:270-292, 545-548) the row also carries the soft cross-entropy of the
// teacher's row against log_softmax of the student's row - the student's chain PROBABILITIES are the logits there - and
// dW = alpha * d(walk loss) + (1 - alpha) * d(soft cross-entropy).
__global__ void __launch_bounds__(256) loss_rows_kernel(float* __restrict__ W, float* __restrict__ rowloss,
                                                        float* __restrict__ rowacc, int64_t rows, int N, float cgrad, int need_grad,
                                                        const float* __restrict__ Wt, float* __restrict__ rowts, float alpha) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += nw) {
        const int n = (int)(row % N);
        float* w = W + row * N;
        const float* wt = Wt ? Wt + row * N : nullptr;
        float rs = 0.f, best = -INFINITY, dot = 0.f, st = 0.f;
        int bi = 0x7fffffff;
        for (int m = lane; m < N; m += 32) {
            const float v = w[m];
            rs += v + kEpsLog;
            if (v > best) { best = v; bi = m; }
            if (wt) { const float t = wt[m]; dot = fmaf(t, v, dot); st += t; }
        }
        rs = warp_sum(rs);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(kFull, best, o);
            const int oi = __shfl_xor_sync(kFull, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const float dg = w[n] + kEpsLog;
        float se = 0.f;
        if (wt) {
            dot = warp_sum(dot);
            st = warp_sum(st);
            for (int m = lane; m < N; m += 32) se += expf(w[m] - best);
            se = warp_sum(se);
        }
        __syncwarp();
        if (lane == 0) {
            rowloss[row] = logf(rs) - logf(dg);
            rowacc[row] = (bi == n) ? 1.f : 0.f;
            if (wt) rowts[row] = (best + logf(se)) * st - dot;              // sum_m t_m (logsumexp - w_m)
        }
        if (need_grad) {
            const float ir = 1.0f / rs, idg = 1.0f / dg;
            if (wt) {
                const float ca = alpha * cgrad, cb = (1.f - alpha) * cgrad, ise = st / se;
                for (int m = lane; m < N; m += 32) {
                    const float v = w[m];
                    w[m] = ca * (ir - (m == n ? idg : 0.f)) + cb * (expf(v - best) * ise - wt[m]);
                }
            } else {
                for (int m = lane; m < N; m += 32) w[m] = cgrad * (ir - (m == n ? idg : 0.f));
            }
        }
    }
}

// sums rowloss / rowacc laid out as (B, J, N) into xent[j], acc[j]: one CTA per walk j with a fixed summation tree, and
// the last CTA to finish (ticket) adds the J results in order -> deterministic
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ rowloss, const float* __restrict__ rowacc,
                                                          float* __restrict__ xent, float* __restrict__ acc, int B, int J, int N,
                                                          const unsigned* __restrict__ tc_err, unsigned* __restrict__ ticket,
                                                          const float* __restrict__ rowts, float* __restrict__ ts, float alpha) {
    __shared__ float sl[256], sa[256], st[256];
    __shared__ unsigned last;
    const int j = blockIdx.x;
    float l = 0.f, a = 0.f, t = 0.f;
    for (int64_t e = threadIdx.x; e < (int64_t)B * N; e += 256) {
        const int64_t b = e / N, n = e - b * N;
        l += rowloss[(b * J + j) * N + n];
        a += rowacc[(b * J + j) * N + n];
        if (rowts) t += rowts[(b * J + j) * N + n];
    }
    sl[threadIdx.x] = l;
    sa[threadIdx.x] = a;
    st[threadIdx.x] = t;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sl[threadIdx.x] += sl[threadIdx.x + s];
            sa[threadIdx.x] += sa[threadIdx.x + s];
            st[threadIdx.x] += st[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        xent[j] = sl[0] / ((float)B * N);
        acc[j] = sa[0] / ((float)B * N);
        if (rowts) ts[j] = st[0] / ((float)B * N);
        __threadfence();
        last = atomicAdd(ticket, 1u) == (unsigned)J - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float tot = 0.f, tts = 0.f;
        for (int jj = 0; jj < J; ++jj) tot += ld_cg(xent + jj);
        tot /= (float)J;                                                                  // the loss of model.py:413
        if (rowts) {
            for (int jj = 0; jj < J; ++jj) tts += ld_cg(ts + jj);
            tts /= (float)J;
            ts[J] = tts;
            tot = alpha * tot + (1.f - alpha) * tts;                                      // teacherstudent.py:575
        }
        // a tensor-core pipeline that timed out (gemm_tc.cu) left garbage behind: poison the loss instead of returning it
        xent[J] = (tc_err && *tc_err) ? __int_as_float(0x7fc00000) : tot;
        *ticket = 0u;
    }
}

// ---- transition-matrix backward rows: overwrites the raw affinity with its gradient contribution -----------------
// rows of Y (transition rows), dY (their gradients); A holds raw affinities in the same orientation.
__global__ void __launch_bounds__(256) stoch_bwd_rows_kernel(float* __restrict__ A, const float* __restrict__ Y,
                                                             const float* __restrict__ dY, const float* __restrict__ denom,
                                                             const unsigned char* __restrict__ codes, unsigned code_mask,
                                                             int64_t rows, int M, float tau, int softmax) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += nw) {
        float dot = 0.f;
        for (int m = lane; m < M; m += 32) dot += Y[row * M + m] * dY[row * M + m];
        dot = warp_sum(dot);
        const float den = denom[row];
        for (int m = lane; m < M; m += 32) {
            float g = 0.f;
            const unsigned code = codes ? codes[row * M + m] : 0u;
            if (!(code & code_mask)) {
                const float y = Y[row * M + m], dy = dY[row * M + m];
                if (softmax) g = y * (dy - dot) / tau;
                else { const float E = expf(A[row * M + m] / tau); g = (dy - dot) / den * (2.0f * (E - 1.0f) * E) / tau; }
            }
            A[row * M + m] = g;
        }
    }
}

static int rows_grid(int64_t rows);

int launch_l2norm_rows(const float* f, float* q, float* invn, float* nrm, int64_t rows, int D, crw_stream_t stream) {
    if (rows <= 0) return CRW_OK;
    CRW_LAUNCH(gnorm_kernel, rows_grid(rows), 256, 0, stream, f, q, invn, nrm, rows, D);
    return check_launch("l2norm_rows");
}

static int rows_grid(int64_t rows) {
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    return (int)(blocks > 0 ? blocks : 1);
}

__global__ void advance_state_kernel(uint64_t* state, uint64_t inc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state[1] += inc;
}

// copies B matrices of MS floats between strided stacks (dX_0 = gP_0, dY_0 = gS_0)
__global__ void __launch_bounds__(256) copy_mats_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t dsb,
                                                        int64_t ssb, int64_t MS, int B) {
    const int64_t n = MS * B;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / MS, o = e - b * MS;
        dst[b * dsb + o] = src[b * ssb + o];
    }
}

// workspace carve-up of the general path: every array holds all clips, (B, T-1 | T-2, N, N) contiguous
struct GenLayout {
    int64_t MS, s1, s2;         // N*N, (T-1)*MS, (T-2)*MS
    int64_t oA, oAT, oF, oG, oP, oS, oW, ogP, ogS, odX, odY, total;
};

static GenLayout gen_layout(int B, int N, int T) {
    GenLayout L;
    L.MS = (int64_t)N * N;
    const int64_t t1 = T > 1 ? T - 1 : 0, t2 = T >= 3 ? T - 2 : 0;
    L.s1 = t1 * L.MS;
    L.s2 = t2 * L.MS;
    int64_t o = 0;
    L.oA = o; o += B * L.s1;
    L.oAT = o; o += B * L.s1;
    L.oF = o; o += B * L.s1;
    L.oG = o; o += B * L.s1;
    L.oP = o; o += B * L.s2;
    L.oS = o; o += B * L.s2;
    L.oW = o; o += B * L.s2;
    L.ogP = o; o += B * L.s1;
    L.ogS = o; o += B * L.s1;
    L.odX = o; o += B * L.s1;
    L.odY = o; o += B * L.s1;
    L.total = o;
    return L;
}

size_t walk_general_mats_floats(int B, int N, int T) { return (size_t)gen_layout(B, N, T).total; }

// s12, s21 (B(T-1)N each) | invn, nrm (B N T each) | rowloss, rowacc, rowts (B (T-2) N each)
size_t walk_general_stat_floats(int B, int N, int T) {
    return (size_t)B * N * (2 * (T > 1 ? T - 1 : 0) + 2 * T + 3 * (T >= 3 ? T - 2 : 0));       // + rowts (teacher-student rows)
}

#define CRW_TRY(x) do { int _e = (x); if (_e != CRW_OK) return _e; } while (0)

int launch_walk_general(const WalkParams& p, crw_stream_t stream) {
    const int B = p.B, N = p.N, T = p.T, D = p.D;
    const GenLayout L = gen_layout(B, N, T);
    const TcCtx tc{p.ws_tc, p.ws_tc_bytes, (p.flags & CRW_WALK_NO_TF32) != 0};
    const int64_t MS = L.MS, s1 = L.s1, s2 = L.s2;
    const int t1 = T > 1 ? T - 1 : 0, t2 = T >= 3 ? T - 2 : 0;
    float* M0 = p.ws_mats;
    float* s12 = p.ws_stat;
    float* s21 = s12 + (int64_t)B * t1 * N;
    float* invn = s21 + (int64_t)B * t1 * N;
    float* nrm = invn + (int64_t)B * N * T;
    float* rowloss = nrm + (int64_t)B * N * T;
    float* rowacc = rowloss + (int64_t)B * t2 * N;
    float* rowts = rowacc + (int64_t)B * t2 * N;
    unsigned char* codesF = p.ws_codes;                          // (B,T-1,N,N), orientation of A
    unsigned char* codesG = p.ws_codes + (int64_t)B * s1;        // orientation of A^T
    const int softmax = (p.flags & CRW_WALK_SOFTMAX) ? 1 : 0;
    const bool flip = (p.flags & CRW_WALK_FLIP) != 0;
    const int64_t gs = (int64_t)T * D, cs = (int64_t)N * gs;     // node stride, clip stride in q / grad
    float *A = M0 + L.oA, *AT = M0 + L.oAT, *F = M0 + L.oF, *G = M0 + L.oG, *P = M0 + L.oP, *S = M0 + L.oS, *W = M0 + L.oW,
          *gP = M0 + L.ogP, *gS = M0 + L.ogS, *dX = M0 + L.odX, *dY = M0 + L.odY;
    float* X = flip ? G : F;
    float* Y = flip ? F : G;

    // 1. normalise (model.py:118)
    const int64_t rows = (int64_t)B * N * T;
    CRW_LAUNCH(gnorm_kernel, rows_grid(rows), 256, 0, stream, p.feats, p.q, invn, nrm, rows, D);
    CRW_TRY(check_launch("gnorm"));
    if (T < 2) return CRW_OK;

    // 2. raw affinities in both orientations (model.py:68): A_i = Q_i Q_{i+1}^T, AT_i = Q_{i+1} Q_i^T
    {
        GemmArgs g[2] = {};
        g[0].nterms = 1; g[0].K[0] = D; g[0].M = N; g[0].N = N; g[0].nj = t1; g[0].accumulate = 0;
        g[0].A[0] = mref(p.q, cs, D, gs, 1);
        g[0].B[0] = mref(p.q + D, cs, D, 1, gs);
        g[0].C = A; g[0].csb = s1; g[0].csj = MS; g[0].ldc = N;
        g[1] = g[0];
        g[1].A[0] = mref(p.q + D, cs, D, gs, 1);
        g[1].B[0] = mref(p.q, cs, D, 1, gs);
        g[1].C = AT;
        CRW_TRY(run_gemms(g, 2, B, stream, &tc));
    }
    // 3. transition rows (model.py:74-90) in both directions
    for (int dir = 1; dir <= 2; ++dir) {
        StochArgs s{};
        s.A = dir == 1 ? A : AT;
        s.out = dir == 1 ? F : G;
        s.denom = dir == 1 ? s12 : s21;
        s.codes = dir == 1 ? codesF : codesG;
        s.codes_in = dir == 2 ? codesF : nullptr;      // each element's two draws are made once, by the forward pass
        s.u1 = p.u12; s.u2 = p.u21p;
        s.seed = p.seed; s.offset = p.offset; s.pthreads = p.pthreads; s.pinc = p.pinc;
        s.dev_state = p.dev_state;
        s.mode = dir; s.softmax = softmax;
        s.rows = (int64_t)B * t1 * N; s.N = N; s.M = N; s.T = T; s.B = B;
        s.tau = p.tau; s.rate = p.rate;
        CRW_LAUNCH(stoch_rows_kernel, rows_grid(s.rows), 256, 0, stream, s);
        CRW_TRY(check_launch("stoch_rows"));
    }
    if (p.dev_state && p.rate > 0.f) {       // both row passes have consumed the state: advance it for the next launch
        CRW_LAUNCH(advance_state_kernel, 1, 32, 0, stream, p.dev_state, (uint64_t)p.pinc * 2u * (unsigned)(T - 1));
        CRW_TRY(check_launch("advance_state"));
    }
    if (T < 3) return CRW_OK;

    // 4. chains (model.py:376-380 re-associated): P_j = P_{j-1} X_j, S_j = Y_j S_{j-1}
    for (int j = 1; j <= T - 2; ++j) {
        GemmArgs g[2] = {};                              // the two chains of a level are independent: one launch
        g[0].nterms = 1; g[0].K[0] = N; g[0].M = N; g[0].N = N; g[0].nj = 1; g[0].accumulate = 0; g[0].csb = s2; g[0].csj = 0; g[0].ldc = N;
        g[0].A[0] = j == 1 ? mref(X, s1, 0, N, 1) : mref(P + (j - 2) * MS, s2, 0, N, 1);
        g[0].B[0] = mref(X + j * MS, s1, 0, N, 1);
        g[0].C = P + (j - 1) * MS;
        g[1] = g[0];
        g[1].A[0] = mref(Y + j * MS, s1, 0, N, 1);
        g[1].B[0] = j == 1 ? mref(Y, s1, 0, N, 1) : mref(S + (j - 2) * MS, s2, 0, N, 1);
        g[1].C = S + (j - 1) * MS;
        CRW_TRY(run_gemms(g, 2, B, stream, &tc));
    }
    {   // W_j = P_j S_j, all walks in one launch
        GemmArgs g{};
        g.nterms = 1; g.K[0] = N; g.M = N; g.N = N; g.nj = t2; g.accumulate = 0; g.csb = s2; g.csj = MS; g.ldc = N;
        g.A[0] = mref(P, s2, MS, N, 1);
        g.B[0] = mref(S, s2, MS, N, 1);
        g.C = W;
        CRW_TRY(run_gemm(g, B, stream, &tc));
    }
    // 5. loss rows (model.py:395-397) and dW in place
    const int need_grad = p.grad != nullptr;
    const float cgrad = 1.0f / ((float)t2 * (float)B * (float)N);
    const int64_t lrows = (int64_t)B * t2 * N;
    if (p.chains_out) cudaMemcpyAsync(p.chains_out, W, sizeof(float) * (size_t)B * s2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    const float* Wt = p.ts_target;
    CRW_LAUNCH(loss_rows_kernel, rows_grid(lrows), 256, 0, stream, W, rowloss, rowacc, lrows, N, cgrad, need_grad, Wt, rowts, p.ts_alpha);
    CRW_TRY(check_launch("loss_rows"));
    CRW_LAUNCH(loss_reduce_kernel, t2, 256, 0, stream, rowloss, rowacc, p.xent, p.acc, B, t2, N, (const unsigned*)p.ws_tc, p.ws_counter,
               Wt ? (const float*)rowts : (const float*)nullptr, p.ts_xent, p.ts_alpha);
    CRW_TRY(check_launch("loss_reduce"));
    if (!need_grad) return CRW_OK;

    // 6. reverse sweep: gP_j = dW_j S_j^T + gP_{j+1} X_{j+1}^T ; gS_j = P_j^T dW_j + Y_{j+1}^T gS_{j+1}
    for (int j = T - 2; j >= 0; --j) {
        GemmArgs g[2] = {};                                  // gP_j and gS_j have the same shape and term count: one launch
        for (int i = 0; i < 2; ++i) { g[i].M = N; g[i].N = N; g[i].nj = 1; g[i].accumulate = 0; g[i].csb = s1; g[i].csj = 0; g[i].ldc = N; }
        int t = 0;
        if (j >= 1) {
            g[0].A[t] = mref(W + (j - 1) * MS, s2, 0, N, 1);
            g[0].B[t] = mref(S + (j - 1) * MS, s2, 0, 1, N);
            g[1].A[t] = mref(P + (j - 1) * MS, s2, 0, 1, N);
            g[1].B[t] = mref(W + (j - 1) * MS, s2, 0, N, 1);
            g[0].K[t] = g[1].K[t] = N;
            ++t;
        }
        if (j + 1 <= T - 2) {
            g[0].A[t] = mref(gP + (j + 1) * MS, s1, 0, N, 1);
            g[0].B[t] = mref(X + (j + 1) * MS, s1, 0, 1, N);
            g[1].A[t] = mref(Y + (j + 1) * MS, s1, 0, 1, N);
            g[1].B[t] = mref(gS + (j + 1) * MS, s1, 0, N, 1);
            g[0].K[t] = g[1].K[t] = N;
            ++t;
        }
        g[0].nterms = g[1].nterms = t;
        g[0].C = gP + j * MS;
        g[1].C = gS + j * MS;
        CRW_TRY(run_gemms(g, 2, B, stream, &tc));
    }
    // 7. dX_j = P_{j-1}^T gP_j, dY_j = gS_j S_{j-1}^T (j >= 1); dX_0 = gP_0, dY_0 = gS_0
    {
        const int cg = (int)((MS * B + 255) / 256 < 1184 ? (MS * B + 255) / 256 : 1184);
        CRW_LAUNCH(copy_mats_kernel, cg, 256, 0, stream, dX, gP, s1, s1, MS, B);
        CRW_LAUNCH(copy_mats_kernel, cg, 256, 0, stream, dY, gS, s1, s1, MS, B);
        CRW_TRY(check_launch("copy_mats"));
        GemmArgs g[2] = {};
        g[0].nterms = 1; g[0].K[0] = N; g[0].M = N; g[0].N = N; g[0].accumulate = 0; g[0].csb = s1; g[0].ldc = N;
        // j = 1 (P_0 = X_0, S_0 = Y_0 live in the transition stacks)
        g[0].nj = 1; g[0].csj = 0;
        g[1] = g[0];
        g[0].A[0] = mref(X, s1, 0, 1, N); g[0].B[0] = mref(gP + MS, s1, 0, N, 1); g[0].C = dX + MS;
        g[1].A[0] = mref(gS + MS, s1, 0, N, 1); g[1].B[0] = mref(Y, s1, 0, 1, N); g[1].C = dY + MS;
        CRW_TRY(run_gemms(g, 2, B, stream, &tc));
        if (T - 2 >= 2) {
            g[0].nj = g[1].nj = T - 3; g[0].csj = g[1].csj = MS;
            g[0].A[0] = mref(P, s2, MS, 1, N); g[0].B[0] = mref(gP + 2 * MS, s1, MS, N, 1); g[0].C = dX + 2 * MS;
            g[1].A[0] = mref(gS + 2 * MS, s1, MS, N, 1); g[1].B[0] = mref(S, s2, MS, 1, N); g[1].C = dY + 2 * MS;
            CRW_TRY(run_gemms(g, 2, B, stream, &tc));
        }
    }
    // 8. transition-matrix backward, in place over the raw affinities: A <- Z (rows of F), AT <- Z2 (rows of G)
    {
        const int64_t r = (int64_t)B * t1 * N;
        const unsigned char* cF = p.rate > 0.f ? codesF : nullptr;
        const unsigned char* cG = p.rate > 0.f ? codesG : nullptr;
        CRW_LAUNCH(stoch_bwd_rows_kernel, rows_grid(r), 256, 0, stream, A, (const float*)F, (const float*)(flip ? dY : dX),
                   (const float*)s12, cF, 1u, r, N, p.tau, softmax);
        CRW_LAUNCH(stoch_bwd_rows_kernel, rows_grid(r), 256, 0, stream, AT, (const float*)G, (const float*)(flip ? dX : dY),
                   (const float*)s21, cG, 3u, r, N, p.tau, softmax);
        CRW_TRY(check_launch("stoch_bwd_rows"));
    }
    // 9. dQ_i += Z_i Q_{i+1} + Z2_i^T Q_{i+1} ; dQ_{i+1} += Z_i^T Q_i + Z2_i Q_i
    cudaMemsetAsync(p.grad, 0, sizeof(float) * (size_t)B * N * T * D, (cudaStream_t)stream);
    {
        // the two products write interleaved frames of dQ (i and i + 1): they stay two launches, the second accumulating
        // onto the first's rows
        GemmArgs g{};
        g.nterms = 2; g.K[0] = g.K[1] = N; g.M = N; g.N = D; g.nj = t1; g.accumulate = 1; g.csb = cs; g.csj = D; g.ldc = gs;
        g.A[0] = mref(A, s1, MS, N, 1);  g.B[0] = mref(p.q + D, cs, D, gs, 1);
        g.A[1] = mref(AT, s1, MS, 1, N); g.B[1] = mref(p.q + D, cs, D, gs, 1);
        g.C = p.grad;
        CRW_TRY(run_gemm(g, B, stream, &tc));
        g.A[0] = mref(A, s1, MS, 1, N);  g.B[0] = mref(p.q, cs, D, gs, 1);
        g.A[1] = mref(AT, s1, MS, N, 1); g.B[1] = mref(p.q, cs, D, gs, 1);
        g.C = p.grad + D;
        CRW_TRY(run_gemm(g, B, stream, &tc));
    }
    // 10. normalisation backward
    CRW_LAUNCH(gnorm_bwd_kernel, rows_grid(rows), 256, 0, stream, (const float*)p.q, p.grad, (const float*)invn,
               (const float*)nrm, rows, D);
    return check_launch("gnorm_bwd");
}

}  // namespace crw

// ---- standalone operator entry points (the reference's CRW.affinity / CRW.stoch_mat methods) ---------------------
using namespace crw;

extern "C" int crw_affinity(const float* x1, const float* x2, int BT, int N1, int N2, int D, float* out, crw_stream_t stream) {
    if (BT < 0 || N1 <= 0 || N2 <= 0 || D <= 0) { set_error("affinity: bad shape"); return CRW_ERR_SHAPE; }
    GemmArgs g{};
    g.nterms = 1; g.K[0] = D; g.M = N1; g.N = N2; g.nj = 1; g.accumulate = 0;
    g.A[0] = mref(x1, (int64_t)N1 * D, 0, D, 1);
    g.B[0] = mref(x2, (int64_t)N2 * D, 0, 1, D);
    g.C = out; g.csb = (int64_t)N1 * N2; g.csj = 0; g.ldc = N2;
    // grid.z is limited to 65535: split the batch
    for (int b0 = 0; b0 < BT; b0 += 32768) {
        const int nb = BT - b0 < 32768 ? BT - b0 : 32768;
        GemmArgs h = g;
        h.A[0].p = x1 + (int64_t)b0 * N1 * D;
        h.B[0].p = x2 + (int64_t)b0 * N2 * D;
        h.C = out + (int64_t)b0 * N1 * N2;
        int e = run_gemm(h, nb, stream);
        if (e != CRW_OK) return e;
    }
    return CRW_OK;
}

extern "C" int crw_stoch_mat(float* A, const float* drop_uniform, float rate, float temperature, unsigned flags,
                             int64_t R, int N, int M, float* out, crw_stream_t stream) {
    if (R < 0 || N <= 0 || M <= 0 || !(temperature > 0.f)) { set_error("stoch_mat: bad shape / temperature"); return CRW_ERR_SHAPE; }
    StochArgs s{};
    s.A = A; s.out = out; s.denom = nullptr; s.codes = nullptr;
    s.u1 = drop_uniform; s.u2 = nullptr;
    s.mode = 0; s.softmax = (flags & CRW_WALK_SOFTMAX) ? 1 : 0;
    s.rows = R * N; s.N = N; s.M = M; s.T = 2; s.B = 1;
    s.tau = temperature; s.rate = drop_uniform ? rate : 0.f;
    if (s.rows == 0) return CRW_OK;
    CRW_LAUNCH(stoch_rows_kernel, rows_grid(s.rows), 256, 0, stream, s);
    return check_launch("stoch_mat");
}

// Backward of crw_stoch_mat (model.py:74-90 / utils/__init__.py:414-422 under autograd): A is the tensor AFTER the in-place
// dropout (dropped entries hold -1e20 and receive no gradient, as index_put_ cuts it in the reference), y the forward output.
//   softmax:      dA_i = y_i (g_i - sum_j g_j y_j) / tau
//   ZeroSoftmax:  y_i = e_i / (sum e + eps), e_i = (exp(A_i / tau) - 1)^2:  dA_i = 2 (exp(z_i) - 1) exp(z_i) / (tau S) (g_i - sum_j g_j y_j)
namespace crw {
__global__ void __launch_bounds__(256) stoch_rows_bwd_kernel(const float* __restrict__ A, const float* __restrict__ y, const float* __restrict__ g,
                                                             float* __restrict__ gA, int64_t rows, int M, float tau, int softmax) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += nw) {
        const float* a = A + row * M;
        const float* yr = y + row * M;
        const float* gr = g + row * M;
        float dot = 0.f, S = 0.f;
        for (int m = lane; m < M; m += 32) {
            dot = fmaf(gr[m], yr[m], dot);
            if (!softmax) { const float E = expf(a[m] / tau) - 1.0f; S = fmaf(E, E, S); }
        }
        dot = warp_sum(dot);
        S = warp_sum(S) + kEpsZs;
        for (int m = lane; m < M; m += 32) {
            float d;
            if (softmax) d = yr[m] * (gr[m] - dot) / tau;
            else {
                const float ez = expf(a[m] / tau);
                d = 2.0f * (ez - 1.0f) * ez / (tau * S) * (gr[m] - dot);
            }
            gA[row * M + m] = d;
        }
    }
}
}  // namespace crw

extern "C" int crw_stoch_mat_bwd(const float* A, const float* out, const float* grad_out, float temperature, unsigned flags,
                                 int64_t R, int N, int M, float* grad_A, crw_stream_t stream) {
    if (R < 0 || N <= 0 || M <= 0 || !(temperature > 0.f)) { set_error("stoch_mat_bwd: bad shape / temperature"); return CRW_ERR_SHAPE; }
    const int64_t rows = R * N;
    if (rows == 0) return CRW_OK;
    CRW_LAUNCH(stoch_rows_bwd_kernel, rows_grid(rows), 256, 0, stream, A, out, grad_out, grad_A, rows, M, temperature,
               (flags & CRW_WALK_SOFTMAX) ? 1 : 0);
    return check_launch("stoch_mat_bwd");
}

// ---- head weight gradient: dW (D,C) = g^T (D,R) x (R,C), split over R so the small output still fills the GPU ----
namespace crw {
// out = beta * out + alpha * sum_j part_j  (alpha = 1, beta = 0: plain reduction, bit-identical to the sum alone)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int S,
                                                            float alpha, float beta) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < S; ++j) s += part[(int64_t)j * n + e];          // fixed order: deterministic
        out[e] = beta == 0.f ? (alpha == 1.f ? s : alpha * s) : fmaf(alpha, s, beta * out[e]);
    }
}
static int wgrad_splits(int64_t R, int D, int C) {
    const int tiles = ((D + BM - 1) / BM) * ((C + BN - 1) / BN);
    int S = (148 * 4 + tiles - 1) / tiles;                                   // ~4 CTAs per SM
    const int maxS = (int)((R + 63) / 64);
    if (S > maxS) S = maxS;
    return S < 1 ? 1 : S;
}
}  // namespace crw

namespace crw {
// split-K factor for the tensor-core weight gradient: equal slices (TMA needs one slice extent), about one wave of CTAs;
// 0 when no divisor of R fits or the shape is not TMA-addressable
static int wgrad_tc_splits(int64_t R, int D, int C) {
    if (D < 64 || C < 64 || (D & 3) || (C & 3) || R < 256) return 0;
    const int tiles = ((D + 127) / 128) * ((C + 127) / 128);
    const int target = 148 / tiles > 1 ? 148 / tiles : 1;
    int best = 0;
    for (int S = 1; S <= 4 * target && S <= 512; ++S) {
        if (R % S != 0 || R / S < 32) continue;
        if (best == 0 || abs(S - target) < abs(best - target)) best = S;
    }
    return (best >= target / 4 && best >= 1) ? best : 0;
}
}  // namespace crw

extern "C" size_t crw_head_wgrad_workspace_bytes(int64_t R, int D, int C) {
    if (R <= 0 || D <= 0 || C <= 0) return 0;
    const int S = wgrad_splits(R, D, C), St = wgrad_tc_splits(R, D, C);
    return sizeof(float) * (size_t)(S > St ? S : St) * D * C + 256;          // + one error word for the tensor-core pipeline
}

extern "C" int crw_head_wgrad(const float* grad_out, const float* x, float* dW, int64_t R, int D, int C, void* workspace,
                              size_t workspace_bytes, crw_stream_t stream) {
    return crw_head_wgrad_axpby(grad_out, x, dW, R, D, C, 1.0f, 0.0f, workspace, workspace_bytes, stream);
}

extern "C" int crw_head_wgrad_axpby(const float* grad_out, const float* x, float* dW, int64_t R, int D, int C, float alpha, float beta,
                                    void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    if (R <= 0 || D <= 0 || C <= 0) { set_error("head_wgrad: bad shape"); return CRW_ERR_SHAPE; }
    if (!workspace || workspace_bytes < crw_head_wgrad_workspace_bytes(R, D, C)) { set_error("head_wgrad: workspace too small"); return CRW_ERR_SHAPE; }
    const int64_t n = (int64_t)D * C;
    // tensor cores: dW = sum over S equal row slices of g_s^T x_s, both operands read in place as MN-major tf32 tiles
    const int St = wgrad_tc_splits(R, D, C);
    if (St > 0) {
        const int64_t Ks = R / St;
        TcGemmCall c{};
        c.ngroups = 1; c.nterms = 1; c.K[0] = (int)Ks; c.M = D; c.N = C; c.nb = St; c.nj = 1;
        c.grp[0].A[0] = TcOperand{grad_out, Ks * D, 0, 1, D};         // A(r = d, k = row) = g[row * D + d]
        c.grp[0].B[0] = TcOperand{x, Ks * C, 0, C, 1};                // B(k = row, c) = x[row * C + c]
        c.grp[0].C = (float*)workspace; c.grp[0].csb = n; c.grp[0].csj = 0; c.grp[0].ldc = C; c.grp[0].accumulate = 0;
        if (gemm_tf32_eligible(c)) {
            unsigned* err = (unsigned*)((char*)workspace + sizeof(float) * (size_t)St * n);
            int e = gemm_tf32_run(c, err, stream);
            if (e != CRW_OK) return e;
            CRW_LAUNCH(splitk_reduce_kernel, (int)((n + 255) / 256), 256, 0, stream, (const float*)workspace, dW, n, St, alpha, beta);
            return check_launch("head_wgrad_reduce");
        }
    }
    const int S = wgrad_splits(R, D, C);
    const int KS = (int)((R + S - 1) / S);
    GemmArgs g{};
    g.nterms = 1; g.K[0] = KS; g.ktot = (int)R; g.M = D; g.N = C; g.nj = S; g.accumulate = 0;
    g.A[0] = mref(grad_out, 0, (int64_t)KS * D, 1, D);          // A(r = d, k = row): g[row*D + d]
    g.B[0] = mref(x, 0, (int64_t)KS * C, C, 1);                 // B(k = row, c)
    g.C = (float*)workspace; g.csb = 0; g.csj = (int64_t)D * C; g.ldc = C;
    int e = run_gemm(g, 1, stream);
    if (e != CRW_OK) return e;
    CRW_LAUNCH(splitk_reduce_kernel, (int)((n + 255) / 256), 256, 0, stream, (const float*)workspace, dW, n, S, alpha, beta);
    return check_launch("head_wgrad_reduce");
}

// head forward / input gradient (model.py:117, nn.Linear(bias=False)) on the fused tf32 GEMM.  CRW_ERR_UNSUPPORTED when the
// shape is not TMA-addressable (the caller then uses a library GEMM).
extern "C" int crw_head_fwd(const float* x, const float* weight, float* out, int64_t R, int D, int C, unsigned* err_word, crw_stream_t stream) {
    if (R <= 0 || D <= 0 || C <= 0 || R > 0x7fffffff) { set_error("head_fwd: bad shape"); return CRW_ERR_SHAPE; }
    TcGemmCall c{};
    c.ngroups = 1; c.nterms = 1; c.K[0] = C; c.M = (int)R; c.N = D; c.nb = 1; c.nj = 1;
    c.grp[0].A[0] = TcOperand{x, 0, 0, C, 1};                        // x (R, C)
    c.grp[0].B[0] = TcOperand{weight, 0, 0, 1, C};                   // B(k, d) = weight[d * C + k]
    c.grp[0].C = out; c.grp[0].csb = 0; c.grp[0].csj = 0; c.grp[0].ldc = D; c.grp[0].accumulate = 0;
    if (!gemm_tf32_eligible(c)) { set_error("head_fwd: shape not addressable by the tensor-core path"); return CRW_ERR_UNSUPPORTED; }
    return gemm_tf32_run(c, err_word, stream);
}

// Split-K variant for short row counts (a micro-batch of 5 clips is 980 rows = 8 row tiles: one K loop of 16 chunks per CTA is a
// 20 us latency chain): `splits` slices of K = C / splits channels as a batched product into the workspace, summed in fixed order.
extern "C" size_t crw_head_fwd_splitk_workspace_bytes(int64_t R, int D, int splits) { return sizeof(float) * (size_t)R * D * (splits > 1 ? splits : 1) + 256; }

extern "C" int crw_head_fwd_splitk(const float* x, const float* weight, float* out, int64_t R, int D, int C, int splits, void* workspace,
                                   size_t workspace_bytes, unsigned* err_word, crw_stream_t stream) {
    if (R <= 0 || D <= 0 || C <= 0 || R > 0x7fffffff) { set_error("head_fwd: bad shape"); return CRW_ERR_SHAPE; }
    if (splits <= 1) return crw_head_fwd(x, weight, out, R, D, C, err_word, stream);
    if (C % splits != 0 || !workspace || workspace_bytes < crw_head_fwd_splitk_workspace_bytes(R, D, splits)) {
        set_error("head_fwd_splitk: C %% splits != 0 or workspace too small");
        return CRW_ERR_SHAPE;
    }
    const int Ks = C / splits;
    const int64_t n = R * D;
    TcGemmCall c{};
    c.ngroups = 1; c.nterms = 1; c.K[0] = Ks; c.M = (int)R; c.N = D; c.nb = splits; c.nj = 1;
    c.grp[0].A[0] = TcOperand{x, Ks, 0, C, 1};                       // slice z: channels [z Ks, (z+1) Ks) of x (R, C)
    c.grp[0].B[0] = TcOperand{weight, Ks, 0, 1, C};                  // B(k, d) = weight[d * C + z Ks + k]
    c.grp[0].C = (float*)workspace; c.grp[0].csb = n; c.grp[0].csj = 0; c.grp[0].ldc = D; c.grp[0].accumulate = 0;
    if (!gemm_tf32_eligible(c)) { set_error("head_fwd_splitk: shape not addressable by the tensor-core path"); return CRW_ERR_UNSUPPORTED; }
    int e = gemm_tf32_run(c, err_word, stream);
    if (e != CRW_OK) return e;
    CRW_LAUNCH(splitk_reduce_kernel, (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8), 256, 0, stream, (const float*)workspace, out, n, splits, 1.0f, 0.0f);
    return check_launch("head_fwd_splitk_reduce");
}

extern "C" int crw_head_dgrad(const float* grad_out, const float* weight, float* grad_x, int64_t R, int D, int C, unsigned* err_word,
                              crw_stream_t stream) {
    if (R <= 0 || D <= 0 || C <= 0 || R > 0x7fffffff) { set_error("head_dgrad: bad shape"); return CRW_ERR_SHAPE; }
    TcGemmCall c{};
    c.ngroups = 1; c.nterms = 1; c.K[0] = D; c.M = (int)R; c.N = C; c.nb = 1; c.nj = 1;
    c.grp[0].A[0] = TcOperand{grad_out, 0, 0, D, 1};                 // g (R, D)
    c.grp[0].B[0] = TcOperand{weight, 0, 0, C, 1};                   // B(k = d, c) = weight[d * C + c]
    c.grp[0].C = grad_x; c.grp[0].csb = 0; c.grp[0].csj = 0; c.grp[0].ldc = C; c.grp[0].accumulate = 0;
    if (!gemm_tf32_eligible(c)) { set_error("head_dgrad: shape not addressable by the tensor-core path"); return CRW_ERR_UNSUPPORTED; }
    return gemm_tf32_run(c, err_word, stream);
}

extern "C" int crw_l2norm_fwd(const float* f, float* q, float* inv_norm, float* norm, int64_t rows, int D, crw_stream_t stream) {
    if (rows < 0 || D <= 0) { set_error("l2norm_fwd: bad shape"); return CRW_ERR_SHAPE; }
    if (rows == 0) return CRW_OK;
    CRW_LAUNCH(gnorm_kernel, rows_grid(rows), 256, 0, stream, f, q, inv_norm, norm, rows, D);
    return check_launch("l2norm_fwd");
}

extern "C" int crw_l2norm_bwd(const float* q, float* grad_inout, const float* inv_norm, const float* norm, int64_t rows, int D,
                              crw_stream_t stream) {
    if (rows < 0 || D <= 0) { set_error("l2norm_bwd: bad shape"); return CRW_ERR_SHAPE; }
    if (rows == 0) return CRW_OK;
    CRW_LAUNCH(gnorm_bwd_kernel, rows_grid(rows), 256, 0, stream, q, grad_inout, inv_norm, norm, rows, D);
    return check_launch("l2norm_bwd");
}

namespace crw {
__global__ void __launch_bounds__(256) philox_uniform_kernel(float* out, int64_t n, uint64_t seed, uint64_t offset, uint32_t threads) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = torch_uniform(seed, offset, threads, (uint64_t)e);
}
}  // namespace crw

extern "C" int crw_philox_uniform(float* out, int64_t n, uint64_t seed, uint64_t offset, uint32_t philox_threads,
                                  crw_stream_t stream) {
    if (n < 0 || philox_threads == 0) { set_error("philox_uniform: bad arguments"); return CRW_ERR_SHAPE; }
    if (n == 0) return CRW_OK;
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    CRW_LAUNCH(philox_uniform_kernel, grid, 256, 0, stream, out, n, seed, offset, philox_threads);
    return check_launch("philox_uniform");
}
