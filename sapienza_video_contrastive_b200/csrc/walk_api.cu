// walk_api.cu - C ABI of the walk (include/crw_b200.h): workspace carve-up and dispatch between the single-CTA
// fused kernel (walk_fused.cu) and the batched multi-kernel path (walk_general.cu).
#include "walk.cuh"
#include "gemm_tc.cuh"

using namespace crw;

namespace {

struct WsLayout {
    size_t o_counter, o_clipcnt, o_partial, o_araw, o_codes, o_mats, o_stat;
    size_t o_F, o_G, o_dF, o_dG, o_s12, o_s21, o_invn, o_nrm, o_dqa, o_dqb, o_tc, tc_bytes, total;
    bool fused;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

WsLayout ws_layout(int B, int N, int T, int D, unsigned flags) {
    WsLayout w{};
    w.fused = !(flags & CRW_WALK_FORCE_GENERAL) && fused_fits(N, T, D);
    const size_t t1 = T > 1 ? T - 1 : 0, t2 = T >= 3 ? T - 2 : 0;
    const size_t mat = sizeof(float) * B * t1 * (size_t)fused_layout(N, T, D).MS;     // fused path: row stride NP
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    w.o_counter = take(256);
    w.o_clipcnt = take(sizeof(unsigned) * B * T);                  // zero-initialised region ends here (see crw_b200.h)
    w.o_partial = take(sizeof(float) * B * t2 * 2 * kChainCluster);       // per-clip (per-CTA in the clustered chain) loss sums
    w.o_araw = take(w.fused ? mat : 0);
    w.o_codes = take((size_t)B * t1 * N * N * (w.fused ? 1 : 2));
    w.o_mats = take(w.fused ? 0 : sizeof(float) * walk_general_mats_floats(B, N, T));
    w.o_stat = take(w.fused ? 0 : sizeof(float) * walk_general_stat_floats(B, N, T));
    if (w.fused) {
        w.o_F = take(mat); w.o_G = take(mat); w.o_dF = take(mat); w.o_dG = take(mat);
        w.o_s12 = take(sizeof(float) * B * t1 * N); w.o_s21 = take(sizeof(float) * B * t1 * N);
        w.o_invn = take(sizeof(float) * B * T * N); w.o_nrm = take(sizeof(float) * B * T * N);
        w.o_dqa = take(sizeof(float) * B * t1 * N * D); w.o_dqb = take(sizeof(float) * B * t1 * N * D);
    }
    // general path on large graphs: operand planes of the tensor-core GEMM (gemm_tc.cu)
    const int NK = N > D ? N : D;
    w.tc_bytes = 0;
    // measured against the SIMT GEMM (tools/bench_tc_gemm.py): 1.3-1.7x at N = 64..68 (T > 4, where the fused kernels do not
    // fit), 1.45x at 128, 2.1x at 196, 2.5x at 512, 4.4x at 1024
    const bool tc = !w.fused && !(flags & CRW_WALK_FORCE_SIMT) && t1 > 0 && (N >= 64 || (flags & CRW_WALK_FORCE_TC)) &&
                    gemm_tc_eligible(N, N < D ? N : D, N < D ? N : D, NK);
    if (tc) {    // the largest of the calls launch_walk_general makes (groups x batch, K-concatenated terms)
        const int Np = (N + 7) & ~7, Dp = (D + 7) & ~7, b = B, m1 = (int)t1, m2 = (int)t2;
        const size_t calls[6] = {gemm_tc_workspace_bytes(N, N, Dp, 2 * b * m1),          // affinities, both orientations
                                 gemm_tc_workspace_bytes(N, N, Np, 2 * b),               // one chain level
                                 gemm_tc_workspace_bytes(N, N, Np, b * m2),              // W_j
                                 gemm_tc_workspace_bytes(N, N, 2 * Np, 2 * b),           // one reverse-sweep level, two terms
                                 gemm_tc_workspace_bytes(N, N, Np, 2 * b * (m2 > 1 ? m2 - 1 : 1)),   // dX_j, dY_j
                                 gemm_tc_workspace_bytes(N, D, 2 * Np, b * m1)};         // dQ
        for (size_t c : calls) w.tc_bytes = c > w.tc_bytes ? c : w.tc_bytes;
    } else w.tc_bytes = 0;
    w.o_tc = take(w.tc_bytes);
    w.total = o;
    return w;
}

}  // namespace

extern "C" size_t crw_walk_workspace_bytes(int B, int N, int T, int D, unsigned flags) {
    if (B <= 0 || N <= 0 || T <= 0 || D <= 0) return 0;
    return ws_layout(B, N, T, D, flags).total;
}

namespace {
struct TsArgs {
    const float* teacher_chains;
    float alpha;
    float* ts_xent;
    float* chains_out;
};

int walk_run(const float* feats, int B, int N, int T, int D, float temperature, float rate,
             const float* u12, const float* u21p, uint64_t philox_seed, uint64_t philox_offset,
             uint32_t philox_threads, uint64_t* philox_state_dev, unsigned flags, float* q, float* xent, float* acc,
             float* grad_feats, void* workspace, size_t workspace_bytes, crw_stream_t stream, const TsArgs* ts);
}  // namespace

extern "C" int crw_walk_fwd_bwd(const float* feats, int B, int N, int T, int D, float temperature, float rate,
                                const float* u12, const float* u21p, uint64_t philox_seed, uint64_t philox_offset,
                                uint32_t philox_threads, uint64_t* philox_state_dev, unsigned flags, float* q, float* xent, float* acc,
                                float* grad_feats, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    return walk_run(feats, B, N, T, D, temperature, rate, u12, u21p, philox_seed, philox_offset, philox_threads, philox_state_dev, flags,
                    q, xent, acc, grad_feats, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int crw_walk_ts_fwd_bwd(const float* feats, int B, int N, int T, int D, float temperature, float rate,
                                   const float* u12, const float* u21p, uint64_t philox_seed, uint64_t philox_offset,
                                   uint32_t philox_threads, uint64_t* philox_state_dev, unsigned flags,
                                   const float* teacher_chains, float alpha, float* chains_out,
                                   float* q, float* xent, float* ts_xent, float* acc,
                                   float* grad_feats, void* workspace, size_t workspace_bytes, crw_stream_t stream) {
    if (teacher_chains && (!ts_xent || !(alpha >= 0.f && alpha <= 1.f))) {
        set_error("walk_ts: a teacher needs ts_xent and 0 <= alpha <= 1 (got %g)", (double)alpha);
        return CRW_ERR_SHAPE;
    }
    const TsArgs ts{teacher_chains, alpha, ts_xent, chains_out};
    return walk_run(feats, B, N, T, D, temperature, rate, u12, u21p, philox_seed, philox_offset, philox_threads, philox_state_dev,
                    flags | CRW_WALK_FORCE_GENERAL, q, xent, acc, grad_feats, workspace, workspace_bytes, stream, &ts);
}

namespace {
int walk_run(const float* feats, int B, int N, int T, int D, float temperature, float rate,
             const float* u12, const float* u21p, uint64_t philox_seed, uint64_t philox_offset,
             uint32_t philox_threads, uint64_t* philox_state_dev, unsigned flags, float* q, float* xent, float* acc,
             float* grad_feats, void* workspace, size_t workspace_bytes, crw_stream_t stream, const TsArgs* ts) {
    if (B <= 0 || N <= 0 || T <= 0 || D <= 0) { set_error("walk: bad shape B=%d N=%d T=%d D=%d", B, N, T, D); return CRW_ERR_SHAPE; }
    if (!(temperature > 0.f)) { set_error("walk: temperature must be > 0"); return CRW_ERR_SHAPE; }
    if ((u12 == nullptr) != (u21p == nullptr)) { set_error("walk: u12 and u21p must both be given or both be NULL"); return CRW_ERR_SHAPE; }
    if (rate > 0.f && !u12 && philox_threads == 0) { set_error("walk: in-kernel dropout needs philox_threads"); return CRW_ERR_SHAPE; }
    const WsLayout w = ws_layout(B, N, T, D, flags);
    if (workspace_bytes < w.total || !workspace) {
        set_error("walk: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return CRW_ERR_SHAPE;
    }
    if ((((uintptr_t)feats | (uintptr_t)q | (uintptr_t)grad_feats) & 15) != 0 || D % 4 != 0) {
        if (w.fused) { set_error("walk: fused path needs 16-byte aligned feats/q/grad and D %% 4 == 0"); return CRW_ERR_UNSUPPORTED; }
    }
    char* ws = (char*)workspace;
    WalkParams p{};
    p.feats = feats; p.q = q; p.xent = xent; p.acc = acc; p.grad = grad_feats;
    p.u12 = u12; p.u21p = u21p;
    p.dev_state = (rate > 0.f && !u12) ? philox_state_dev : nullptr;
    p.seed = philox_seed; p.offset = philox_offset; p.pthreads = philox_threads ? philox_threads : 256;
    // torch advances the Philox offset by 4 * ceil(numel / (threads * 4)) per rand call (DistributionTemplates.h)
    const int64_t numel = (int64_t)B * N * N;
    p.pinc = (uint32_t)(4 * ((numel - 1) / ((int64_t)p.pthreads * 4) + 1));
    p.B = B; p.N = N; p.T = T; p.D = D; p.tau = temperature; p.rate = rate; p.flags = flags;
    p.ws_counter = (unsigned*)(ws + w.o_counter);
    p.ws_partial = (float*)(ws + w.o_partial);
    p.ws_araw = (float*)(ws + w.o_araw);
    p.ws_codes = (unsigned char*)(ws + w.o_codes);
    p.ws_mats = (float*)(ws + w.o_mats);
    p.ws_stat = (float*)(ws + w.o_stat);
    p.ws_clipcnt = (unsigned*)(ws + w.o_clipcnt);
    p.ws_tc = w.tc_bytes ? (void*)(ws + w.o_tc) : nullptr;
    p.ws_tc_bytes = w.tc_bytes;
    if (w.fused) {
        p.ws_F = (float*)(ws + w.o_F); p.ws_G = (float*)(ws + w.o_G);
        p.ws_dF = (float*)(ws + w.o_dF); p.ws_dG = (float*)(ws + w.o_dG);
        p.ws_s12 = (float*)(ws + w.o_s12); p.ws_s21 = (float*)(ws + w.o_s21);
        p.ws_invn = (float*)(ws + w.o_invn); p.ws_nrm = (float*)(ws + w.o_nrm);
        p.ws_dqa = (float*)(ws + w.o_dqa); p.ws_dqb = (float*)(ws + w.o_dqb);
    }
    if (ts) {
        p.ts_target = ts->teacher_chains; p.ts_alpha = ts->alpha; p.ts_xent = ts->ts_xent; p.chains_out = ts->chains_out;
    }
    return w.fused ? launch_walk_fused(p, stream) : launch_walk_general(p, stream);
}
}  // namespace
