/* crw_b200.h - C ABI of libcrw_b200.so (sm_100a only).
 *
 * The reference (paolomandica/sapienza-video-contrastive) has no FFI layer: its hot path is a chain of
 * PyTorch ATen calls inside code/model.py and code/utils/test_utils.py.  This header is the boundary a
 * binding for that path targets instead.  Every entry point cites the reference lines it replaces
 * (paths relative to the reference root).  Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name says host;
 *   - the caller owns every buffer, including workspaces (query the size with *_workspace_bytes);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises or allocates - with ONE
 *     documented exception, crw_sinkhorn_knopp, whose stop rule the reference evaluates on the host once per sweep;
 *   - a tensor-core pipeline that stalls (bounded mbarrier waits, seconds) raises its error word and TRAPS: the caller sees
 *     a CUDA error at its next API call, never silently wrong output;
 *   - return value: CRW_OK or a negative error; crw_last_error() gives a thread-local message;
 *   - fp32 arithmetic throughout; indices int64.
 */
#ifndef CRW_B200_H
#define CRW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRW_OK 0
#define CRW_ERR_SHAPE (-1)
#define CRW_ERR_UNSUPPORTED (-2)
#define CRW_ERR_CUDA (-3)

/* walk flags */
#define CRW_WALK_SOFTMAX 1u   /* F.softmax rows (teacherstudent.py:80) instead of ZeroSoftmax (model.py:90) */
#define CRW_WALK_FLIP 2u      /* args.flip: reversed product order, model.py:380-382 */
#define CRW_WALK_FORCE_GENERAL 4u /* skip the fused small-graph kernels even when the clip fits shared memory */
#define CRW_WALK_FORCE_SIMT 8u    /* large-graph path: exact-fp32 SIMT GEMMs instead of the tcgen05 hi/lo-split GEMM */
#define CRW_WALK_FORCE_TC 16u     /* large-graph path: tcgen05 GEMM wherever the shapes allow it (N, D >= 64), not only above the measured crossover */
#define CRW_WALK_NO_TF32 64u      /* large-graph path: always the fp16 hi/lo-split GEMM, never the fused kind::tf32 one */
#define CRW_WALK_NO_CLUSTER 32u   /* small-graph path: one CTA per clip for the chain instead of a 4-CTA cluster */
#define CRW_LP_FORCE_SIMT 1u      /* label propagation: exact-fp32 SIMT scores instead of the tcgen05 kernel */
#define CRW_LP_EXACT_ONLY 2u      /* label propagation, tensor-core path: skip the fp16 pre-ranking pass, every tile on the 3-MMA fp32-faithful pass */

typedef void* crw_stream_t;

int crw_version(void);
const char* crw_last_error(void);

/* ---- a1: patch mean pooling, model.py:116  (feats = maps.sum(-1).sum(-1) / (H*W)) -------------------
 * maps (rows, hw) contiguous fp32 with rows = B*N*C*T; pooled (rows).  bwd broadcasts g/hw. */
int crw_pool_patch_fwd(const float* maps, float* pooled, int64_t rows, int hw, crw_stream_t stream);
int crw_pool_patch_bwd(const float* grad_pooled, float* grad_maps, int64_t rows, int hw, crw_stream_t stream);

/* The same two kernels confined to `sms` streaming multiprocessors (one 1024-thread CTA per SM, placed alone on its SM by a
 * shared-memory reservation), so that kernels of ANOTHER stream - the walk of a different micro-batch, whose CTAs need a
 * whole SM - run beside the HBM-bound pooling pass instead of behind it (sapienza_video_contrastive_b200/pipeline.py).
 * sms <= 0, an unaligned base or hw not in {16, 32, 64}: identical to the unrestricted entry points. */
int crw_pool_patch_fwd_sm(const float* maps, float* pooled, int64_t rows, int hw, int sms, crw_stream_t stream);
int crw_pool_patch_bwd_sm(const float* gpooled, float* gmaps, int64_t rows, int hw, int sms, crw_stream_t stream);
/* ... with the gradient scaled on the way: gmaps = gpooled * scale / hw (evaluated as gpooled / (hw / scale)), scale > 0 -
 * the 1 / n of "mean over the clips of all n micro-batches". */
int crw_pool_patch_bwd_scaled(const float* gpooled, float* gmaps, int64_t rows, int hw, float scale, int sms, crw_stream_t stream);

/* ---- a2/a3: superpixel segment-mean pooling, model.py:296-325 + utils/__init__.py:433-584 -------------
 * maps (B,C,T,Hm,Wm) contiguous; labels int64 addressed as labels[b*ls_b + t*ls_t + y*ls_y + x*ls_x]
 * (so channel 0 of the (B,T,3,h,w) mask is passed without a copy, model.py:298); h = sy*Hm, w = sx*Wm.
 * out (B,SP,T,C) (node-major, the layout the walk consumes): mean of maps[b,:,t,y/sy,x/sx] over pixels labelled s; labels outside [0,SP) ignored;
 * empty segments -> 0.  The workspace keeps the per-cell label histogram for the backward.
 * Hm*Wm a multiple of 32, C a multiple of 128 and SP <= 256 (BASELINE configs[2]) run as a tcgen05 product of the window-count
 * matrix with the map (fp32-faithful: TF32 big/small split, fp32 accumulation; summation order differs from the scatter-reduce
 * other shapes take, results agree to 1e-6); both paths are deterministic. */
size_t crw_segmean_workspace_bytes(int B, int T, int Hm, int Wm, int h, int w, int SP);
int crw_segmean_fwd(const float* maps, const int64_t* labels, int64_t ls_b, int64_t ls_t, int64_t ls_y, int64_t ls_x,
                    int B, int C, int T, int Hm, int Wm, int h, int w, int SP,
                    float* out, void* workspace, size_t workspace_bytes, crw_stream_t stream);
int crw_segmean_bwd(const float* grad_out, const void* workspace, size_t workspace_bytes,
                    int B, int C, int T, int Hm, int Wm, int h, int w, int SP,
                    float* grad_maps, crw_stream_t stream);

/* ---- f2: the same pooling with dilated superpixel masks, model.py:303-309 + utils/__init__.py:590-608 ----
 * Every label's mask is dilated by a ksize x ksize structuring element before the window counts (the reference's
 * depthwise conv2d(onehot, kernel, padding=ksize/2) > 0): a pixel then belongs to every label that has a pixel inside the
 * element centred on it, masks overlap, and sizes / window counts are those of the dilated masks.  `shape` picks the
 * element: diamond ("L1"), disc ("circle") or one row + one column ("cross").  ksize odd, <= 127; SP <= 255.
 * Same layouts as crw_segmean_*; the backward reads the workspace the forward filled. */
#define CRW_DILATE_L1 0
#define CRW_DILATE_CIRCLE 1
#define CRW_DILATE_CROSS 2
size_t crw_segmean_dilated_workspace_bytes(int B, int T, int Hm, int Wm, int h, int w, int SP);
int crw_segmean_dilated_fwd(const float* maps, const int64_t* labels, int64_t ls_b, int64_t ls_t, int64_t ls_y, int64_t ls_x,
                            int B, int C, int T, int Hm, int Wm, int h, int w, int SP, int ksize, int shape,
                            float* out, void* workspace, size_t workspace_bytes, crw_stream_t stream);
int crw_segmean_dilated_bwd(const float* grad_out, const void* workspace, size_t workspace_bytes,
                            int B, int C, int T, int Hm, int Wm, int h, int w, int SP,
                            float* grad_maps, crw_stream_t stream);

/* ---- a4: affinity, model.py:63-72  (einsum 'bctn,bctm->btnm') ----------------------------------------
 * x1, x2 node-major: x1[(b*T + t)*N1*D + n*D + d] (unit-norm or not), out (B*T, N1, N2). */
int crw_affinity(const float* x1, const float* x2, int BT, int N1, int N2, int D, float* out, crw_stream_t stream);

/* ---- a5: stoch_mat, model.py:74-90 + utils/__init__.py:414-422 ----------------------------------------
 * A (R, N, M) contiguous.  If drop_uniform != NULL, entries with u < rate are overwritten IN PLACE with
 * -1e20 (the reference mutates its argument).  out = ZeroSoftmax(A / temperature) (or softmax)
 * along the last dim. */
int crw_stoch_mat(float* A, const float* drop_uniform, float rate, float temperature, unsigned flags,
                  int64_t R, int N, int M, float* out, crw_stream_t stream);
/* Its backward (the reference's stoch_mat / ZeroSoftmax are differentiable torch ops): A as LEFT by the forward call (dropped
 * entries = -1e20: they get gradient 0), out = the forward result, grad_out -> grad_A, all (R, N, M). */
int crw_stoch_mat_bwd(const float* A, const float* out, const float* grad_out, float temperature, unsigned flags,
                      int64_t R, int N, int M, float* grad_A, crw_stream_t stream);

/* ---- a5, Sinkhorn branch: utils/__init__.py:615-641 (sinkhorn_knopp) as called by stoch_mat(do_sinkhorn=True), model.py:83-87 ----
 * A (R, N, M) contiguous, IN PLACE: optionally A <- exp(A / temperature), A <- A / sum(A) per matrix, then sweeps of
 * "L1-normalise the columns, L1-normalise the rows" (F.normalize semantics, eps 1e-12) while the unbiased standard
 * deviation of all R*M column sums exceeds tol, at least one and at most max_iter sweeps (the reference's own stop rule).
 * The stop rule is evaluated on the HOST once per sweep, as in the reference: this call synchronises `stream`.
 * *iterations (host, may be NULL) receives the number of sweeps. */
size_t crw_sinkhorn_workspace_bytes(int64_t R, int N, int M);
int crw_sinkhorn_knopp(float* A, int64_t R, int N, int M, int apply_exp, float temperature, float tol, int max_iter,
                       int* iterations, void* workspace, size_t workspace_bytes, crw_stream_t stream);

/* ---- a5/a6: the walk, model.py:366-413, forward AND backward in one call -------------------------------
 * feats (B,N,T,D) fp32, NOT yet normalised (output of selfsim_fc, model.py:117); the L2 normalisation of
 * model.py:118 is folded in.  Edge dropout: either u12/u21p (each (T-1,B,N,N), the reference's rand_like
 * draws in its order, u21p in its physical = transposed layout) or, if both are NULL and
 * rate > 0, an in-kernel Philox4x32-10 replay of torch's CUDA generator: draw j (0 <= j < 2(T-1)) uses
 * (philox_seed, philox_offset + inc*j), inc = 4*ceil(B*N*N / (4*philox_threads)), with torch's
 * element->thread mapping for `philox_threads` = grid*block threads of torch's rand kernel.
 * If philox_state_dev != NULL it points to DEVICE memory {seed, offset}: the kernel reads the state from there and
 * advances the offset by 2(T-1)*inc when it is done, so a launch captured in a CUDA graph draws fresh dropout
 * masks on every replay (philox_seed / philox_offset are then ignored).
 * The workspace must be zero-filled once after allocation; every call leaves it reusable.
 * Outputs: q (B,N,T,D) unit-norm nodes; xent (T-2 + 1): mean cross-entropies of walks i = 1..T-2
 * followed by their mean, i.e. the loss of model.py:413 (T < 3: untouched); acc (T-2) argmax accuracies; grad_feats (B,N,T,D) = d(sum_i xent_i / max(1,T-2)) / d feats, or NULL
 * to skip the backward. */
size_t crw_walk_workspace_bytes(int B, int N, int T, int D, unsigned flags);
int crw_walk_fwd_bwd(const float* feats, int B, int N, int T, int D, float temperature, float rate,
                     const float* u12, const float* u21p, uint64_t philox_seed, uint64_t philox_offset,
                     uint32_t philox_threads, uint64_t* philox_state_dev, unsigned flags, float* q, float* xent, float* acc,
                     float* grad_feats, void* workspace, size_t workspace_bytes, crw_stream_t stream);

/* ---- f3: teacher-student walk, teacherstudent.py:472-580 (CRWTeacherStudent.forward) + :270-292 (SoftCrossEntropyLoss) ----
 * The same walk with two additions, both on the chain products W_i (B,T-2,N,N) = the palindrome products `aar` (`aal`
 * with CRW_WALK_FLIP) of walks i = 1..T-2:
 *   chains_out != NULL      W_i are copied out before the loss (the TEACHER call: rate 0, grad_feats NULL);
 *   teacher_chains != NULL  the STUDENT call: ts_xent[i] = mean over rows of  -sum_m teacher[m] * log_softmax(W_i row)[m]
 *                           (the student's chain probabilities are the logits there, as in the reference), ts_xent[T-2] their
 *                           mean, xent[T-2] = alpha * mean_i xent_i + (1 - alpha) * ts_xent[T-2]  (teacherstudent.py:572-575),
 *                           and grad_feats is the gradient of THAT loss.
 * Runs the batched multi-kernel path: size the workspace with crw_walk_workspace_bytes(..., flags | CRW_WALK_FORCE_GENERAL).
 * The reference's CRWBase transition matrix is the softmax one: pass CRW_WALK_SOFTMAX for it. */
int crw_walk_ts_fwd_bwd(const float* feats, int B, int N, int T, int D, float temperature, float rate,
                        const float* u12, const float* u21p, uint64_t philox_seed, uint64_t philox_offset,
                        uint32_t philox_threads, uint64_t* philox_state_dev, unsigned flags,
                        const float* teacher_chains, float alpha, float* chains_out,
                        float* q, float* xent, float* ts_xent, float* acc,
                        float* grad_feats, void* workspace, size_t workspace_bytes, crw_stream_t stream);

/* ---- a1: the projection head, model.py:117 (nn.Linear(C_e,128,bias=False)) ------------------------------------------
 * out (R,D) = x (R,C) weight^T, grad_x (R,C) = grad_out (R,D) weight, with R = B*N*T rows, on the fused tcgen05 kind::tf32
 * GEMM (fp32-faithful: tf32 big + small operand parts, error ~2^-22 |a||b|).  err_word: one zero-initialised device word
 * that turns non-zero if the pipeline times out.  CRW_ERR_UNSUPPORTED when the shape is not TMA-addressable (C, D multiples
 * of 4, >= 64 / 32; 16-byte aligned pointers): use a library GEMM then.
 * dW (D,C) = grad_out^T (D,R) x (R,C): a tiny 128 x 512 output with a long reduction, split over R into equal slices on the
 * same tensor-core kernel (or ragged slices on the exact-fp32 SIMT kernel) and folded in a fixed order: deterministic.  The
 * workspace (crw_head_wgrad_workspace_bytes) must be zero-filled once. */
int crw_head_fwd(const float* x, const float* weight, float* out, int64_t R, int D, int C, unsigned* err_word, crw_stream_t stream);
int crw_head_dgrad(const float* grad_out, const float* weight, float* grad_x, int64_t R, int D, int C, unsigned* err_word,
                   crw_stream_t stream);
size_t crw_head_wgrad_workspace_bytes(int64_t R, int D, int C);
int crw_head_wgrad(const float* grad_out, const float* x, float* dW, int64_t R, int D, int C, void* workspace,
                   size_t workspace_bytes, crw_stream_t stream);
/* dW = beta * dW + alpha * grad_out^T x: accumulation of micro-batch contributions (pipeline.py) without a separate pass. */
int crw_head_wgrad_axpby(const float* grad_out, const float* x, float* dW, int64_t R, int D, int C, float alpha, float beta,
                         void* workspace, size_t workspace_bytes, crw_stream_t stream);
/* Head forward with the K = C contraction split into `splits` slices (a batched product into the workspace, summed in fixed
 * order): for micro-batches whose few row tiles would otherwise run one long serial K loop each.  splits <= 1 = crw_head_fwd. */
size_t crw_head_fwd_splitk_workspace_bytes(int64_t R, int D, int splits);
int crw_head_fwd_splitk(const float* x, const float* weight, float* out, int64_t R, int D, int C, int splits, void* workspace,
                        size_t workspace_bytes, unsigned* err_word, crw_stream_t stream);

/* L2 normalisation of rows (F.normalize, eps 1e-12; model.py:118,329): q = f / max(|f|, eps).  inv_norm and norm
 * (rows each) are kept for the backward, which overwrites grad in place: g <- (g - q (q.g)) * inv_norm. */
int crw_l2norm_fwd(const float* f, float* q, float* inv_norm, float* norm, int64_t rows, int D, crw_stream_t stream);
int crw_l2norm_bwd(const float* q, float* grad_inout, const float* inv_norm, const float* norm, int64_t rows, int D,
                   crw_stream_t stream);

/* Batched fp32-faithful GEMM on the tensor cores (the contraction engine of the large-graph walk, exported for
 * tests and for callers of CRW.affinity at large N): C[z] (M,N) (+)= op(A[z]) (M,K) op(B[z]) (K,N), all fp32 row-major
 * contiguous; trans_a: A stored (Z,K,M); trans_b: B stored (Z,N,K).  Operands are split into per-row-scaled fp16
 * hi/lo planes and multiplied with three tcgen05 MMAs per k-step (error ~2^-22 |a||b|).  Needs M >= 128, N >= 64,
 * K >= 64 (CRW_ERR_UNSUPPORTED otherwise).  Workspace: zero-filled once, crw_bmm_tc_workspace_bytes. */
size_t crw_bmm_tc_workspace_bytes(int Z, int M, int N, int K);
int crw_bmm_tc(const float* A, const float* B, float* C, int Z, int M, int N, int K, int trans_a, int trans_b,
               int accumulate, void* workspace, size_t workspace_bytes, crw_stream_t stream);

/* Same product on tcgen05 kind::tf32: TMA reads A and B in place (16-byte aligned bases, M/N/K strides multiples of 4
 * floats) and the kernel splits them into tf32 big + small parts itself; no workspace.  err_word: one zero-initialised
 * device word that receives a non-zero value if the pipeline times out.  CRW_ERR_UNSUPPORTED if a shape is not
 * addressable. */
int crw_bmm_tf32(const float* A, const float* B, float* C, int Z, int M, int N, int K, int trans_a, int trans_b,
                 int accumulate, unsigned* err_word, crw_stream_t stream);

/* torch-compatible uniform draw (same Philox stream as torch.rand on CUDA); used by tests to pin the
 * in-kernel replay.  out (n). */
int crw_philox_uniform(float* out, int64_t n, uint64_t seed, uint64_t offset, uint32_t philox_threads,
                       crw_stream_t stream);

/* ---- a9-a11: label-propagation affinity + top-k, utils/test_utils.py:148-179 (+ mask test.py:118-122) --
 * feats (Nf, hw, C) fp32, channel-last, already L2-normalised if the caller wants cosine affinities
 * (test.py:93).  key_frames (Nt, S) int64: context frame ids of target n (context_index_bank); the first
 * n_long slots are long-memory (unmasked), the rest are radius-restricted: key (ky,kx) is admissible for
 * query (qy,qx) iff (ky-qy)^2 + (kx-qx)^2 < radius^2 (float32 sqrt(d2) < radius in the reference).
 * radius <= 0 disables the restriction.  Target n's query frame is query_frames[n].
 * dense_mask (optional, (hw_keys, hw_queries) additive fp32 as built by test.py:118-122): when non-NULL the
 * restricted slots instead visit every key and add dense_mask[key, query] to the score - the reference's literal
 * semantics for an arbitrary mask, without the window skipping.
 * Ws (Nt,k,hw) fp32 softmax over the k best scores / temperature; Is (Nt,k,hw) int64 = slot*hw + key_pos,
 * sorted by descending score, ties broken by ascending index.
 * Two implementations with identical results.  (1) The tensor-core path (C % 64 == 0, C <= 256, k <= 12, radius <= 12, no
 * dense mask; |feats| < 6e4, csrc/lp_tc.cu): a tcgen05/TMEM/TMA kernel pre-ranks every admissible key on the fp16 "hi" part of
 * the features (one MMA per 16 channels) and keeps a 16-key shortlist per query; a second kernel re-scores the shortlist in
 * exact fp32 (sequential fmaf over the channels), ranks it and PROVES the result equal to a full exact evaluation from the
 * error bound of the pre-scores (|pre - exact| <= 1.05e-3 |q| max|k|); tiles holding a query that cannot be certified (exact
 * ties of replicated frames, vos.py:148-149; CRW_LP_EXACT_ONLY) are redone on the fp32-faithful pass (fp16 hi/lo
 * split, three MMAs per step, ~2^-22 relative) and re-ranked the same way.  (2) An exact-fp32 SIMT kernel for everything else
 * (or CRW_LP_FORCE_SIMT); both rank by the same fp32 scores, so Ws and Is agree bit for bit.  feats holds Nf frames.
 * Workspace header (32-bit words, valid once the stream has drained): [0] error (1 = internal barrier timeout - the kernel
 * also traps -, 2 = features outside the fp16 range), [1] tiles sent to the fp32-faithful pass, [4] queries not certified. */
size_t crw_lp_topk_workspace_bytes(int Nf, int Nt, int S, int h, int w, int C, int k);
int crw_lp_topk(const float* feats, int Nf, const int64_t* key_frames, const int64_t* query_frames, int Nt, int S,
                int n_long, int h, int w, int C, float radius, const float* dense_mask, float temperature, int k,
                unsigned flags, float* Ws, int64_t* Is, void* workspace, size_t workspace_bytes, crw_stream_t stream);

/* Which kernel crw_lp_topk takes for a configuration: 1 = the tcgen05 tensor-core kernel, 0 = the exact-fp32 SIMT kernel
 * (what the dispatcher of crw_lp_topk decides, given a workspace of crw_lp_topk_workspace_bytes and no CRW_LP_FORCE_SIMT).
 * Lets callers and tests assert that the path they mean to measure is the one that runs. */
int crw_lp_topk_uses_tensor_cores(int C, int k, float radius, int has_dense_mask);

/* feats (C, Nf, hw) channel-first (the encoder's layout, test.py:90-93) -> (Nf, hw, C) channel-last with
 * optional L2 normalisation over C (eps 1e-12). */
int crw_lp_prepare(const float* feats_cf, int C, int Nf, int hw, int normalize, float* feats_cl, crw_stream_t stream);

/* ---- a12: one step of the label gather, test.py:147-154 ----------------------------------------------
 * lbls (Nf, hw, L) soft labels; pred[q,l] = sum_k lbls[key_frames_n[Is[k,q] / hw], Is[k,q] % hw, l] * Ws[k,q],
 * written to lbls[out_frame] (test.py:157).  Sequential in the target index by construction. */
int crw_lp_gather(float* lbls, const int64_t* key_frames_n, const float* Ws_n, const int64_t* Is_n,
                  int hw, int L, int k, int64_t out_frame, crw_stream_t stream);

/* The whole loop test.py:145-157 of one video in one call: for t = first_target .. Nt-1, lbls[out_frame0 + t] = the gather of
 * target t (key_frames (Nt, S), Ws / Is (Nt, k, hw) as crw_lp_topk returns them): one gather launch per frame, enqueued back to
 * back in stream order (the recurrence is sequential by construction).  The caller handles what precedes first_target (test.py:158-160: the first
 * target keeps the ground truth, so first_target = 1 and out_frame0 = n_context). */
int crw_lp_gather_all(float* lbls, const int64_t* key_frames, const float* Ws, const int64_t* Is, int Nt, int S, int hw, int L,
                      int k, int first_target, int64_t out_frame0, crw_stream_t stream);

/* ---- f1 (SURVEY 8f "next"): label-map post-processing, utils/test_utils.py:85-123 (dump_predictions) + test.py:162-164 ----
 * pred (n, h, w, L) fp32 soft label maps of n target frames -> cv2.resize(pred, (W, H)) (bilinear, OpenCV's float path and
 * border rule) -> arg-max over L (first maximum wins) -> cls (n, H, W) uint8 class index and / or rgb (n, H, W, 3) uint8 =
 * palette[cls] (palette (L, 3) uint8 = np.uint8(lbl_set); NULL: rgb repeats the index).  norm_mask != 0 first applies
 * pred -= min_L; pred /= max_L per source pixel.  The upsampled (H, W, L) tensor is never materialised.  L <= 255. */
int crw_lp_upsample_argmax(const float* pred, int n, int h, int w, int L, int H, int W, int norm_mask,
                           const unsigned char* palette, unsigned char* cls, unsigned char* rgb, crw_stream_t stream);

/* test.py:162-164 (--norm_mask) as an in-place operator: maps (rows, L) fp32, every row r: maps[r] -= min(maps[r]);
 * maps[r] /= max(maps[r]) (0/0 -> NaN, as in the reference).  The evaluator applies it to the ground-truth frame 0 after the
 * first target (the reference's `pred = lbls[0]` is a view, so its in-place normalisation rewrites frame 0) and to every
 * map it returns. */
int crw_lp_minmax_normalize(float* maps, int64_t rows, int L, crw_stream_t stream);

/* ---- f1, JHMDB branch: key-point coordinates, utils/test_utils.py:60-84 (process_pose) called at test.py:171-172 ------------
 * pred (n, h, w, L) fp32 soft label maps, channel 0 = background.  For every frame and channel c = 1..L-1: the topk
 * (default 3, at most 4; clipped to h*w) largest positions, values normalised to sum 1, coords[f][0][c-1] = sum x_i v_i,
 * coords[f][1][c-1] = sum y_i v_i (fp32, in rank order), or -1, -1 when the channel is zero everywhere.
 * coords (n, 2, L-1) fp32.  Equal values rank by position (torch.topk leaves that order unspecified). */
int crw_lp_pose_coords(const float* pred, int n, int h, int w, int L, int topk, float* coords, crw_stream_t stream);

/* ---- f4: the patch-grid producer, code/utils/augs.py:59-82 (patch_grid) --------------------------------
 * frames (F, H, W, 3) uint8, DEVICE: the frames of a batch of clips after the frame-level transforms.  The (H - win) / stride + 1
 * by (W - win) / stride + 1 windows of win x win pixels (skimage view_as_windows order, augs.py:75-76) are each cropped to
 * boxes[f][p] = {top, left, height, width} (the RandomResizedCrop(win, scale (0.7, 0.9)) parameters, drawn by the caller),
 * resized to out_size x out_size with Pillow's BILINEAR rule (bit-exact on the 8-bit image), divided by 255 and normalised
 * with mean3 / std3 (HOST pointers, 3 floats each; augs.py:10-12).  out (F, P * 3, out_size, out_size) fp32 - reshaped to
 * (B, T, P * 3, out_size, out_size) it is the tensor CRW.forward takes (model.py:348-349).  win, out_size <= 64, out_size >= win. */
int crw_patch_grid(const unsigned char* frames, const int* boxes, int F, int H, int W, int win, int stride, int out_size,
                   const float* mean3, const float* std3, float* out, crw_stream_t stream);

/* ---- f4, second half: the superpixel label-map producer, code/data/superpixels.py:9-16 (compute_sp_slic) for every frame of a
 * clip (compute_mask, superpixels.py:24-63) ----------------------------------------------------------------------------------
 * video (F, 3, H, W) fp32, DEVICE (the (T, C, H, W) clip compute_mask receives; it permutes each frame to HWC on the host).
 * Per frame: cv2.normalize(0, 255, NORM_MINMAX, CV_8U) (bit-exact with OpenCV 4.x) -> skimage.segmentation.slic(img,
 * n_segments[f], compactness): Lab, regular grid of centres, n_iter (the reference: 10) rounds of windowed nearest-centre
 * assignment and centre update, then - connectivity != 0, the reference's default - the scan-order breadth-first
 * enforce_connectivity pass.  labels (F, H, W) int32, DEVICE, starting at 1 (0 only where enforce_connectivity found nothing
 * to merge into, as in scikit-image).  n_segments: HOST array of F counts (--randomise-superpixels draws one per frame).
 * scikit-image is not in this image, so the definition is oracle/slic_oracle.py's restatement of its published algorithm
 * (deviations listed there: 2^-24 feature quantisation, fixed Newton cube root, empty segments stay empty): parity with
 * scikit-image itself is UNPINNED.  At most 2048 grid centres per frame.  Workspace: crw_slic_workspace_bytes (0 = bad
 * arguments). */
size_t crw_slic_workspace_bytes(int F, int H, int W, const int* n_segments, int n_iter);
int crw_slic(const float* video, int F, int H, int W, const int* n_segments, double compactness, int n_iter, int connectivity,
             int* labels, void* workspace, size_t workspace_bytes, crw_stream_t stream);

/* The connectivity pass of crw_slic on its own (scikit-image _enforce_label_connectivity_cython, called at the end of slic()):
 * segments (F, H, W) int32 DEVICE, any labelling -> labels (F, H, W) int32: 4-connected components relabelled from 1 in scan
 * order, components below min_size merged into the last labelled neighbour the breadth-first search met (0 if none), the search
 * cut at max_size pixels.  Workspace: crw_label_connectivity_workspace_bytes. */
size_t crw_label_connectivity_workspace_bytes(int F, int H, int W);
int crw_label_connectivity(const int* segments, int F, int H, int W, int min_size, int max_size, int* labels, void* workspace,
                           size_t workspace_bytes, crw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CRW_B200_H */
