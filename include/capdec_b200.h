/*
 * capdec_b200 — C ABI of the B200-native (sm_100a) CapDec training-step hot path.
 *
 * The reference (DavidHuji/CapDec) has no FFI: its hot path is Python calling torch / HuggingFace modules
 * (train.py:345-354).  This header is the seam we introduce beneath the reference's Python class surface
 * (SURVEY.md §8b): one entry point per fused op, plain device pointers + sizes + a cudaStream_t, no torch
 * types.  Each declaration cites the reference code it replaces (file:line into DavidHuji/CapDec, `HF:` =
 * transformers' modeling_gpt2.py / pytorch_utils.py / activations.py as called from train.py:259).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator); callee never allocates;
 *   - launches are asynchronous on `stream` (pass torch.cuda.current_stream()); CUDA-graph capturable;
 *   - return 0 on success, <0 on error; capdec_last_error() returns a thread-local message; nothing throws;
 *   - all tensors fp32 row-major unless noted; token ids are int64 (as torch.int64 in train.py:57);
 *   - RNG: `seed_dev` is a DEVICE pointer to a 64-bit Philox seed (NULL = 0); dropout masks / noise are functions of
 *     (*seed_dev, stream_id, element index) and are regenerated in backward, never stored.  Keeping the seed in
 *     device memory (advanced by capdec_step_clock) gives fresh masks on every CUDA-graph replay.
 */
#ifndef CAPDEC_B200_H_
#define CAPDEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* capdec_stream_t; /* cudaStream_t */

const char* capdec_last_error(void);
int capdec_version(void);
/* number of kernels launched by this library since load (the bench's `gpu_launches` evidence) */
int64_t capdec_launch_count(void);

/* ---- GEMM: C[M,N] = act( A[M,K] . B[N,K]^T + bias[N] )  on tcgen05 (TF32 in, FP32 accumulate in TMEM) -----------
 * Replaces every dense contraction on the path: nn.Linear (train.py:113-118, :124-126, :144-147, :241),
 * HF Conv1D addmm (HF:pytorch_utils.py:97-123, used by c_attn/c_proj/c_fc HF:modeling_gpt2.py:106-107,232-233),
 * lm_head (HF:modeling_gpt2.py:703-706) and all their autograd dgrad/wgrad products (train.py:351).
 *   a_major / b_major: 0 = K-major  (A stored [M,K] / B stored [N,K], K contiguous, leading dim = row pitch)
 *                      1 = MN-major (A stored [K,M] / B stored [K,N], M resp. N contiguous)
 *   precision: 0 = 1xTF32 (perf mode); 1 = 3xTF32 (fp32-grade): the fp32 operand tiles are split into
 *              hi = RN_tf32(x), lo = x - hi INSIDE the kernel's shared-memory pipeline (converter warps) and every k-step
 *              issues lo*hi + hi*lo + hi*hi into the same TMEM accumulator; operands are read from HBM once, epilogue
 *              activations use exact tanhf.  a_lo / b_lo are ignored (kept for ABI stability; pass NULL).
 *   act: 0 none, 1 gelu_new (HF:activations.py:59-66), 2 tanh (train.py:106 MLP act), 3 relu (train.py:121),
 *        4 gelu_new with aux <- gelu_new'(pre-activation) instead of the pre-activation (feeds mul_act 4 below)
 *   aux: optional second output receiving the PRE-activation (needed by backward); ld = ldc
 *   accumulate: C += result (TMA reduce-add in L2); required for split_k > 1 (wgrad over M = B*T)
 *   block_n / split_k: 0 = auto
 */
int capdec_gemm_tf32(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb, float* C,
                     int64_t ldc, int M, int N, int K, const float* bias, int act, float* aux, int accumulate,
                     int precision, const float* a_lo, const float* b_lo, int block_n, int split_k,
                     capdec_stream_t stream);

/* Same, with data-dependent extents read from DEVICE scalars at run time (shapes stay static, CUDA-graph friendly):
 * rows >= *m_limit_dev are not computed (whole 256/512-row cluster tiles are skipped), the reduction stops at
 * *k_limit_dev (rounded up to 32; the caller guarantees that rows of the K-tail hold zeros in one operand).
 * Used for the LM head over the NON-IGNORED caption tokens only (train.py:349-350). Either pointer may be NULL. */
int capdec_gemm_tf32_ex(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb, float* C,
                        int64_t ldc, int M, int N, int K, const float* bias, int act, float* aux, int accumulate,
                        int precision, const float* a_lo, const float* b_lo, int block_n, int split_k,
                        const int32_t* m_limit_dev, const int32_t* k_limit_dev, capdec_stream_t stream);

/* dgrad GEMM with a fused activation backward and bias gradient:  C[M,N] = (A . B^T) * act'(mul_in[M,N]),
 * colsum[n] += sum_m C[m,n] (may be NULL).  mul_act: 4 = mul_in already holds the derivative (forward act 4),
 * 1 = gelu_new'(pre-activation u) (HF:activations.py:59-66),
 * 2 = tanh' = 1 - a^2 with a the activated output (train.py:106), 3 = relu mask from the activated output
 * (train.py:121).  mul_in shares C's leading dimension.  capdec_gemm_tf32_mul runs 1xTF32; the _ex form takes the
 * `precision` of capdec_gemm_tf32 (1 = 3xTF32 with exact derivative arithmetic). */
int capdec_gemm_tf32_mul(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb, float* C,
                         int64_t ldc, int M, int N, int K, const float* mul_in, int mul_act, float* colsum,
                         int block_n, const int32_t* m_limit_dev, capdec_stream_t stream);
int capdec_gemm_tf32_mul_ex(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb, float* C,
                            int64_t ldc, int M, int N, int K, const float* mul_in, int mul_act, float* colsum,
                            int block_n, const int32_t* m_limit_dev, int precision, capdec_stream_t stream);

/* debug/bring-up override of the UMMA shared-memory descriptor encoding for MN-major operands
 * (layout_type, LBO bytes, SBO bytes, TMA swizzle enum); pass -1 to keep the default. Not used in production. */
void capdec_gemm_debug_mn_encoding(int layout_type, int lbo_bytes, int sbo_bytes, int tma_swizzle);
/* tile engine selection override: -1 auto (CTA pairs / cta_group::2 whenever M > 128 and N >= 128), 0 = always one CTA
 * per 128-row tile (cta_group::1), 1 = always CTA pairs.  Used by the tests to cover both engines. */
void capdec_gemm_debug_force_pair(int mode);
/* bring-up: following GEMM launches write clock64 stamps of CTA 0's roles into trace_dev (int64 [4][64], device; NULL = off) */
void capdec_gemm_debug_trace(void* trace_dev);
/* Tile order of the persistent GEMM kernels launched from now on (process-wide; returns the previous setting).
 * 0 = static: cluster c takes tiles c, c + #clusters, ... - the fastest order while a GEMM owns the whole GPU (default).
 * 1 = dynamic: one scheduler thread per cluster draws tile ids from a device-wide counter and publishes them to a ring in
 *     every CTA of the cluster; a cluster that shares its SMs with a collective or another kernel simply draws fewer tiles
 *     (used by the data-parallel Trainer while gradient buckets are all-reduced beside the backward pass).
 * Results do not depend on the order (split-K reduce-add order is not deterministic in either).  Environment override:
 * CAPDEC_GEMM_SCHED=static|dynamic. */
int capdec_gemm_set_schedule(int dynamic);
/* Tiling hint for the GEMMs this thread launches next with a device-side row limit (m_limit_dev): the live row count
 * of a packed caption batch is data dependent, so tile width / engine / wave quantisation are chosen for `rows`
 * (0 clears the hint = plan for the static M).  Never changes results.  MN-major operands whose extent is not a
 * multiple of 32 are fetched with 32-wide boxes when the row pitch covers the rounded extent: the caller guarantees
 * that every row is readable over its full pitch (true for any pitch-allocated matrix). */
void capdec_gemm_set_row_hint(int rows);
/* Measured plan selection.  enable = 1: the first call for each problem signature (shape, operand majors, epilogue,
 * limits) times a short list of (engine, tile width, split-K) plans on the caller's stream with the caller's operands
 * and remembers the fastest; later calls - and calls under CUDA-graph capture - reuse it.  A measuring call launches
 * the problem several times, so accumulate outputs / fused column sums of that call are garbage: run it on a
 * throw-away step (Trainer.autotune does).  enable = 0 stops measuring (remembered plans stay in use); -1 forgets
 * them.  Returns the number of remembered plans. */
int capdec_gemm_autotune(int enable);
/* The heuristic plan for a problem, without launching (host arithmetic only).  row_limited != 0: plan as for a GEMM with
 * a device-side row limit, i.e. for the rows given to capdec_gemm_set_row_hint.  Returns
 * engine (0 single CTA, 1 CTA pair, 2/3 quads) | tile width << 8 | split-K factor << 20, or a negative error code. */
int capdec_gemm_plan_query(int M, int N, int K, int b_major, int accumulate, int block_n, int split_k, int row_limited);

/* fp32 CUDA-core GEMM with the same contract (verification kernel: exact fp32 FMA, no tensor cores). */
int capdec_gemm_fp32_simt(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb,
                          float* C, int64_t ldc, int M, int N, int K, const float* bias, int act, float* aux,
                          int accumulate, capdec_stream_t stream);

/* hi = RN_tf32(x), lo = RN_tf32(x - hi) (both exact TF32 values).  n % 4 == 0.  Stand-alone restatement of the split the
 * 3xTF32 GEMM performs in shared memory (test / analysis helper; the GEMM no longer needs pre-split operands). */
int capdec_split_tf32(const float* x, float* hi, float* lo, int64_t n, capdec_stream_t stream);

/* ---- noise injection: train.py:27-39 (+ :18-24 uniform-ball variant) ---------------------------------------------
 * out = normalize( normalize(x) [unless dont_norm] + noise + offset ), rows of length D.
 * noise: if `noise` != NULL it is used verbatim (parity mode: caller supplies torch.randn*std); otherwise a
 * Philox Gaussian N(0, variance) (uniform_ball=0) or uniform-ball of radius sqrt(variance) (uniform_ball=1).
 * variance == 0 -> identity copy (train.py:28-29: no normalisation at all).  offset may be NULL ([D]). */
int capdec_noise_injection(const float* x, float* out, int B, int D, float variance, const float* noise,
                           const float* offset, int uniform_ball, int dont_norm, const uint64_t* seed_dev, uint64_t step,
                           capdec_stream_t stream);

/* ---- embedding assembly: train.py:253-255 + HF:modeling_gpt2.py:579-585,612 ----------------------------------------
 * h[b,t,:] = (t < P ? prefix_proj[b,t,:] : wte[tokens[b,t-P],:]) + wpe[t,:], then dropout(p) (train mode).
 * tokens int64 [B,L]; prefix_proj [B,P,d]; h [B,P+L,d].  tokens may be NULL with L = 0 (inputs_embeds path). */
int capdec_embed_fwd(const int64_t* tokens, const float* prefix_proj, const float* wte, const float* wpe, float* h,
                     int B, int P, int L, int d, int vocab, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                     capdec_stream_t stream);
/* backward: d_wte[tokens] += dh (atomic scatter; may be NULL = frozen), d_wpe[t] += sum_b dh (may be NULL),
 * d_prefix_proj[b,t<P] = dh (may be NULL). dropout mask regenerated. */
int capdec_embed_bwd(const int64_t* tokens, const float* dh, float* d_prefix_proj, float* d_wte, float* d_wpe, int B,
                     int P, int L, int d, int vocab, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                     capdec_stream_t stream);

/* ---- (residual add +) LayerNorm: nn.LayerNorm(eps=1e-5) HF:modeling_gpt2.py:252-254,505 ; train.py:184-188 -------
 * fwd: r = h_in + dropout(y) (y may be NULL -> r = h_in);  x = LN(r)*gamma + beta;  stats[row] = (mean, rstd).
 *      h_out receives r (may alias h_in; may be NULL when y == NULL).  rows x d, d % 128 == 0, d <= 1024. */
int capdec_add_ln_fwd(const float* h_in, const float* y, float* h_out, float* x, float* stats, const float* gamma,
                      const float* beta, int rows, int d, float eps, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                      const int32_t* rows_dev, capdec_stream_t stream);
/* bwd: dr = dh_res (running residual gradient, may be NULL = 0) + LN_bwd(dx; r, stats, gamma) -> written to dh_out
 *      (may alias dh_res); if dy != NULL: dy = dropout_mask * dr (gradient of the branch output y).
 *      dgamma/dbeta accumulated (+=) unless NULL (frozen GPT-2, train.py:276-284).
 *      dbias_branch (may be NULL): += sum over rows of (dropout_mask * dr) = the bias gradient of the Linear/Conv1D
 *      whose output was the branch y (attn.c_proj / mlp.c_proj / mapper project / fc2), fused here to save a pass. */
int capdec_add_ln_bwd(const float* dx, const float* r, const float* stats, const float* gamma, const float* dh_res,
                      float* dh_out, float* dy, float* dgamma, float* dbeta, float* dbias_branch, int rows, int d,
                      float p_drop,
                      const uint64_t* seed_dev, uint32_t stream_id, const int32_t* rows_dev, capdec_stream_t stream);
/* rows_dev (both LayerNorm entry points, may be NULL): device scalar with the live row count of a packed batch
 * (capdec_pack_plan); rows >= *rows_dev are not touched, except that bwd zero-fills dh_out / dy up to the next multiple
 * of 32 rows so that K-limited weight-gradient GEMMs can read whole k-blocks. */

/* ---- attention core ----------------------------------------------------------------------------------------------
 * GPT-2 (HF:modeling_gpt2.py:54-72,185-191): qkv [B,T,3*H*hd] (q|k|v thirds, heads contiguous hd slices),
 * causal softmax(q k^T * scale) (dropout p on probabilities) v -> ctx [B,T,H*hd].
 * Mapper (train.py:150-167): q [B,T,H*hd] from to_queries, kv [B,S,2*H*hd] from to_keys_values, no mask.
 * Generic strided form: element (b,t,h,:) of q at q + b*q_bs + t*q_ts + h*hd.  lse [B,H,T] saved for backward.
 * key_len: optional int32 [B] number of valid keys (padding mask, HF attention_mask); NULL = all valid. */
int capdec_attention_fwd(const float* q, const float* k, const float* v, float* ctx, float* lse, int B, int H, int T,
                         int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs,
                         int64_t o_ts, float scale, int causal, const int32_t* key_len, float p_drop,
                         const uint64_t* seed_dev, uint32_t stream_id, capdec_stream_t stream);
/* dbias_qkv (may be NULL): [3*H*hd] += column sums of (dq | dk | dv) over all rows = c_attn.bias gradient */
int capdec_attention_bwd(const float* q, const float* k, const float* v, const float* ctx, const float* dctx,
                         const float* lse, float* dq, float* dk, float* dv, float* dbias_qkv, int B, int H, int T,
                         int S, int hd,
                         int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs, int64_t o_ts,
                         float scale, int causal, const int32_t* key_len, float p_drop,
                         const uint64_t* seed_dev, uint32_t stream_id, capdec_stream_t stream);

/* Tensor-core variants (mma.sync m16n8k8, TF32 in / FP32 accumulate) with the identical contract and dropout
 * mapping; used in the 1xTF32 precision mode.  Return -3 (unsupported) if the tile exceeds shared memory. */
int capdec_attention_tc_fwd(const float* q, const float* k, const float* v, float* ctx, float* lse, int B, int H, int T,
                            int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs,
                            int64_t o_ts, float scale, int causal, const int32_t* key_len, float p_drop,
                            const uint64_t* seed_dev, uint32_t stream_id, const int32_t* cu_rows, capdec_stream_t stream);
int capdec_attention_tc_bwd(const float* q, const float* k, const float* v, const float* ctx, const float* dctx,
                            const float* lse, float* dq, float* dk, float* dv, float* dbias_qkv, int B, int H, int T,
                            int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs,
                            int64_t o_ts, float scale, int causal, const int32_t* key_len, float p_drop,
                            const uint64_t* seed_dev, uint32_t stream_id, const int32_t* cu_rows, capdec_stream_t stream);
/* cu_rows (may be NULL): packed rows — caption b owns rows [cu_rows[b], cu_rows[b+1]) of q/k/v/ctx (causal
 * self-attention, T = S = that count <= the T argument, which stays the pitch of lse and of the dropout counters). */
/* fp32-grade variants of the two above for the 3xTF32 mode: every MMA fragment is split into hi = RN_tf32(x) and
 * lo = x - hi in registers and accumulated as lo*hi + hi*lo + hi*hi; expf / logf instead of the MUFU approximations.
 * Same contract, same dropout mapping, packed rows included. */
int capdec_attention_tc_fwd_x3(const float* q, const float* k, const float* v, float* ctx, float* lse, int B, int H, int T,
                               int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs,
                               int64_t o_ts, float scale, int causal, const int32_t* key_len, float p_drop,
                               const uint64_t* seed_dev, uint32_t stream_id, const int32_t* cu_rows, capdec_stream_t stream);
int capdec_attention_tc_bwd_x3(const float* q, const float* k, const float* v, const float* ctx, const float* dctx,
                               const float* lse, float* dq, float* dk, float* dv, float* dbias_qkv, int B, int H, int T,
                               int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs,
                               int64_t o_ts, float scale, int causal, const int32_t* key_len, float p_drop,
                               const uint64_t* seed_dev, uint32_t stream_id, const int32_t* cu_rows, capdec_stream_t stream);

/* ---- masked cross entropy: train.py:349-350 (nnf.cross_entropy(..., ignore_index=0), mean over targets != 0) -----
 * logits [rows, ld] (ld >= V, padded pitch), targets int64 [rows].  loss_sum/n_valid are device scalars (float);
 * capdec_ce_count writes the number of non-ignored targets to *n_valid and zeroes *loss_sum_to_zero (may be NULL).  fwd_bwd overwrites logits with
 * dlogits = (softmax - onehot) * grad_scale / *n_valid (zero rows where target == ignore_index; n_valid == NULL
 * means 1, i.e. sum-reduced gradients for exact data-parallel averaging) and adds the row losses into loss_sum.
 * If write_grad == 0 logits are left intact (validation, train.py:383-386). */
int capdec_ce_count(const int64_t* targets, int64_t n, int64_t ignore_index, float* n_valid, float* loss_sum_to_zero,
                    capdec_stream_t stream);
int capdec_ce_fwd_bwd(float* logits, int64_t ld, const int64_t* targets, int rows, int V, int64_t ignore_index,
                      const float* n_valid, float grad_scale, float* loss_sum, int write_grad,
                      const int32_t* row_limit_dev /* NULL or device scalar: rows >= it are skipped */,
                      capdec_stream_t stream);
/* Compaction of the consumed logits rows: targets int64 [B,L] -> row_src[r] (hidden row b*T+off+j feeding compact row
 * r), dst_of[b*T+t] (compact row or -1), targets_c[r], counts = {n_valid, n_valid rounded up to 32}; also writes
 * *n_valid (float) and zeroes *loss_sum_to_zero.  Replaces capdec_ce_count on the compacted path. */
int capdec_compact_targets(const int64_t* targets, int B, int L, int T, int off, int64_t ignore_index,
                           int32_t* row_src, int32_t* dst_of, int64_t* targets_c, int32_t* counts, float* n_valid,
                           float* loss_sum_to_zero, const int32_t* cu_rows, capdec_stream_t stream);
/* cu_rows (may be NULL): hidden rows are packed, caption b starts at row cu_rows[b] instead of b*T.
 * dst[r] = src[row_src[r]] for r < counts[0], zero rows up to counts[1]; dst has max_rows rows of width d */
int capdec_rows_gather_idx(const float* src, float* dst, const int32_t* row_src, const int32_t* counts, int max_rows,
                           int d, capdec_stream_t stream);
/* dst[row] = dst_of[row] >= 0 ? src[dst_of[row]] : 0 for every one of `rows` rows */
int capdec_rows_scatter_idx(const float* src, float* dst, const int32_t* dst_of, int rows, int d, capdec_stream_t stream);

/* ---- packed-row execution: run the trunk only over the positions that can reach the loss ------------------------------
 * Caption b (tokens [B,L], right-padded with 0, train.py:55-63) keeps positions 0 .. P + len_b - 2, len_b = 1 + index of
 * its last non-zero token: later rows only feed ignored targets (train.py:349-350) and no live row attends to them.
 * pack_plan: cu[B+1] (first packed row per caption), rows = {live rows, live rounded up to 32}, row_bt[r] = b<<8 | t. */
int capdec_pack_plan(const int64_t* tokens, int B, int L, int P, int32_t* cu, int32_t* rows, int32_t* row_bt,
                     capdec_stream_t stream);
/* packed forms of capdec_embed_fwd / capdec_embed_bwd (train.py:253-255): h / dh have one row per live position */
int capdec_embed_fwd_packed(const int64_t* tokens, const float* prefix_proj, const float* wte, const float* wpe, float* h,
                            const int32_t* row_bt, const int32_t* rows, int B, int P, int L, int d, int vocab, float p_drop,
                            const uint64_t* seed_dev, uint32_t stream_id, capdec_stream_t stream);
int capdec_embed_bwd_packed(const int64_t* tokens, const float* dh, float* d_prefix_proj, float* d_wte, float* d_wpe,
                            const int32_t* cu, int B, int P, int L, int d, int vocab, float p_drop, const uint64_t* seed_dev,
                            uint32_t stream_id, capdec_stream_t stream);
/* rows [rows[0], rows[1]) of buf [*, ld] <- 0 */
int capdec_zero_tail_rows(float* buf, int64_t ld, const int32_t* rows, capdec_stream_t stream);

/* ---- KV-cached batched beam search (replaces gpt2_prefix_eval.py:50-115 generate_beam) -----------------
 * R = n_img*beam physical rows; caches are per layer [R][Tmax][d]; `src` is the int32 [2][R][Tmax] lineage table
 * (position t of logical beam b lives in physical row src[c&1][b][t]); `step` is the device-resident count c of
 * selections done, which every kernel below reads so that one decode step is a replayable CUDA graph.
 * beam_init: c=0, scores=0, seq_len=1 (:59), stopped=0 (:60), prefix lineage -> row img*beam. */
int capdec_beam_init(int32_t* step, float* scores, float* seq_len, int32_t* stopped, int32_t* src, int32_t* img_done,
                     int32_t* ticket, int n_img, int beam, int P, int Tmax, capdec_stream_t stream);
/* K/V thirds of the prefill's fused QKV rows [n_img*P, 3d] -> cache rows img*beam, positions 0..P-1 */
int capdec_kv_prefill(const float* qkv, float* kcache, float* vcache, int n_img, int beam, int P, int Tmax, int d,
                      capdec_stream_t stream);
/* x[b,:] = wte[hist_tok[c-1][b]] + wpe[P+c-1]   (:105 `model.gpt.transformer.wte(next_tokens)` + HF position embedding) */
int capdec_decode_embed(const int32_t* step, const int32_t* hist_tok, const float* wte, const float* wpe, float* x, int rows,
                        int P, int d, capdec_stream_t stream);
/* one query per row over its lineage in the cache; appends this token's K/V at (row, P+c-1).  head_dim 64. */
int capdec_decode_attention(const float* qkv, float* kcache, float* vcache, const int32_t* src, const int32_t* step,
                            float* ctx, int rows, int H, int head_dim, int P, int Tmax, float scale, capdec_stream_t stream);
/* per logits row: lse of logits/temperature (:77-79) and its k largest entries (value, index), 8 slots per row */
int capdec_row_topk(const float* logits, int64_t ld, int rows, int V, float temperature, int k, float* cand_val,
                    int32_t* cand_idx, float* row_lse, capdec_stream_t stream);
/* per image: top-`beam` of (scores + logp)/seq_len over beam*V (:80-100), stop bookkeeping (:106), lineage update,
 * history (hist_tok/hist_parent [max_sel][R]), img_done[img] = all beams stopped (:107), then c += 1. */
int capdec_beam_select(const float* cand_val, const int32_t* cand_idx, const float* row_lse, int32_t* step, float* scores,
                       float* seq_len, int32_t* stopped, int32_t* src, int32_t* hist_tok, int32_t* hist_parent,
                       int32_t* img_done, int32_t* ticket, int n_img, int beam, int P, int Tmax, int V, int stop_token,
                       capdec_stream_t stream);

/* ---- small fused elementwise / reduction ops ----------------------------------------------------------------------
 * colsum: out[n] += sum_m x[m,n]  (bias gradients of every Linear/Conv1D; autograd of addmm bias) */
int capdec_colsum_acc(const float* x, int64_t ld, float* out, int M, int N, capdec_stream_t stream);
/* dx[M,N] = dy * act'(.) ; act: 1 gelu_new (pre = pre-activation), 2 tanh (pre = activated output a: 1-a^2),
 * 3 relu (pre = activated output).  dbias (may be NULL): [N] += column sums of dx (bias gradient of the Linear that
 * produced the pre-activation), fused to save a pass.  dx may alias dy.  Contiguous [M,N], N % 4 == 0. */
int capdec_act_bwd(const float* dy, const float* pre, float* dx, float* dbias, int M, int N, int act,
                   capdec_stream_t stream);
/* row gather / scatter of d-wide rows: dst[i,:] = src[map(i),:] with map(i) = (i / L)*T + off + i % L
 * (logits slice [:, P-1:-1] of train.py:349, expressed on the hidden states) */
int capdec_rows_gather(const float* src, float* dst, int B, int T, int L, int off, int d, capdec_stream_t stream);
int capdec_rows_scatter(const float* src, float* dst, int B, int T, int L, int off, int d, capdec_stream_t stream);
/* TransformerMapper input assembly (train.py:230-233): x[b, s<C] = lin[b, s], x[b, C+s] = prefix_const[s] */
int capdec_mapper_concat_fwd(const float* lin, const float* prefix_const, float* x, int B, int C, int P, int d,
                             capdec_stream_t stream);
int capdec_mapper_concat_bwd(const float* dx, float* dlin, float* dprefix_const, int B, int C, int P, int d,
                             capdec_stream_t stream);

/* ---- AdamW, HuggingFace-4.24 semantics (train.py:326,352) ----------------------------------------
 * g' = g / *grad_denom_dev (NULL = 1: e.g. the all-reduced count of non-ignored targets) ;
 * m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2 ; p -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps) ;
 * p -= lr*wd*p.  lr and the step count t are read from device scalars so that the step is CUDA-graph
 * replayable.  zero_grad != 0 clears g in the same pass (train.py:354). */
int capdec_adamw_step(float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev, const float* t_dev,
                      float beta1, float beta2, float eps, float weight_decay, const float* grad_denom_dev,
                      int zero_grad, capdec_stream_t stream);

/* ---- data-parallel optimizer step over NVLink peer memory (train.py:352-354 under N ranks) --------------
 * One process per GPU.  capdec_peer_export: CUDA-IPC handle (64 bytes) of the device allocation that contains `ptr` and
 * the byte offset of `ptr` inside it; capdec_peer_open (in ANOTHER process of the same node): maps that allocation with
 * peer access and returns the address that corresponds to `ptr`; capdec_peer_close unmaps it.
 * capdec_adamw_peer_step: g_slices / p_peers are HOST arrays of `world` device pointers.  This rank owns elements
 * [lo, lo + n) of the flat parameter buffer.  g_slices[r] = where the gradients of that slice, as computed by rank r, can
 * be read ([n] floats: a peer's gradient buffer + lo, or a local staging area peers pushed into with capdec_copy_async);
 * p_peers[r] = the parameter buffer of rank r (element 0 = the first trainable parameter; entry `rank` is this process's
 * own).  g = sum over r of g_slices[r][i] (rank order), the HF-AdamW update of capdec_adamw_step with this rank's moments
 * m, v ([n] each), and the new value stored to p_peers[r][lo + i] for EVERY r - reduce-scatter + update + all-gather in
 * one kernel.  The caller orders it across ranks: every rank's gradients are final (and pushed) before any rank launches,
 * and no rank reads its parameters / clears its gradients before every rank's launch has finished
 * (capdec_b200/trainer.py brackets it with two 16-byte all-reduces).
 * capdec_copy_async: cudaMemcpyAsync(dst, src, bytes, default kind) on `stream` - device / peer pointers; a memcpy node
 * under CUDA-graph capture, executed by the copy engines. */
int capdec_peer_export(const void* ptr, void* handle64, int64_t* offset_out);
int capdec_peer_open(const void* handle64, int64_t offset, void** ptr_out);
int capdec_peer_close(void* ptr, int64_t offset);
int capdec_copy_async(void* dst, const void* src, int64_t bytes, capdec_stream_t stream);
int capdec_adamw_peer_step(void* const* g_slices, void* const* p_peers, int world, int rank, int64_t lo, int64_t n,
                           float* m, float* v, const float* lr_dev, const float* t_dev, float beta1, float beta2,
                           float eps, float weight_decay, const float* grad_denom_dev, capdec_stream_t stream);

/* ---- device-side training clock (train.py:328-330,352-354) --------------------------------------------------------
 * *seed_dev += golden ratio (next step's RNG stream); if step_dev != NULL: n = *step_dev (updates done so far),
 * *lr_dev = base_lr * linear-warm-up/decay factor(n) (HF get_linear_schedule_with_warmup), *t_dev = n + 1 (Adam bias
 * correction step), *step_dev = n + 1.  One thread; exists so that the whole step is CUDA-graph replayable. */
int capdec_step_clock(uint64_t* seed_dev, float* step_dev, float* lr_dev, float* t_dev, float base_lr,
                      int warmup_steps, int total_steps, capdec_stream_t stream);

/* ---- device-resident data feed (ClipCocoDataset.pad_tokens / __getitem__, train.py:52-72; collate + H2D :327,:346) --
 * tokens_all int32 [N, L]: captions pre-padded / truncated to L = max_seq_len with -1 in the padding; cap2emb int32 [N];
 * table [E, D] fp32 (table_fp16 = 0) or fp16 (= 1); idx int64 [B] caption indices of this batch.
 * tokens int64 [B, L] = max(t, 0); mask fp32 [B, P+L] = [1]*P ++ (t >= 0) (may be NULL: the fast path never needs it);
 * prefix fp32 [B, D] = table[cap2emb[i]] (/ its L2 norm when normalize != 0; no epsilon, as train.py:71). */
int capdec_batch_gather(const int32_t* tokens_all, const int32_t* cap2emb, const void* table, int table_fp16,
                        const int64_t* idx, int64_t* tokens, float* mask, float* prefix, int B, int L, int P, int D,
                        int normalize, capdec_stream_t stream);

/* zero-fill `bytes` bytes at p on `stream` (cudaMemsetAsync; a memset node under graph capture) */
int capdec_zero_fill(void* p, int64_t bytes, capdec_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CAPDEC_B200_H_ */
