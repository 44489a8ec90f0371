/*
 * ovqa.h — C ABI of libovqa_b200.so: the B200 (sm_100a) kernels behind OpenViVQA's cross-modal
 * attention hot path.
 *
 * The reference (hieunghia-pat/OpenViVQA) is pure Python/PyTorch and has no FFI of its own; the boundary a
 * maintainer binds is therefore "one entry point per op site of the hot path" (SURVEY.md §2.1 table /
 * §8a).  Each entry point cites the reference lines it replaces.  Conventions:
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless stated otherwise;
 *   - the CALLER owns every buffer (inputs, outputs, saved statistics, workspaces);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); entry points never synchronise;
 *   - return value: 0 on success, a negative OVQA_ERR_* otherwise (never throws, never falls back to CPU);
 *   - dtype arguments are OVQA_F32 / OVQA_BF16 and name the STORAGE type; arithmetic is always fp32
 *     accumulate (tcgen05 bf16xbf16->fp32 for the tensor-core GEMMs).
 */
#ifndef OVQA_H_
#define OVQA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVQA_F32  0
#define OVQA_BF16 1

#define OVQA_OK                 0
#define OVQA_ERR_BAD_ARG       -1   /* null pointer, negative size, unsupported dtype            */
#define OVQA_ERR_BAD_SHAPE     -2   /* shape outside the kernel's supported range                 */
#define OVQA_ERR_ALIGNMENT     -3   /* pointer / leading dimension not aligned as required        */
#define OVQA_ERR_CUDA          -4   /* a CUDA runtime/driver call failed (see ovqa_last_cuda_error) */
#define OVQA_ERR_NO_DEVICE     -5   /* no sm_100 device / driver entry point unavailable          */

/* epilogue flags for ovqa_gemm */
#define OVQA_EPI_GELU        1      /* exact erf GELU (positionwise_feed_forward.py:24)           */
#define OVQA_EPI_DGELU       2      /* multiply by GELU'(aux) (backward of the above)             */
#define OVQA_EPI_SAVE_DGELU  4      /* with GELU + aux_out: aux_out receives GELU'(v) x dropout scale */
#define OVQA_EPI_MUL_AUX     8      /* multiply by aux (the derivative saved by the forward pass) */
#define OVQA_EPI_RELU       16      /* max(v, 0) (the attribute-reduce MLP, models/mcan.py:21-25)    */

#define OVQA_OP_N 0                 /* operand stored with the reduction (K) dimension contiguous */
#define OVQA_OP_T 1                 /* operand stored with its M/N dimension contiguous           */

int         ovqa_version(void);
const char* ovqa_error_string(int code);
int         ovqa_last_cuda_error(void);           /* cudaError_t of the last failing call (thread local) */
int         ovqa_device_sm_count(void);          /* SMs the persistent kernels size their grids for (device count minus the reserve) */
int         ovqa_set_sm_reserve(int sms);         /* leave `sms` SMs (rounded up to even) to co-running kernels of other streams (NCCL during a
                                                     data-parallel backward pass); returns the previous reserve.  Initial value: $OVQA_SM_RESERVE or 0 */
uint64_t    ovqa_launch_count(void);              /* kernels launched by this library so far (process wide) */

/* ---------------------------------------------------------------------------------------------------
 * Dense contraction with fused epilogue.            C[m,n] = epi( sum_k A[m,k] * B[n,k] )
 *   replaces: nn.Linear at attentions.py:49-51,58 (fc_q/fc_k/fc_v/fc_o), positionwise_feed_forward.py:24-25
 *             (fc1+GELU, fc2), mmf_m4c.py:387-392 (OcrPtrNet query/key) and their autograd backward
 *             (dgrad / wgrad).
 *   A: a_op==OVQA_OP_N -> stored [m, k] (row stride lda);  OVQA_OP_T -> stored [k, m] (row stride lda)
 *   B: b_op==OVQA_OP_N -> stored [n, k] (row stride ldb);  OVQA_OP_T -> stored [k, n] (row stride ldb)
 *      (y = x W^T : N,N.   dx = dy W : A N, B T.   dW = dy^T x : A T, B T.)
 *   ab_dtype: OVQA_BF16 runs on tcgen05/TMEM fed by TMA; OVQA_F32 runs the exact fp32 FFMA path
 *             (the 1e-4 parity path).  c_dtype: storage of C (and of c2).
 *   epilogue, in order:  v = acc*alpha + bias[n];  if aux_out: aux_out = v (pre-activation, c_dtype) — or, with
 *             OVQA_EPI_SAVE_DGELU, d(dropout(GELU(v)))/dv = GELU'(v) * s, s = this element's dropout scale (0 or 1/(1-p); 1 if
 *             p == 0), which lets the backward GEMM use OVQA_EPI_MUL_AUX alone (no transcendental, no RNG);
 *             GELU / DGELU(aux_in) / MUL_AUX(aux_in);  dropout(p, seed, offset [+ *offset_base]) if p>0;  v += residual[m,n] (fp32);
 *             if beta!=0: v += beta*C_old (fp32 C only);  C = v;  if c2: c2 = bf16(v).
 *   offset_base: optional DEVICE pointer to a u64 added to `offset` inside the kernel.  A host that replays a captured
 *             CUDA graph advances that word on the device between replays (ovqa_advance_u64), so every replay draws
 *             fresh dropout masks although the launch parameters are frozen in the graph.
 *   colsum_out: optional fp32 [n], += the column sums of the final C (atomically; the caller zero-fills).  When C is the
 *             gradient of a Linear's output this IS that Linear's bias gradient: the backward pass needs no separate
 *             reduction over the B*L rows (autograd of nn.Linear's bias at attentions.py:49-51, positionwise_feed_forward.py:24).
 *   split_k>1 (wgrad): partial sums are atomically added into fp32 C, which the caller zero-fills (or
 *             pre-loads); bias/epilogue flags are not allowed with split_k != 1.  split_k == 0: the library
 *             picks the split (and tile width) that fills the SMs; C must be zero-filled as for split_k>1.
 * ------------------------------------------------------------------------------------------------- */
int ovqa_gemm(const void* A, const void* B, void* C,
              int64_t m, int64_t n, int64_t k,
              int64_t lda, int64_t ldb, int64_t ldc,
              int a_op, int b_op, int ab_dtype, int c_dtype,
              float alpha, float beta,
              const float* bias, int epi_flags,
              const void* aux_in, void* aux_out, int64_t ld_aux,
              const float* residual, int64_t ld_res,
              void* c2_bf16, int64_t ldc2,
              float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* offset_base,
              float* colsum_out,
              int split_k, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Attention core.   out = softmax(q k^T * scale + mask [+ geometry log-bias]) v
 *   replaces: ScaledDotProductAttention.forward attentions.py:53-57,
 *             AugmentedGeometryScaledDotProductAttention.forward attentions.py:113-134 (+ models/utils.py
 *             :102-162 box_relational_embedding, recomputed on the fly from boxes),
 *             HF eager_attention_forward as called from mmf_m4c.py:349-351.
 *   q [B, Lq, *] rows of stride ldq, head hd at column hd*dk;  k [B, Lk, *] (ldk);  v [B, Lk, *] (ldv) head at
 *   hd*dv;  out [B, Lq, h*dv] (ldo).  Batch strides are Lq*ldq etc. (dense batches).
 *   mask: additive fp32 (value -1e5 in the reference, models/utils.py:56) with element strides
 *         (mask_sb, mask_sh, mask_sq) and unit stride over keys; any stride may be 0 (broadcast); NULL = none.
 *         The add happens in fp32 BEFORE the row max is taken.
 *   geometry (all NULL/0 for the plain core): boxes [B, N, 4] fp32 (x_min,y_min,x_max,y_max), requires
 *         Lq==Lk==N; geo_w [h, d_g], geo_b [h]; d_g = 64-style trig embedding (geo_trig=1, d_g%8==0) or 4;
 *         geo_freq [d_g/8] fp32 optional: the 1/1000^(k/(d_g/8)) table (models/utils.py:146-148) as the host
 *         computed it (NULL -> computed in-kernel with powf).
 *   lse  [B, h, Lq] fp32, optional: row log-sum-exp, saved for backward / for ovqa_sdpa_probs.
 *   att  [B, h, Lq, Lk] fp32, optional: the probabilities (direct callers of the core consume them:
 *         mmf_lorra.py:118-146).
 *   prob_dropout_*: dropout applied to the PROBABILITIES before they multiply v (HF BertSelfAttention's attention_probs_dropout_prob, the
 *         layers MMF_M4C trains through: mmf_m4c.py:274-276,349-351); p = 0 for the repo's own cores.  Keep decision of element
 *         ((b*h + head)*Lq + i)*Lk + j from the library's counter stream (seed, offset [+ *offset_base]); ovqa_sdpa_bwd regenerates it from
 *         the same triple.  p > 0 runs on the CUDA-core kernels.
 * ------------------------------------------------------------------------------------------------- */
int ovqa_sdpa_fwd(const void* q, const void* k, const void* v, void* out,
                  int64_t B, int64_t h, int64_t Lq, int64_t Lk, int64_t dk, int64_t dv,
                  int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                  const float* mask, int64_t mask_sb, int64_t mask_sh, int64_t mask_sq,
                  const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq,
                  int64_t d_g, int geo_trig,
                  float scale, int dtype, float* lse, float* att,
                  float prob_dropout_p, uint64_t prob_dropout_seed, uint64_t prob_dropout_offset, const uint64_t* prob_dropout_offset_base,
                  void* stream);

/* Same core under M4C's prefix-LM mask (MMT.forward, mmf_m4c.py:310-343: (B,1,L,L) = per-key padding row repeated over the queries, with
 * the last T x T block — the decoder steps — replaced by a causal pattern) WITHOUT materialising / reading the dense mask:
 *   key_mask [B, Lk] additive fp32 row (batch stride key_mask_sb), causal_tail = T, causal_value = the additive value above the diagonal of
 *   the tail block (queries Lq-T.., keys Lk-T..).  bf16, head dim 96 or 128, Lk <= 256 (the tensor-memory kernel); returns
 *   OVQA_ERR_BAD_SHAPE for anything else: the caller then passes the dense mask to ovqa_sdpa_fwd. */
int ovqa_sdpa_fwd_prefix_lm(const void* q, const void* k, const void* v, void* out,
                            int64_t B, int64_t h, int64_t Lq, int64_t Lk, int64_t dk, int64_t dv,
                            int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                            const float* key_mask, int64_t key_mask_sb, int64_t causal_tail, float causal_value,
                            float scale, int dtype, float* lse, void* stream);

/* Backward of the core (recomputes the probabilities from q,k,mask,lse):
 *   dq, dk, dv have the layouts of q, k, v (lddq/lddk/lddv).  `delta` [B,h,Lq] fp32 is caller-provided
 *   scratch (rowsum(dout*out)).  Geometry bias participates in the recomputation; its weight gradient
 *   (geo_dw [h,d_g], geo_db [h], fp32, ACCUMULATED into) is produced when non-NULL.
 *   `att` / `datt` [B,h,Lq,Lk] fp32 (both or neither; NULL from MultiHeadAttention): the probabilities ovqa_sdpa_fwd returned and the
 *   gradient arriving THROUGH them — a direct caller of the core may consume `att` itself (mmf_lorra.py:118-146 builds its
 *   features from the attention weights), so dP = dout v^T + datt.
 *   `dbias` [B,h,Lq,Lk] fp32 or NULL: receives dS = P (dP - delta), the gradient of an additive per-(b,h,q,k) term that was passed as
 *   `mask` — how the box-geometry bias trains its weights (ovqa_geo_bias_fwd / _bwd).  bf16 only (OVQA_ERR_BAD_SHAPE otherwise). */
int ovqa_sdpa_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                  void* dq, void* dk, void* dv,
                  int64_t B, int64_t h, int64_t Lq, int64_t Lk, int64_t dk_, int64_t dv_,
                  int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo,
                  int64_t lddq, int64_t lddk, int64_t lddv,
                  const float* mask, int64_t mask_sb, int64_t mask_sh, int64_t mask_sq,
                  const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq,
                  int64_t d_g, int geo_trig,
                  float* geo_dw, float* geo_db,
                  float scale, int dtype, const float* lse, float* delta,
                  const float* att, const float* datt, float* dbias,
                  float prob_dropout_p, uint64_t prob_dropout_seed, uint64_t prob_dropout_offset, const uint64_t* prob_dropout_offset_base,
                  void* stream);

/* ---------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension with optional fused residual add.
 *   replaces: `self.layer_norm(queries + out)` attentions.py:331, positionwise_feed_forward.py:26,
 *             encoders.py:113 (+ pos), HF BertSelfOutput/BertOutput LayerNorm (eps 1e-12).
 *   y = LN(x [+ residual]) * gamma + beta [+ post_add];  x/residual/post_add fp32 [rows, D];
 *   y_f32 and/or y_bf16 (either may be NULL);  mean/rstd [rows] fp32 optional (saved for backward);
 *   sum_out fp32 optional: x+residual (the LN input, saved for backward when residual is fused).
 * ------------------------------------------------------------------------------------------------- */
int ovqa_layernorm_fwd(const float* x, const float* residual, const float* gamma, const float* beta,
                       const float* post_add, int64_t post_add_rows,
                       float* y_f32, void* y_bf16, float* sum_out, float* mean, float* rstd,
                       int64_t rows, int64_t D, float eps, void* stream);

/* dx = LN backward(dy; x, mean, rstd, gamma) [+ dx_add];  dgamma/dbeta are ACCUMULATED into (fp32 [D]).
 * dx_bf16 optional second output (feeds the following bf16 dgrad/wgrad GEMMs), optionally with the
 * forward dropout mask re-applied (dropout_p, seed, offset over the [rows, D] index space).
 * dbias optional (fp32 [D], ACCUMULATED into): column sums of the masked dx, i.e. the bias gradient of the
 * GEMM whose output fed this LayerNorm (fc_o / fc2) — saves a separate pass over dx. */
int ovqa_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                       const float* gamma, const float* dx_add,
                       float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dbias,
                       int64_t rows, int64_t D,
                       float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* offset_base, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Element-wise helpers of the path.
 * ------------------------------------------------------------------------------------------------- */
/* y(bf16) = x(fp32), n elements — staging of GEMM operands (weights once per optimiser step). */
/* FeatureEmbedding front-end (vision_embeddings.py:19-25, models/utils.py:44-57): one pass over fp32 features [rows, K]
 * (row pitch ldx) writes the contiguous bf16 GEMM operand y [rows, K] and the padding mask[rows] (mask_value where the row sums to 0, else 0;
 * NULL to skip).  K % 8 == 0. */
int ovqa_cast_rows_padmask(const float* x, void* y_bf16, float* mask, int64_t rows, int64_t K, int64_t ldx, float mask_value, void* stream);
/* The same padding mask for features that ARRIVE as bf16 (an input feed staged as bf16 in pinned host memory halves the host-to-device
 * bytes; the first GEMM rounds its operand to bf16 anyway, so results are bit-identical): mask[row] = mask_value where the row sums to 0. */
int ovqa_rows_padmask_bf16(const void* x_bf16, float* mask, int64_t rows, int64_t K, int64_t ldx, float mask_value, void* stream);
/* y_bf16[i] = a[i] * b[i]   (a fp32, b fp32 or bf16): gradient of a fused activation from its saved derivative */
int ovqa_mul_to_bf16(const float* a, const void* b, int b_dtype, void* y_bf16, int64_t n, void* stream);
int ovqa_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
int ovqa_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream);
/* out[n] (+)= sum_m x[m, n]  — bias gradients.  x dtype f32/bf16, out fp32, accumulate!=0 adds. */
int ovqa_colsum(const void* x, float* out, int64_t m, int64_t n, int64_t ldx, int dtype, int accumulate,
                void* stream);
/* y = dropout(x): y[i] = x[i] * keep(i)/(1-p) with the same counter-based mask as the GEMM/LN epilogues
 * (element index i, seed, offset); x fp32, y fp32 or bf16 (out_dtype).  p == 0 is a plain cast/copy. */
int ovqa_dropout_apply(const float* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset,
                       const uint64_t* offset_base, int out_dtype, void* stream);
/* *word += delta on the device (one thread): advances the dropout offset base between CUDA-graph replays. */
int ovqa_advance_u64(uint64_t* word, uint64_t delta, void* stream);
/* Many contiguous fp32 -> {fp32, bf16} copies in ONE launch: the once-per-optimiser-step staging of the GEMM operands
 * (bf16 weight copies, row-concatenated [Wq;Wk;Wv] / [bq;bk;bv]).  `descs` is a DEVICE array of n_desc descriptors. */
typedef struct {
    const float* src;       /* device, fp32, n elements                        */
    void*        dst;       /* device, dst_dtype, n elements (may be an offset into a larger buffer) */
    int64_t      n;
    int32_t      dst_dtype; /* OVQA_F32 | OVQA_BF16 */
    int32_t      reserved;
} ovqa_stage_desc;
int ovqa_stage_many(const ovqa_stage_desc* descs, int64_t n_desc, void* stream);
/* y = a + b (fp32, n elements) */
int ovqa_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Box-geometry bias, materialised (only for callers that want it; the fused core recomputes it):
 *   out[b,h,i,j] = log(max(relu(geo_w[h] . emb(box_i, box_j) + geo_b[h]), 1e-6))
 *   replaces models/utils.py:102-162 + attentions.py:113-119,131-132.
 * ------------------------------------------------------------------------------------------------- */
int ovqa_box_geometry_bias(const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq,
                           float* out, int64_t B, int64_t N, int64_t h, int64_t d_g, int geo_trig, void* stream);

/* Box-geometry attention bias as a tensor (the tensor-core route of AugmentedGeometryScaledDotProductAttention, attentions.py:113-134):
 *   out[b,hd,i,j] = log(clamp(relu(geo_w[hd] . emb(box_i, box_j) + geo_b[hd]), 1e-6)) [+ key_mask[b*key_mask_sb + j]]      fp32 [B,h,N,N]
 *   with emb = box_relational_embedding (models/utils.py:102-162), evaluated once per pair for all heads.  The result is passed to
 *   ovqa_sdpa_fwd / ovqa_sdpa_bwd as `mask` with strides (h*N*N, N*N, N).  ovqa_geo_bias_bwd turns the attention backward's `dbias`
 *   into the weight gradients: geo_dw [h,d_g], geo_db [h] (fp32, ACCUMULATED into).  h <= 16, d_g <= 64. */
int ovqa_geo_bias_fwd(const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq,
                      const float* key_mask, int64_t key_mask_sb, float* out,
                      int64_t B, int64_t N, int64_t h, int64_t d_g, int geo_trig, void* stream);
int ovqa_geo_bias_bwd(const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq,
                      const float* dbias, float* geo_dw, float* geo_db,
                      int64_t B, int64_t N, int64_t h, int64_t d_g, int geo_trig, void* stream);

/* Attribute-reduce pooling (models/mcan.py:70-77; the same lines in cross_modality_transformer.py / visiolinguistic_transformer.py;
 * SURVEY §8f rank 3).  hidden [B,N,Dh] = dropout(relu(fc1 x)) (ovqa_gemm with OVQA_EPI_RELU + dropout), x [B,N,D] fp32:
 *   logits[b,n] = hidden[b,n,:] . w2 + *b2;  probs = softmax over n (no padding mask, as upstream);  pooled[b,:] = sum_n probs[b,n] x[b,n,:].
 *   probs [B,N] fp32 optional (saved for backward).
 * Backward: dx [B,N,D] fp32 (optional) = probs dpooled^T;  dpre [B,N,Dh] (hidden's dtype) = the gradient of fc1's PRE-activation,
 *   dlogit[n] w2[c] / (1 - dropout_p) where hidden[b,n,c] > 0 (the unit was active and kept), else 0;  dw2 [Dh], db2 [1] fp32 ACCUMULATED into. */
int ovqa_attr_pool_fwd(const void* hidden, const float* w2, const float* b2, const float* x, float* pooled, float* probs,
                       int64_t B, int64_t N, int64_t Dh, int64_t D, int hidden_dtype, void* stream);
int ovqa_attr_pool_bwd(const float* dpooled, const void* hidden, const float* w2, const float* x, const float* probs,
                       float* dx, void* dpre, float* dw2, float* db2,
                       int64_t B, int64_t N, int64_t Dh, int64_t D, float dropout_p, int hidden_dtype, void* stream);

/* Output projection + bias + dropout + residual + LayerNorm in ONE kernel for d_model = 512 (attentions.py:58,331 fc_o -> layer_norm;
 * positionwise_feed_forward.py:25-26 fc2 -> layer_norm):  s = dropout(A W^T + bias) + residual;  y = LayerNorm(s) * gamma + beta.
 * A [M,K] bf16 (row pitch lda), W [N,K] bf16 (ldw), N must be 512 (a 128-row tile keeps its full fp32 rows in tensor memory; other widths
 * return OVQA_ERR_BAD_SHAPE and the caller runs ovqa_gemm + ovqa_layernorm_fwd); residual, s_out, y [M,N] fp32 contiguous; y_bf16 [M,N].
 * s_out (LayerNorm input, saved for the backward pass), y_bf16, mean / rstd [M] may be NULL.  Dropout: the library's counter stream
 * (element row*N + col, as ovqa_gemm), so ovqa_layernorm_bwd regenerates the same mask. */
int ovqa_gemm_ln(const void* A, const void* W, const float* bias, const float* residual, const float* gamma, const float* beta,
                 float* s_out, float* y, void* y_bf16, float* mean, float* rstd, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldw,
                 float eps, float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* offset_base, void* stream);

/* Recurrent part of the single-layer, batch-first LSTM of LSTMTextEmbedding (models/modules/text_embeddings.py:234,244: nn.LSTM(d_model, d_model,
 * batch_first=True) called without an initial state).  Gate order i, f, g, o (PyTorch).  bf16 operands, fp32 accumulate / state.
 *   xg [B,L,4H] fp32 = x W_ih^T + b_ih for all steps (one ovqa_gemm);  w_hh_bf16 [4H,H];  b_hh [4H] fp32 or NULL.
 *   y [B,L,H] fp32 = h_t.  Optional outputs (all or none are needed by the backward pass): y_bf16 [B,L,H], hprev_bf16 [B,L,H] = h_{t-1} rows
 *   (zero at t = 0: the operand of the W_hh weight gradient), gates_bf16 [B,L,4H] = activated i,f,g,o, c_out [B,L,H] fp32 = c_t.
 * One kernel per direction: a thread-block cluster of H/32 CTAs keeps W_hh in registers for all L steps and exchanges h_t (forward) /
 * the partial products of dh_{t-1} (backward) through distributed shared memory.  H must be 128, 256 or 512 (ovqa_lstm_supported);
 * other sizes return OVQA_ERR_BAD_SHAPE (the caller keeps the library LSTM).
 * Backward: dy [B,L,H] fp32 -> dgates_bf16 [B,L,4H] = gradient of the gate PRE-activations; the caller forms dx = dgates W_ih,
 * dW_ih = dgates^T x, dW_hh = dgates^T hprev, db_ih = db_hh = column sums of dgates with ovqa_gemm / ovqa_colsum. */
int ovqa_lstm_supported(int64_t H);
int ovqa_lstm_fwd(const float* xg, const void* w_hh_bf16, const float* b_hh, float* y, void* y_bf16, void* hprev_bf16, void* gates_bf16,
                  float* c_out, int64_t B, int64_t L, int64_t H, void* stream);
int ovqa_lstm_bwd(const float* dy, const void* w_hh_bf16, const void* gates_bf16, const float* c_all, void* dgates_bf16,
                  int64_t B, int64_t L, int64_t H, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * M4C dynamic pointer network and greedy decode step.
 * ------------------------------------------------------------------------------------------------- */
/* scores[b,t,j] = (q[b,t,:] . k[b,j,:]) * scale + mask[b,j]   written into a row of width ld_scores at
 * column col_offset (so the caller's (B,T,V+Locr) score tensor is filled in place, no torch.cat).
 *   replaces OcrPtrNet.forward mmf_m4c.py:394-396 (+ the cat at :227) and DynamicPointerNetwork m4c.py:27-33
 *   (mask_is_bool=1: uint8 mask, masked positions get -inf as in m4c.py:31). */
int ovqa_ptr_scores(const void* q, const void* k, const void* mask, float* scores,
                    int64_t B, int64_t T, int64_t L, int64_t d, int64_t ldq, int64_t ldk,
                    int64_t ld_scores, int64_t col_offset, float scale, int dtype, int mask_is_bool,
                    void* stream);

/* Backward of ovqa_ptr_scores (autograd of mmf_m4c.py:394-396; the M4C training path):
 *   dq[b,t,:] = scale * sum_j dscores[b,t,j] k[b,j,:],   dk[b,j,:] = scale * sum_t dscores[b,t,j] q[b,t,:];
 *   dscores fp32 rows of pitch ld_dscores (the gradient of the pointer columns, e.g. a slice of d(scores)); the additive mask is a
 *   constant, `bool_mask` (uint8 [B,L], or NULL) marks the keys ovqa_ptr_scores filled with -inf: they carry no gradient. */
int ovqa_ptr_scores_bwd(const float* dscores, const void* q, const void* k, const void* bool_mask,
                        void* dq, void* dk, int64_t B, int64_t T, int64_t L, int64_t d,
                        int64_t ld_dscores, int64_t ldq, int64_t ldk, int64_t lddq, int64_t lddk,
                        float scale, int dtype, void* stream);

/* PrevPredEmbeddings: out[b,t,:] = LN_src(table_row(prev_inds[b,t])) + LN_emb(pos_emb[t] + type_emb[is_ocr])
 *   where table_row(i) = ans_emb[i] (i < V, LayerNorm ans_*) else ocr_emb[b, i-V] (LayerNorm ocr_*).
 *   replaces mmf_m4c.py:416-459 without materialising the (B, V+Locr, D) table. fp32 in, fp32 out
 *   (row stride ld_out so it can write straight into the joint (B, L, D) encoder input). */
int ovqa_prev_pred_embed(const float* ans_emb, const float* ocr_emb, const int64_t* prev_inds,
                         const float* pos_emb, const float* type_emb,
                         const float* ans_g, const float* ans_b, const float* ocr_g, const float* ocr_b,
                         const float* emb_g, const float* emb_b,
                         float* out, int64_t B, int64_t T, int64_t V, int64_t Locr, int64_t D,
                         int64_t ld_out, int64_t out_batch_stride, float eps, void* stream);

/* One greedy step (mmf_m4c.py:247-256): argmax over the score row (first maximal index, as torch.argmax),
 *   prev_inds[b,1:] = argmax[b,:-1];  last_ids[b] = (last_ids[b]==eos) ? eos : argmax[b,step];
 *   *all_eos = 1 iff every last_ids[b]==eos (device flag; the host may poll it instead of syncing).
 *   argmax_out [B,T] int64 (required scratch/output). */
int ovqa_greedy_step(const float* scores, int64_t* prev_inds, int64_t* last_ids, int64_t* argmax_out,
                     int32_t* all_eos, int64_t B, int64_t T, int64_t W, int64_t step, int64_t eos_idx,
                     void* stream);

/* -------------------------------------------------------------------------------------------------
 * Incremental decoding: ONE MultiHeadAttention block of a decode step over a cache of PROJECTED keys / values.
 *   replaces: MultiHeadAttention.forward inside `statefulness()` (attentions.py:319-331 with :321-325's torch.cat of the raw
 *             running_keys / running_values and the full re-projection at :48-50), as driven by Decoder / DecoderLayer
 *             (decoders.py:21-26, :53-71) from BeamSearch.iter (beam_search.py:41-58).  Eval mode (dropout off).
 *   x_f32   [B*Lq, D] fp32: the block input (query rows AND residual).   x_op: the same rows in the compute dtype (bf16), or NULL (cast here).
 *   self_attention != 0: keys / values are the query rows (Lnew == Lq): one GEMM over the fused weight.
 *   self_attention == 0: kv_src_op [B*Lnew, D] (compute dtype) are the rows to append, or NULL with Lnew == 0 (attend over the cache as is:
 *                        encoder attention, whose keys were projected once).
 *   w_qkv [(2*h*dk + h*dv), D] = [Wq; Wk; Wv] row-major in the compute dtype, b_qkv fp32 [2*h*dk + h*dv];  w_o [D, h*dv], b_o [D].
 *   kv_cache [B, cap, h*dk + h*dv] compute dtype: row (b, j) = [K_j | V_j]; `len` positions are valid on entry; the call files the Lnew new
 *            rows at positions len .. len+Lnew-1 (the caller then advances len) and attends over all len+Lnew positions.
 *   mask: additive fp32, element (b, head, lq, j) at b*mask_sb + head*mask_sh + lq*mask_sq + j, j < len+Lnew; NULL = none.
 *   y_f32 [B*Lq, D] = LayerNorm(x + fc_o(attention));  y_op: optional second copy in bf16 (bf16 mode only).
 *   workspace: >= ovqa_mha_decode_workspace_bytes(...) bytes, 16-byte aligned, caller-owned; launches: (cast) + 1-2 GEMM + attention + GEMM + LN.
 * ------------------------------------------------------------------------------------------------- */
int64_t ovqa_mha_decode_workspace_bytes(int64_t B, int64_t Lq, int64_t Lnew, int64_t D, int64_t h, int64_t dk, int64_t dv, int dtype);
int ovqa_mha_decode_step(const float* x_f32, const void* x_op, const void* kv_src_op, int self_attention,
                         int64_t B, int64_t Lq, int64_t Lnew, int64_t D, int64_t h, int64_t dk, int64_t dv,
                         const void* w_qkv, const float* b_qkv, const void* w_o, const float* b_o,
                         const float* gamma, const float* beta, float eps,
                         void* kv_cache, int64_t cap, int64_t len,
                         const float* mask, int64_t mask_sb, int64_t mask_sh, int64_t mask_sq,
                         float scale, int dtype, void* workspace, int64_t workspace_bytes,
                         float* y_f32, void* y_op, void* stream);

/* MultiHeadAttention block, inference, ONE call (attentions.py:319-331 + :46-60 in eval mode): projection GEMM(s) over the fused
 *   [Wq; Wk; Wv] operand -> attention core -> fc_o + bias + residual -> LayerNorm.  Same arithmetic as the per-op entry points above; what it
 *   removes is host dispatch (one FFI call instead of six), which bounds the block below batch ~256.
 *   kind 0: self-attention (xk / xv ignored, Lk == Lq);  1: keys == values (xk [B*Lk, D]);  2: three inputs.
 *   x*_f32: fp32 rows (xq_f32 is also the residual); x*_op: the same rows in the compute dtype (bf16) or NULL (cast into the workspace);
 *           for dtype fp32 the *_op pointers are ignored and xk / xv are fp32.
 *   w_qkv [(2*h*dk + h*dv), D] = [Wq; Wk; Wv] row-major in the compute dtype, b_qkv fp32;  w_o [D, h*dv], b_o [D];  gamma / beta [D].
 *   mask: additive fp32 with element strides (0 = broadcast) or NULL.  prefix_*: optional structured description of the SAME mask (per-key
 *           row + causal tail, see ovqa_sdpa_fwd_prefix_lm): used when that kernel supports the shape, the dense mask otherwise.
 *   y_f32 [B*Lq, D] fp32, y_op: optional bf16 copy.  Workspace: ovqa_mha_block_infer_workspace_bytes. */
int64_t ovqa_mha_block_infer_workspace_bytes(int64_t B, int64_t Lq, int64_t Lk, int64_t D, int64_t h, int64_t dk, int64_t dv, int kind, int dtype);
int ovqa_mha_block_infer(const float* xq_f32, const void* xq_op, const void* xk, const void* xk_op, const void* xv, const void* xv_op,
                         int kind, int64_t B, int64_t Lq, int64_t Lk, int64_t D, int64_t h, int64_t dk, int64_t dv,
                         const void* w_qkv, const float* b_qkv, const void* w_o, const float* b_o, const float* gamma, const float* beta, float eps,
                         const float* mask, int64_t mask_sb, int64_t mask_sh, int64_t mask_sq,
                         const float* prefix_key_mask, int64_t prefix_sb, int64_t prefix_tail, float prefix_value,
                         float scale, int dtype, void* workspace, int64_t workspace_bytes, float* y_f32, void* y_op, void* stream);

/* PositionWiseFeedForward in eval mode from one call (positionwise_feed_forward.py:23-27, dropout off):
 *   y = LayerNorm(x + fc2(gelu(fc1 x)));  x_f32 [rows, D], x_op as above;  w1 [d_ff, D], w2 [D, d_ff] compute dtype; biases fp32. */
int64_t ovqa_pwff_infer_workspace_bytes(int64_t rows, int64_t D, int64_t d_ff, int dtype);
int ovqa_pwff_infer(const float* x_f32, const void* x_op, int64_t rows, int64_t D, int64_t d_ff,
                    const void* w1, const float* b1, const void* w2, const float* b2,
                    const float* gamma, const float* beta, float eps, int dtype,
                    void* workspace, int64_t workspace_bytes, float* y_f32, void* y_op, void* stream);

/* Beam selection: the k largest entries of every row, in descending order (NaN greatest, equal values by ascending index).
 *   replaces: BeamSearch.select (beam_search.py:35-38: torch.sort over all beam x vocabulary candidates of a sample, then [:beam_size]).
 *   x fp32 [rows, n] (row pitch ldx);  out_val fp32 [rows, k], out_idx int64 [rows, k];  1 <= k <= min(16, n), rows <= 65535
 *   (OVQA_ERR_BAD_SHAPE otherwise).
 *   workspace >= ovqa_topk_rows_workspace_bytes(rows, n, k) bytes, caller-owned;  two launches. */
int64_t ovqa_topk_rows_workspace_bytes(int64_t rows, int64_t n, int k);
int ovqa_topk_rows(const float* x, int64_t rows, int64_t n, int64_t ldx, int k, float* out_val, int64_t* out_idx,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* Verification twin of ovqa_gemm's bf16 path on CUDA cores (same operand forms, bias/GELU epilogue): lets the
 * GPU tests cross-check the tcgen05 kernel on device.  Not on the product path. */
int ovqa_gemm_bf16_simt(const void* A, const void* B, void* C, int64_t m, int64_t n, int64_t k,
                        int64_t lda, int64_t ldb, int64_t ldc, int a_op, int b_op, int c_dtype,
                        float alpha, const float* bias, int epi_flags, void* stream);

/* Diagnostics for tests: fill `out[0..n)` with the dropout keep-mask (1/0 as float) the epilogues use. */
int ovqa_dropout_mask(float* out, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_base,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OVQA_H_ */
