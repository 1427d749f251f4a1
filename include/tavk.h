/* tavk.h — C ABI of libtavk.so, the sm_100a kernel library behind the TAV fusion hot path.
 *
 * The reference (g8a9/multi-modal-emotion) is pure PyTorch: it has no FFI/plugin layer of its own.  The drop-in
 * boundary is therefore the nn.Module API (TAVForMAE / PreFormer / VideoMAEEncoder / TransformerEncoder) and each
 * entry point below replaces one family of torch op sites inside those modules.  Citations are
 * <reference file>:<line> relative to the reference repository root.
 *
 * Conventions (every function):
 *   - raw device pointers + explicit sizes/strides; no torch / C++ types cross this boundary;
 *   - `stream` is a cudaStream_t passed as void*; kernels are only enqueued, never synchronised;
 *   - no allocation and no hidden global state: outputs and workspaces are caller-owned;
 *   - returns 0 on success; 1 = bad argument, 2 = unsupported shape, 3 = CUDA error.  tavk_last_error()
 *     returns a thread-local message for the last non-zero return;
 *   - there is NO CPU fallback: without an sm_100 device every compute entry point returns 3.
 */
#ifndef TAVK_H_
#define TAVK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAVK_VERSION 100

enum { TAVK_F32 = 0, TAVK_BF16 = 1 };

/* GEMM epilogues */
enum {
    TAVK_EPI_LINEAR = 0,   /* out = alpha*acc (+bias) (+rowbias) (+resid)                               */
    TAVK_EPI_GELU = 1,     /* out = bf16(pre), out2 = bf16(gelu_erf(pre)), pre = alpha*acc + bias      */
    TAVK_EPI_GELU_BWD = 2, /* out = bf16((alpha*acc + bias) * gelu_erf'(aux)); bf16 output only             */
    TAVK_EPI_GELU_GRAD = 3,/* out = bf16(gelu_erf'(pre)), out2 = bf16(gelu_erf(pre)): the forward saves the derivative
                              instead of the pre-activation ...                                              */
    TAVK_EPI_MUL = 4       /* ... so that the backward is out = bf16((alpha*acc + bias) * aux), aux = that derivative */
};

/* attention mask modes */
enum {
    TAVK_ATTN_NONE = 0,     /* softmax(s*QK^T) V                       (HF Wav2Vec2 / VideoMAE encoder layers)  */
    TAVK_ATTN_KEY_BIAS = 1  /* softmax(s*QK^T + bias[b,k]) V           (utils/TAVFormer.py:68-72, HF RoBERTa)   */
    /* The fusion encoder's post-softmax mask add (utils/TAVFormer.py:372-375) is NOT a softmax mask:
     * P+m => ctx = softmax(s*QK^T)V + sum_k m[b,k] V[b,k,:].  It is served by TAVK_ATTN_NONE plus
     * tavk_masked_colsum (the rank-1 term) routed through the out-projection as an fp32 row-bias. */
};

const char* tavk_last_error(void);
int tavk_version(void);
/* 0 when an sm_100 device is current and usable, 3 otherwise (message in tavk_last_error). */
int tavk_device_check(void);
int tavk_sm_count(void);
/* Caller-owned scratch sizes (bytes): tavk_attn_bwd's `delta`, the GroupNorm entry points' `sums_ws`, and the GEMM (none:
 * accumulators live in tensor memory, split-K partial sums are reduced into the output). */
int64_t tavk_workspace_bytes_attn_bwd(int B, int S, int nh);
int64_t tavk_workspace_bytes_groupnorm(int B, int C);
struct tavk_gemm_args;
int64_t tavk_workspace_bytes_gemm(const struct tavk_gemm_args* args);

/* ------------------------------------------------------------------------------------------------------------
 * tavk_gemm_bf16 — C[M,N] = epi(alpha * sum_k A[m,k] * B[n,k]); bf16 operands, fp32 accumulate (tcgen05/TMEM, TMA).
 * Replaces: F.linear for key/value/query (utils/TAVFormer.py:348-350, fused to N=2304), out-proj dense (:422),
 * intermediate dense + GELU (:404-405), output dense + residual (:434-437), wav_2_768 / wav_2_768_2
 * (models/tav.py:363,478), MultiHeadAttention query/key/value/out (utils/TAVFormer.py:44-49,88) and the HF encoder
 * layers' Linear modules; plus the autograd backward of each (dgrad / wgrad) through the *_mn_major flags.
 *   K-major operand : element (row r, k) at ptr[r*ld + k]          (contraction index contiguous)
 *   MN-major operand: element (row r, k) at ptr[k*ld + r]          (row index contiguous)
 * forward   Y = X W^T      : A=X  (K-major)      B=W   (K-major)
 * dgrad     dX = dY W      : A=dY (K-major)      B=W   (MN-major, K := N_out)
 * wgrad     dW = dY^T X    : A=dY (MN-major)     B=X   (MN-major, K := rows), k_splits>1 + accumulate
 * Requirements: N % 8 == 0, lda/ldb % 8 == 0, 16-byte aligned bases; K tail and M/N edges are handled
 * (TMA zero-fill on loads, predicated stores). */
typedef struct tavk_gemm_args {
    const void* A;  int64_t lda;  int32_t a_mn_major;
    const void* B;  int64_t ldb;  int32_t b_mn_major;
    int32_t M, N, K;
    void* out;      int64_t ldo;  int32_t out_dtype;     /* TAVK_F32 | TAVK_BF16                                 */
    void* out2;     int64_t ldo2;                        /* bf16, TAVK_EPI_GELU only                             */
    const float* bias;                                   /* [N] or NULL                                          */
    const float* resid; int64_t ldr;                     /* f32 [M,N] residual or NULL                           */
    const float* rowbias; int32_t rows_per_group;        /* f32 [ceil(M/rows_per_group), N] or NULL              */
    const void* aux;  int64_t ldaux;                     /* bf16 [M,N] pre-activation, TAVK_EPI_GELU_BWD only    */
    float* colsum;                                       /* f32 [N] or NULL: += column sums of the stored result
                                                            (the bias gradient of the Linear that produced the
                                                            GEMM's A operand), accumulated with red.global.add  */
    int32_t epilogue;                                    /* TAVK_EPI_*                                           */
    int32_t accumulate;                                  /* 1: out += result (f32 out, red.global.add)           */
    int32_t k_splits;                                    /* >=1; >1 requires accumulate                          */
    int32_t block_n;                                     /* 0 = heuristic, or force 64 / 128 / 256               */
    float alpha;
    /* ---- grouped / convolution-walk extension (all zero = one plain GEMM) ----
     * `groups` > 1 runs that many independent [M,N,K] problems in one launch over the SAME two tensor maps; group g
     * shifts the TMA coordinates by g*a_g_mn / g*a_g_k (A: row-or-mn coordinate, k coordinate), g*b_g_mn / g*b_g_k
     * (B) and writes its [M,N] block at rows +g*out_g_row, columns +g*out_g_col of out/out2/resid/aux (bias and colsum
     * are indexed by the output column).
     * a_kstep (K-major A): elements between consecutive 64-wide k-blocks (0 = 64).  With a_kstep = lda a k-block is
     * "the next row": a sliding-window (stride-1) convolution tap walk over channels-last activations.
     * b_box_k_shift (MN-major B): every 64-wide box of the N extent uses the same mn coordinate and k coordinate
     * + (box index)*shift: the wgrad of that convolution (box = tap).
     * a_rows/a_cols/b_rows/b_cols: extents of the row-major matrices the tensor maps describe (0 = derived from
     * M, N, K); cols may exceed ld (overlapping rows).  The strided Conv1d layers of the Wav2Vec2 feature encoder
     * are plain GEMMs with lda = stride*C, K = kernel*C.  See multi-modal-emotion_b200/frontends.py. */
    int32_t groups;
    int32_t a_kstep;
    int32_t a_g_mn, a_g_k, b_g_mn, b_g_k;
    int32_t b_box_k_shift;
    int32_t out_g_row, out_g_col;
    int64_t a_rows, a_cols, b_rows, b_cols;
    /* Per-call SM budget of the persistent grid (0 = every SM).  A data-parallel host passes tavk_sm_count() - n to keep
     * n SMs free for the NCCL all-reduce kernels that run concurrently with backward; there is no process-wide state. */
    int32_t max_ctas;
    /* Optional dynamic tile scheduler: int32[2] of caller-owned device memory, zero before the first use (the kernel
     * leaves it zero again when it finishes, so consecutive launches on one stream may share it; launches that can run
     * CONCURRENTLY need one each).  NULL = static round-robin tiles.  With it, persistent CTAs claim output tiles from a
     * counter: a CTA whose SM is still held by another stream's kernel or an NCCL collective claims fewer tiles
     * instead of holding the whole grid back. */
    void* sched_workspace;
    /* CTA pairs (tcgen05 cta_group::2: two CTAs of one cluster share a 256-row output tile and each loads half of the B
     * tile): 0 = the library decides (plain problems with >= 4 row blocks), 1 = never, 2 = whenever the problem allows it
     * (ungrouped, no convolution walk, N > 64, M > 128).  Pair launches use the static tile order (sched_workspace unused). */
    int32_t cta_pair;
} tavk_gemm_args;
int tavk_gemm_bf16(const tavk_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused attention, head_dim 64 (flash-style, online softmax in registers, no [S,S] materialisation).
 * Replaces utils/TAVFormer.py:357-387 (VideoMAESelfAttention scores/softmax/PV), :57-86 (MultiHeadAttention)
 * and the HF encoders' attention interface.  q/k/v/o are bf16 with layout [B, S, nh, 64] at arbitrary row
 * strides (elements): element (b,s,h,d) at ptr[(b*S + s)*ld + h*64 + d]  — so a packed [B,S,3,nh,64] QKV buffer
 * is addressed with ld = 3*nh*64 and three base pointers.  key_bias: f32 [B,S] (TAVK_ATTN_KEY_BIAS) or NULL.
 * lse: f32 [B, nh, S] log-sum-exp of the scaled (+biased) scores, saved for backward. */
typedef struct tavk_attn_args {
    const void* q; const void* k; const void* v; int64_t ld_qkv;
    void* o; int64_t ld_o;
    float* lse;
    const float* key_bias;
    int32_t B, S, nh;
    int32_t mode;          /* TAVK_ATTN_*  */
    float scale;           /* 1/sqrt(64) in every reference call site */
} tavk_attn_args;
int tavk_attn_fwd(const tavk_attn_args* args, void* stream);

/* Backward of the above.  dq/dk/dv bf16 [B,S,nh,64] at row stride ld_dqkv.  delta: f32 [B,nh,S] scratch
 * (rowsum(dO∘O)), written then read by this call.  Rank-1 extra term of the fusion encoder's post-softmax mask
 * (SURVEY Q1): when dv_rowscale (f32 [B,S]) and dv_rank1 (f32 [B, nh*64]) are non-NULL,
 * dV[b,k,h,:] += dv_rowscale[b,k] * dv_rank1[b, h*64:(h+1)*64]. */
typedef struct tavk_attn_bwd_args {
    const void* q; const void* k; const void* v; int64_t ld_qkv;
    const void* o; const void* d_o; int64_t ld_o;
    const float* lse; float* delta;
    const float* key_bias;
    void* dq; void* dk; void* dv; int64_t ld_dqkv;
    const float* dv_rowscale; const float* dv_rank1;
    int32_t B, S, nh;
    int32_t mode;
    float scale;
    /* optional f32 [nh*64] each: += column sums of dq / dk / dv over all B*S rows — the q/k/v projection bias
     * gradients — accumulated in the backward kernels' epilogues (tensor-core path) or by column-sum passes (v1) */
    float* dbq; float* dbk; float* dbv;
} tavk_attn_bwd_args;
int tavk_attn_bwd(const tavk_attn_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (H in {768, 1024}); fp32 statistics, two-pass variance.
 * Replaces nn.LayerNorm at utils/TAVFormer.py:237,239,108,118 and models/tav.py:439,443,445,447 (+ HF layers).
 * x f32 [M,H] -> y_bf16 (optional) and/or y_f32 (optional); mean/rstd f32 [M] saved for backward. */
int tavk_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean,
                       float* rstd, int M, int H, float eps, void* stream);
/* dx = (dx_resid or 0) + LN'(dy); dgamma += sum dy*xhat; dbeta += sum dy (atomic accumulation into f32 [H]).
 * dy is f32 [M,H]; dx_f32 and/or dx_bf16 (a bf16 copy for the next dgrad/wgrad GEMM) may be NULL.
 * dx_colsum (f32 [H] or NULL): += column sums of dx — the bias gradient of the Linear whose output x is, so no
 * separate pass over dx is needed for it. */
int tavk_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                       const float* dx_resid, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                       float* dx_colsum, int M, int H, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Memory-bound helpers of the path. */
/* y[b,s,:] = x[b,s,:] + table[idx[b,s],:]  — models/tav.py:474 (hidden_states + embedding(pos_embed)); idx int64. */
int tavk_embed_add_fwd(const float* x, const int64_t* idx, const float* table, float* y, int rows, int H, int n_embed,
                       void* stream);
/* dtable[j,:] += sum_{rows: idx==j} dy[row,:]  (n_embed <= 8). */
int tavk_embed_add_bwd(const float* dy, const int64_t* idx, float* dtable, int rows, int H, int n_embed, void* stream);
/* HF RobertaEmbeddings.forward up to its LayerNorm (models/tav.py:349 `bert.embeddings(input_ids)`, and inside `bert(...)`
 * at :485): y[b,t,:] = word[ids[b,t]] + pos[p[b,t]] + type0, with p = pad_id + (count of non-pad tokens up to t) for
 * non-pad tokens and pad_id for pad tokens; pos_ids (int64 [B,T]) receives p for the backward scatter. */
int tavk_roberta_embed_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, float* y,
                           int64_t* pos_ids, int B, int T, int H, int vocab, int n_pos, int pad_id, void* stream);
/* dtable[idx[r],:] += dy[r,:] for r < rows (idx outside [0,n_embed) and idx == skip_idx skipped; skip_idx = the
 * nn.Embedding padding_idx, whose row receives no gradient, or -1): embedding backward as a row scatter into a
 * caller-owned (pre-zeroed or accumulating) table gradient. */
int tavk_embedding_scatter_add(const float* dy, const int64_t* idx, float* dtable, int rows, int H, int n_embed,
                               int skip_idx, void* stream);
/* y[b,:] = (1/S) sum_s x[b,s,:]  — torch.mean(dim=1) at models/tav.py:478,481,488 (unmasked). */
int tavk_mean_pool_fwd(const float* x, float* y, int B, int S, int H, void* stream);
/* dx[b,s,:] = dy[b,:] / S ; also emits a bf16 copy when dx_bf16 != NULL. */
int tavk_mean_pool_bwd(const float* dy, float* dx, void* dx_bf16, int B, int S, int H, void* stream);
/* Masked form (SURVEY 8b "optional mask/lengths"): y[b,:] = (1/len_b) sum_{s < len_b} x[b,s,:] with len_b = lengths[b]
 * clamped to [0,S] (int32 [B]; NULL = the unmasked mean above; len_b = 0 gives zeros); backward: dx[b,s,:] =
 * dy[b,:]/len_b for s < len_b, else 0. */
int tavk_masked_mean_pool_fwd(const float* x, const int* lengths, float* y, int B, int S, int H, void* stream);
int tavk_masked_mean_pool_bwd(const float* dy, const int* lengths, float* dx, void* dx_bf16, int B, int S, int H,
                              void* stream);
/* out[n] (+)= sum_m x[m,n]; x is bf16 or f32 [M,N] (row stride ld) — bias gradients. */
int tavk_colsum(const void* x, int x_dtype, int64_t ld, float* out, int M, int N, int accumulate, void* stream);
/* out[b,n] = sum_s w[b,s] * x[b,s,n]  (w NULL => 1) ; x bf16/f32 [B,S,N] row stride ld — the rank-1 term
 * c[b,:] = sum_k m[b,k] V[b,k,:] of utils/TAVFormer.py:375,383 and its backward reductions. */
int tavk_masked_colsum(const void* x, int x_dtype, int64_t ld, const float* w, float* out, int B, int S, int N,
                       void* stream);
/* Small dense fp32 linear for launch-bound shapes (classifier head models/tav.py:499; W_o·c GEMV):
 * y[m,n] = sum_k x[m,k] w[n,k] + b[n]      (fwd)
 * dx[m,k] = sum_n dy[m,n] w[n,k]           (bwd_x)
 * dw[n,k] += sum_m dy[m,n] x[m,k]; db[n] += sum_m dy[m,n]   (bwd_w) */
int tavk_small_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int N, int K, void* stream);
int tavk_small_linear_bwd_x(const float* dy, const float* w, float* dx, int M, int N, int K, int accumulate,
                            void* stream);
int tavk_small_linear_bwd_w(const float* dy, const float* x, float* dw, float* db, int M, int N, int K, void* stream);
/* f32 -> bf16 cast of a flat buffer (weight shadows, activations). */
int tavk_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);
/* y = x * scale (f32, in place allowed). */
int tavk_scale_f32(const float* x, float* y, float scale, int64_t n, void* stream);
/* y[r, c] = keep[r,c] ? x[r,c]/(1-p) : 0 with a counter-based RNG (seed, offset): head dropout models/tav.py:497-498.
 * offset_dev (optional, device uint64): a step counter added (<<32) to `offset` on the device, so a captured CUDA
 * graph draws a fresh mask on every replay. */
int tavk_dropout(const float* x, float* y, uint8_t* keep_mask, int64_t n, float p, uint64_t seed, uint64_t offset,
                 const uint64_t* offset_dev, void* stream);
int tavk_dropout_bwd(const float* dy, const uint8_t* keep_mask, float* dx, int64_t n, float p, void* stream);
/* dx = resid + dropout_bwd(dy): the feed-forward input dropout of the reference TransformerBlock sits on a branch whose
 * un-dropped source also feeds the residual (utils/TAVFormer.py:111,136). */
int tavk_dropout_bwd_add(const float* dy, const uint8_t* keep_mask, const float* resid, float* dx, int64_t n, float p,
                         void* stream);
/* bf16 [B,S,nh,d] -> [B,nh,d,S] (inverse=0) or back (inverse=1): the reference MultiHeadAttention "concat" reinterprets
 * a [B,nh,d,S] buffer as [B,S,nh*d] (utils/TAVFormer.py:86); backward uses the inverse permutation. */
int tavk_permute_bshd_bhds(const void* in, void* out, int B, int S, int nh, int d, int inverse, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Wav2Vec2 convolutional feature encoder, first layer, channels-last (HF Wav2Vec2GroupNormConvLayer: Conv1d(1, C, k,
 * stride s, bias optional) -> GroupNorm(C groups, eps) -> GELU; reached from the reference at models/tav.py:352
 * (wav2vec2.feature_extractor) and :476 (wav2vec2(...))).  Activations are bf16 [B, R, C] with R >= T padded rows
 * per sample (rows t >= T are zero); the layers after it run on tavk_gemm_bf16 with lda = stride*C (overlapping TMA
 * rows), see multi-modal-emotion_b200/frontends.py.
 *   conv0_fwd           : u[b,t,c] = sum_j wav[b, s*t+j] w[c,j] (+bias[c])          wav f32 [B,L], w f32 [C,k]
 *   groupnorm_gelu_fwd  : per (b,c) mean/rstd over t < T (shifted one-pass sums, f32); z = gamma*xhat+beta,
 *                         a = gelu_erf(z); sums_ws: f32 [B,2,C] workspace
 *   groupnorm_bwd       : from dz (gradient w.r.t. z, bf16): du = gamma*rstd*(dz - mean_t dz - xhat*mean_t(dz*xhat))
 *                         (bf16, may alias dz; zero in the padding rows) and sums_ws[b,0,c] = sum_t dz (-> dbeta),
 *                         sums_ws[b,1,c] = sum_t dz*xhat (-> dgamma)
 *   wave_windows        : win[b*R+t, j] = bf16(wav[b, s*t+j]) (j < k, else 0), bf16 [B*R, 16]: with it the conv0 weight
 *                         gradient is the wgrad GEMM du^T x win; the waveform needs no gradient. */
int tavk_conv0_fwd(const float* wav, const float* w, const float* bias, void* u_bf16, int B, int L, int R, int T, int C,
                   int k, int s, void* stream);
int tavk_groupnorm_gelu_fwd(const void* u_bf16, const float* gamma, const float* beta, void* z_bf16, void* a_bf16,
                            float* mean, float* rstd, float* sums_ws, int B, int R, int T, int C, float eps,
                            void* stream);
int tavk_groupnorm_bwd(const void* dz_bf16, const void* u_bf16, const float* mean, const float* rstd,
                       const float* gamma, void* du_bf16, float* sums_ws, int B, int R, int T, int C, void* stream);
int tavk_wave_windows(const float* wav, void* win_bf16, int B, int L, int R, int T, int k, int s, void* stream);
/* Layer-norm family of the same encoder (HF Wav2Vec2LayerNormConvLayer: Conv1d + bias -> LayerNorm over the C channels ->
 * GELU; the wav2vec2-large checkpoints the reference names, models/tav.py:257,455).  Channels-last bf16 [B, R, C] rows,
 * C % 256 == 0, C <= 1024:
 *   chan_ln_gelu_fwd : a = gelu_erf(gamma*(u-mean)*rstd + beta) per row; mean/rstd f32 [B*R] saved; padding rows zero
 *   chan_ln_gelu_bwd : du = LN'(gelu'(z)*da) (z recomputed), dgamma/dbeta accumulated (f32 [C], atomics) */
int tavk_chan_ln_gelu_fwd(const void* u_bf16, const float* gamma, const float* beta, void* a_bf16, float* mean, float* rstd,
                          int B, int R, int T, int C, float eps, void* stream);
int tavk_chan_ln_gelu_bwd(const void* da_bf16, const void* u_bf16, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, void* du_bf16, float* dgamma, float* dbeta, int B, int R, int T, int C,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Weighted softmax cross-entropy (utils/global_functions.py:63-64,76,83 — nn.CrossEntropyLoss(weight=w), mean).
 * logits f32 [B,C], target int64 [B], class_weight f32 [C] or NULL (all ones).
 * Emits loss_num = sum_i w[y_i]*l_i and loss_den = sum_i w[y_i] separately (data-parallel ranks all-reduce the
 * denominator) plus per-row softmax probs (f32 [B,C]) for backward. */
int tavk_softmax_ce_fwd(const float* logits, const int64_t* target, const float* class_weight, float* probs,
                        float* loss_num, float* loss_den, int B, int C, void* stream);
/* dlogits[i,:] = gscale * w[y_i] * (probs[i,:] - onehot(y_i)) ; gscale = dL / loss_den is read from device memory
 * (*gscale_dev) so no host sync is needed. */
int tavk_softmax_ce_bwd(const float* probs, const int64_t* target, const float* class_weight, const float* gscale_dev,
                        float* dlogits, int B, int C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser step over FLAT fp32 buffers (train_model/tav_train.py:61-62: clip_grad_norm_ + AdamW.step).
 * tavk_grad_sqnorm: *out (+)= sum g^2 (f32 accumulate in fp32 with per-block Kahan-free tree; out must be zeroed).
 * tavk_adamw: p,m,v,g flat f32 [n]; clip scale = min(1, max_norm/(sqrt(*sqnorm_dev)+1e-6)) computed on device when
 * sqnorm_dev != NULL (max_norm<=0 disables); decoupled weight decay; bias correction with `step` (1-based);
 * optionally writes the bf16 shadow of the updated weights (the GEMM operands) in the same pass and zeroes g. */
int tavk_grad_sqnorm(const float* g, int64_t n, float* out, void* stream);
int tavk_adamw(float* p, float* m, float* v, float* g, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
               float eps, float weight_decay, int step, const float* sqnorm_dev, float max_norm, float grad_prescale,
               int zero_grad, void* stream);
/* CUDA-graph-safe form of the same step (train_model/tav_train.py:61-63,148-149: AdamW's bias correction and the
 * CosineAnnealingWarmRestarts learning rate both advance EVERY iteration, so neither may be a by-value kernel argument
 * of a captured launch).  The optimiser clock lives in caller-owned device memory:
 *   step_dev  int32[1]  number of completed steps;
 *   hyper_dev f32[4]    {lr, 1-beta1^step, sqrt(1-beta2^step), unused}; the host writes hyper_dev[0] (the scheduler's
 *                       current learning rate) before each step / graph replay.
 * tavk_adamw_prep: ++*step_dev, recomputes hyper_dev[1..2] (double precision) and zeroes *sqnorm_dev (NULL allowed) —
 * it takes the place of the memset in front of tavk_grad_sqnorm.  tavk_adamw_dev: tavk_adamw with lr and the bias
 * corrections read from hyper_dev. */
int tavk_adamw_prep(int* step_dev, float* hyper_dev, float* sqnorm_dev, double beta1, double beta2, void* stream);
int tavk_adamw_dev(float* p, float* m, float* v, float* g, void* p_bf16, int64_t n, const float* hyper_dev, double beta1,
                   double beta2, float eps, float weight_decay, const float* sqnorm_dev, float max_norm,
                   float grad_prescale, int zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAVK_H_ */
