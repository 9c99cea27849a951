/* sdumc_b200 — C ABI of the B200-native SDUMC hot path (libsdumc_b200.so).
 *
 * The reference (WarmCongee/SDUMC) has no FFI: its boundary is the Python nn.Module API
 *   toolkit/models/wengnet_mosei_mult_views_text_missing.py:186-370 (model)
 *   toolkit/utils/loss.py:19-51,243-315                              (MSE / RMSE / RnC losses)
 *   main_frame_val_text_missing.py:89-158,317-321                    (train step, Adam)
 * Each entry point below names the reference lines it replaces.  sdumc_b200/_lib.py is the ctypes
 * binding; toolkit/ at the repo root re-exports the reference-named Python classes on top of it.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (sdumc_last_error() = thread-local text);
 *     no C++ exception crosses the boundary;
 *   - all pointers are DEVICE pointers unless named host_*; the caller owns every buffer,
 *     including workspaces whose size is queried with the *_workspace_bytes call;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the device;
 *   - matrices are row-major; `ld*` are leading dimensions in elements.
 */
#ifndef SDUMC_B200_H_
#define SDUMC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDUMC_ABI_VERSION 1

int sdumc_version(void);
const char* sdumc_last_error(void);

/* ------------------------------------------------------------------------------------------
 * 1. tcgen05 GEMM operator  C = epilogue(op(A) op(B))
 *    replaces every nn.Linear / torch.bmm site of the model (reference :282-284, :60, :82, :85,
 *    MLP() :264-273) and their autograd backward (dX = dY W, dW = dY^T X).
 * ------------------------------------------------------------------------------------------ */
enum { SDUMC_ACT_NONE = 0, SDUMC_ACT_RELU = 1, SDUMC_ACT_TANH = 2 };
enum { SDUMC_OUT_STORE = 0, SDUMC_OUT_ADD = 1, SDUMC_OUT_ATOMIC = 2 };
enum { SDUMC_EPI_GENERIC = 0, SDUMC_EPI_INPROJ = 1, SDUMC_EPI_KEYPROJ = 2 };

typedef struct sdumc_gemm_desc {
  int32_t M, N, K;
  int32_t a_mn;     /* 0: A stored [M,K] (K contiguous); 1: A stored [K,M] */
  int32_t b_mn;     /* 0: B stored [N,K] (nn.Linear weight);  1: B stored [K,N] */
  int32_t tf32;     /* 0: bf16 operands; 1: fp32 operands multiplied as tf32 */
  int32_t k_splits; /* >1: split the reduction, fp32 atomics */
  int32_t block_n;  /* 0 = auto, else 64/128/256 */
  int32_t max_ctas; /* 0 = one per SM */
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  int32_t epi_kind;
  int32_t act;
  const float* bias; /* [N] or NULL */
  const float* gate; /* [M,N] fp32 or NULL: C *= gate_scale * (gate > 0)  (ReLU+dropout backward) */
  int64_t ld_gate;
  float gate_scale;
  float drop_p; /* element dropout after the activation */
  uint32_t drop_site;
  uint32_t fmask_site; /* != 0: multiply by the p=0.5 frame mask of this site */
  uint32_t fmask_site2; /* with fmask_split > 0: rows >= fmask_split take the mask of this site at row r - fmask_split */
  int64_t fmask_split;  /* (two tensors with their own dropout sites stacked along M: the passes of a modality) */
  float* out_f32;
  int64_t ld_f32;
  int32_t f32_mode;
  void* out_bf16;
  int64_t ld_bf16;
  int32_t bf16_mode;
  int32_t n_tgt; /* in-proj epilogue: dropped bf16 copies */
  void* tgt[4];
  uint32_t tgt_site[4];
  const float* qv; /* key-proj epilogue: queries [n_samples, nq, N] (q_stride = nq*N) or shared (0) */
  int64_t q_stride;
  int32_t nq;
  int32_t L;
  float* scores; /* [M, nq] */
  uint64_t seed; /* dropout RNG */
  uint32_t step;
  const uint32_t* step_dev; /* optional device counter added to step */
} sdumc_gemm_desc;

int sdumc_gemm(const sdumc_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * 2. Dropout RNG: Philox4x32-10 keyed by (seed, step); `site` numbers the dropout site.
 *    Test hooks materialise the masks the kernels apply (values 0 or 1/(1-p)).
 * ------------------------------------------------------------------------------------------ */
typedef struct sdumc_dropkey {
  uint32_t seed_lo, seed_hi, step;
  uint32_t reserved;
  const uint32_t* step_dev; /* optional device counter added to `step` at run time (CUDA-graph replay) */
} sdumc_dropkey;

int sdumc_frame_mask(uint64_t seed, uint32_t step, uint32_t site, int64_t rows, int32_t cols, float* out,
                     void* stream);
int sdumc_elem_mask(uint64_t seed, uint32_t step, uint32_t site, int64_t n, float p, float* out, void* stream);

#if defined(__CUDACC__) && defined(SDUMC_INTERNAL)
#define SDUMC_BF16 __nv_bfloat16
#else
#define SDUMC_BF16 uint16_t /* raw bfloat16 bits */
#endif

/* ------------------------------------------------------------------------------------------
 * 3. Pooling attention, frame level (FRA2UTT_new.forward :56-68, Cross_Attention.forward :79-95)
 *    forward : scores come from sdumc_gemm(SDUMC_EPI_KEYPROJ); this op does softmax over the L
 *              frames of each sample, the weighted pooling and the output dropout.
 *    backward: row-wise part (dS, dZ, dQp, db_in, value path); the dense parts are sdumc_gemm.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdumc_pool_fwd_args {
  const SDUMC_BF16* X; /* X' [B*L,256] dropped frames, row pitch ldx */
  int64_t ldx;
  float* S;            /* [B*L,nq] in: scores (unless Kt is set), out: attention probabilities */
  int32_t B, L, nq;
  float alpha;         /* softmax_scale = 0.3 */
  float* O_pre;        /* [B,nq,256] pooled output before the output dropout */
  float* out;          /* out[b*out_stride_b + q*256 + g] */
  int64_t out_stride_b;
  SDUMC_BF16* out_bf16; /* optional copy, same indexing */
  float drop_p;
  uint32_t site;
  sdumc_dropkey key;
  const SDUMC_BF16* Kt; /* optional tanh keys [B*L,256]: when set, the scores S = Kt Qp^T are computed here */
  int64_t ldk;
  const float* Qp;      /* with Kt: projected queries [B,nq,G] (qp_stride_b = nq*G) or shared context (0) */
  int64_t qp_stride_b;
  int32_t G;            /* general_dim: 0 or 256 (reference), or 1024 (BASELINE config 4); 256 in the comments above */
  int32_t reserved;
  /* Length-aware (varlen) eval execution with the reference's padding semantics.  The reference
   * right-zero-pads every batch and has NO mask (read_data.py:223-248; `pads` ignored, feat_data.py:239,253): a padded
   * frame still takes part in the softmax with h_pad = in-projection bias, k_pad = tanh(W_in h_pad + b_in).  With
   * row_off set, X / Kt / S hold only the VALID frames, packed: sample b owns rows [row_off[b], row_off[b+1]); L is the
   * padded length the reference would see, and the (L - T_b) padded frames enter in closed form: the denominator gains
   * (L - T_b) exp(alpha s_pad,q) and the pooled value (L - T_b) p_pad,q h_pad.  Eval mode only (input dropout makes
   * padded rows differ in train mode: those batches run padded).  Qp is required (s_pad,q = k_pad . Qp_q). */
  const int32_t* row_off;   /* [B+1] device, or NULL = dense padded layout */
  const SDUMC_BF16* Hpad;   /* [G] */
  const SDUMC_BF16* Kpad;   /* [G] */
} sdumc_pool_fwd_args;
int sdumc_pool_fwd(const sdumc_pool_fwd_args* a, void* stream);

typedef struct sdumc_attn_bwd_args {
  const SDUMC_BF16* X;
  int64_t ldx;
  const SDUMC_BF16* Kt; /* tanh(X' W_in^T + b) [B*L,256] */
  int64_t ldk;
  const float* P;       /* [B*L,nq] */
  const float* dOut;    /* gradient of the dropped pooled output: dOut[b*dout_stride_b + q*256 + g] */
  int64_t dout_stride_b;
  const float* O_pre;
  const float* Qp;      /* projected queries [B,nq,256] (qp_stride_b = nq*256) or shared context (0) */
  int64_t qp_stride_b;
  int32_t B, L, nq;
  float alpha;
  float out_drop_p;
  uint32_t out_site;
  SDUMC_BF16* dZ;       /* [B*L,256] */
  int64_t lddz;
  SDUMC_BF16* dH;       /* value-path gradient, masked by fmask_site */
  int64_t lddh;
  int32_t dh_mode;      /* 0 store, 1 accumulate */
  uint32_t fmask_site;  /* 0: no input dropout */
  float* dQp;           /* accumulated with atomicAdd (zero-fill first): [B,nq,256], or [nq,256] when qp_stride_b == 0 */
  int64_t dqp_stride_b;
  float* db;            /* [G] atomicAdd */
  sdumc_dropkey key;
  int32_t G;            /* general_dim: 0 or 256 (reference), or 1024 */
  int32_t max_ctas;     /* 0 = one persistent CTA per SM; > 0: at most this many (a share of the GPU, so that the
                           blocks of several modalities run side by side) */
  int32_t split_b;      /* > 0: two problems with their own dropout sites stacked along B (the passes of a modality):
                           samples >= split_b use out_site2 / fmask_site2 with the sample index b - split_b */
  uint32_t out_site2;
  uint32_t fmask_site2;
  uint32_t reserved2;
} sdumc_attn_bwd_args;
int sdumc_attn_bwd(const sdumc_attn_bwd_args* a, void* stream);

int sdumc_cast_bf16(const float* src, SDUMC_BF16* dst, int64_t n, void* stream);
int sdumc_colsum_bf16(const SDUMC_BF16* X, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream);

/* Collate of a device-resident feature store: the right-zero-padding batch builder of
 * toolkit/utils/read_data.py:223-248 (pad_to_maxlen_pre_modality_tensor_4) as a gather.
 *   packed      all utterances of one modality, rows back to back: bf16 [sum_T, D]
 *   row_offset  [n_utt + 1] first row of each utterance (device, int64)
 *   idx         [b] utterances of the batch (device, int32)
 *   out         bf16 [b, Lpad, D]: out[i, l] = packed[row_offset[idx[i]] + l] for l < T_i, else 0
 *   out_off     NULL, or [b + 1] device int32: PACKED output for the varlen path - only the valid rows are written,
 *               utterance i at rows [out_off[i], out_off[i+1]) of out (no padding rows at all)
 * D % 8 == 0, 16-byte aligned pointers; utterances longer than Lpad are an error the caller rules out. */
int sdumc_collate_pad(const SDUMC_BF16* packed, const int64_t* row_offset, const int32_t* idx, int32_t b,
                      int32_t Lpad, int32_t D, SDUMC_BF16* out, const int32_t* out_off, void* stream);

/* ------------------------------------------------------------------------------------------
 * 4. Utterance-level glue between the MLP GEMMs (reference :301-320, :346-364) and the
 *    ReLU/dropout backward.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdumc_act_bwd_args {
  const float* dY; int64_t ld_dy;
  const float* dY2; int64_t ld_dy2; /* optional second addend (fan-in of two gradients) or NULL */
  const float* Y;  int64_t ld_y;   /* saved post-activation output or NULL */
  float scale;                     /* 1/(1-p) of the dropout after the ReLU */
  int32_t rows, cols;
  SDUMC_BF16* dZ; int64_t ld_dz;
  float* db;                       /* [cols] atomicAdd or NULL */
} sdumc_act_bwd_args;
int sdumc_act_bwd(const sdumc_act_bwd_args* a, void* stream);

typedef struct sdumc_gate_fwd_args {
  const float* a2; int64_t ld_a2;  /* attention_mlp output [R,256] */
  const float* Wg; const float* bg;/* fc_att [3,256], [3] */
  const float* h; int64_t ld_h;    /* [R,768] = (h_a|h_t|h_v) */
  int32_t R;
  int32_t G;                       /* general_dim (256 in the shapes quoted here; a multiple of 256 up to 1024) */
  float* g;                        /* [R,4] (3 used): raw gate, not softmaxed (:303-304) */
  float* qin; int64_t qin_stride;  /* 4 x [R,256]: fused, audio+text, text+video, audio+video */
} sdumc_gate_fwd_args;
int sdumc_gate_fwd(const sdumc_gate_fwd_args* a, void* stream);

typedef struct sdumc_gate_bwd_args {
  const float* dqin; int64_t dqin_stride;
  const float* dg_extra;           /* [R,4] gradient of g from the cross weighting (:346-349) */
  const float* g;
  const float* h; int64_t ld_h;
  const float* a2; int64_t ld_a2;
  const float* Wg;
  int32_t R;
  int32_t G;                       /* general_dim */
  float* dh; int64_t ld_dh;        /* [R,768] += */
  float* da2; int64_t ld_da2;      /* [R,256] store */
  float* dWg; float* dbg;          /* atomicAdd */
} sdumc_gate_bwd_args;
int sdumc_gate_bwd(const sdumc_gate_bwd_args* a, void* stream);

typedef struct sdumc_weight_fwd_args {
  const float* c[3];               /* each [R,7,128] */
  const float* g;                  /* [R,4] */
  int32_t R;
  float* W;                        /* [R,7,128] */
} sdumc_weight_fwd_args;
int sdumc_weight_fwd(const sdumc_weight_fwd_args* a, void* stream);

typedef struct sdumc_weight_bwd_args {
  const float* dW;
  const float* c[3];
  const float* g;
  int32_t R;
  const float* dc_extra[3];        /* optional external gradient added to dc[m] (e.g. the cross_text output) */
  float* dc[3];                    /* store */
  float* dg;                       /* [R,4] store */
} sdumc_weight_bwd_args;
int sdumc_weight_bwd(const sdumc_weight_bwd_args* a, void* stream);

typedef struct sdumc_final_fwd_args {
  const float* x2; int64_t ld_x2;  /* [R,128] */
  const float* Wr; const float* br;/* cross_fc_att [7,128],[7] */
  const float* W;                  /* [R,7,128] */
  const float* Wv; const float* bv;/* fc_out_v [1,128],[1] */
  int32_t R;
  float* r;                        /* [R,8] (7 used) */
  float* f;                        /* [R,128] */
  float* vals;                     /* [R] */
} sdumc_final_fwd_args;
int sdumc_final_fwd(const sdumc_final_fwd_args* a, void* stream);

typedef struct sdumc_final_bwd_args {
  const float* dvals;              /* [R] or NULL */
  const float* df_ext;             /* [R,128] or NULL */
  const float* x2; int64_t ld_x2;
  const float* Wr;
  const float* W;
  const float* r;
  const float* f;
  const float* Wv;
  int32_t R;
  float* dWc;                      /* [R,7,128] store */
  float* dx2; int64_t ld_dx2;      /* [R,128] store */
  float* dWr; float* dbr; float* dWv; float* dbv; /* atomicAdd */
} sdumc_final_bwd_args;
int sdumc_final_bwd(const sdumc_final_bwd_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * 5. Losses (toolkit/utils/loss.py:19-51, :243-315; combination main...:137-148)
 * ------------------------------------------------------------------------------------------ */
typedef struct sdumc_loss_sums_args {
  const float* v0; const float* v1; const float* y;  /* [B] */
  const float* th0; const float* th1;                /* [B,G] (G = 256 in the reference) */
  const float* ct0; const float* ct1;                /* [B,896] */
  const float* f0; const float* f1;                  /* [B,128] */
  int32_t B;
  int32_t G;                                         /* general_dim: width of th0 / th1; 0 = 256 */
  float* sums;                                       /* [8], atomicAdd of 5 sums of squares */
} sdumc_loss_sums_args;
int sdumc_loss_sums(const sdumc_loss_sums_args* a, void* stream);

typedef struct sdumc_loss_finish_args {
  sdumc_loss_sums_args in;
  const float* sums;   /* [8] reduced over the whole (global) batch */
  const float* rnc;    /* [1] global RnC value or NULL */
  int32_t B_global;
  float w[6];          /* full_mse, missing_mse, text_feat, text_query_feat, features, rnc */
  float* terms;        /* [8]: 6 terms, total, 0 */
  float* d_v0; float* d_v1;
  float* d_th1;        /* teacher side detached (:148) */
  float* d_ct1;        /* teacher side detached */
  float* d_f0; float* d_f1;
} sdumc_loss_finish_args;
int sdumc_loss_finish(const sdumc_loss_finish_args* a, void* stream);

int sdumc_sqdiff_sum(const float* a, const float* b, int64_t n, float* out_sum, void* stream);
int sdumc_sqdiff_grad(const float* a, const float* b, int64_t n, const float* coef_dev, float* da,
                      float* db_or_null, void* stream);

typedef struct sdumc_rnc_args {
  const float* feats;  /* [n,D], rows = (view-0 samples..., view-1 samples...) */
  const float* labels; /* [n] */
  int32_t n, D;
  int32_t row_begin, row_end; /* anchors of this call (a data-parallel rank passes its slice) */
  float temperature;
  float* loss;         /* [1] atomicAdd */
  float* dfeats;       /* [n,D] atomicAdd, or NULL for forward only */
  float grad_scale;
  void* workspace;
  uint64_t workspace_bytes;
  int32_t reuse_sort;  /* 1: the workspace still holds the label sort of a previous call with the same labels */
  int32_t phase;       /* 0: everything.  1: only the part that depends on the labels (sort, bucket index, the four
                          boundaries of every (anchor, element) pair) - a trainer runs it early, off the critical path;
                          feats / loss / dfeats may be NULL.  2: only the feature-dependent part (distances, loss,
                          gradient), after a phase-1 call with the same labels, row range and workspace.
                          3: sort + bucket index only; a phase-1 call with reuse_sort = 1 then adds the boundaries. */
} sdumc_rnc_args;
uint64_t sdumc_rnc_workspace_bytes(int32_t n, int32_t D);            /* all n anchor rows in one call */
uint64_t sdumc_rnc_workspace_bytes_rows(int32_t n, int32_t D, int32_t rows); /* calls with at most `rows` anchors */
int sdumc_rnc(const sdumc_rnc_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * 6. Optimiser: torch.optim.Adam semantics (main...:317), fused with the bf16 shadow refresh
 * ------------------------------------------------------------------------------------------ */
typedef struct sdumc_adam_args {
  float* p; const float* g; float* m; float* v;
  SDUMC_BF16* p_bf16; /* optional */
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, grad_scale;
  int32_t step; /* 1-based */
  const int32_t* step_dev; /* optional: device-resident step (overrides `step`), for CUDA-graph replay */
  const float* lr_dev;     /* optional: device-resident learning rate (overrides `lr`) */
} sdumc_adam_args;
int sdumc_adam(const sdumc_adam_args* a, void* stream);

/* sizeof() of every argument block, for binding self-checks: index = order of declaration above
 * (0 gemm_desc, 1 pool_fwd, 2 attn_bwd, 3 act_bwd, 4 gate_fwd, 5 gate_bwd, 6 weight_fwd, 7 weight_bwd,
 *  8 final_fwd, 9 final_bwd, 10 loss_sums, 11 loss_finish, 12 rnc, 13 adam) */
int sdumc_struct_size(int which);

#ifdef __cplusplus
}
#endif
#endif /* SDUMC_B200_H_ */
