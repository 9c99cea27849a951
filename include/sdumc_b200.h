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
  uint32_t dbg_lbo, dbg_sbo; /* test-only descriptor overrides, 0 = default */
} sdumc_gemm_desc;

int sdumc_gemm(const sdumc_gemm_desc* d, void* stream);

/* Test hooks: materialise the dropout masks the kernels apply (values 0 or 1/(1-p)). */
int sdumc_frame_mask(uint64_t seed, uint32_t step, uint32_t site, int64_t rows, int32_t cols, float* out,
                     void* stream);
int sdumc_elem_mask(uint64_t seed, uint32_t step, uint32_t site, int64_t n, float p, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDUMC_B200_H_ */
