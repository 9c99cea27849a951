// sdumc_b200 — extern "C" entry points (include/sdumc_b200.h).
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"

namespace sdumc {
const char* last_error();

__global__ void frame_mask_kernel(DropKey key, uint32_t site, long rows, int cols, float* out) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, 32-col chunk)
  const int chunks = (cols + 31) / 32;
  if (idx >= rows * chunks) return;
  const long r = idx / chunks;
  const int c = (int)(idx - r * chunks);
  const int n0 = c * 32;
  const U4 w = frame_mask_words(key, site, (uint32_t)r, (uint32_t)(n0 >> 7));
  const int wsel = (n0 >> 5) & 3;
  const uint32_t bits = wsel == 0 ? w.x : (wsel == 1 ? w.y : (wsel == 2 ? w.z : w.w));
  for (int j = 0; j < 32 && n0 + j < cols; ++j) out[r * cols + n0 + j] = ((bits >> j) & 1u) ? 2.f : 0.f;
}

__global__ void elem_mask_kernel(DropKey key, uint32_t site, long n, uint32_t thr, float scale, float* out) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  out[e] = elem_rand(key, site, (uint32_t)e) >= thr ? scale : 0.f;
}
}  // namespace sdumc

using namespace sdumc;

extern "C" {

int sdumc_version(void) { return SDUMC_ABI_VERSION; }
const char* sdumc_last_error(void) { return sdumc::last_error(); }

int sdumc_gemm(const sdumc_gemm_desc* d, void* stream) {
  SDUMC_CHECK_ARG(d != nullptr, "sdumc_gemm: null descriptor");
  GemmShape sh{};
  sh.M = d->M; sh.N = d->N; sh.K = d->K;
  sh.a_mn = d->a_mn; sh.b_mn = d->b_mn;
  sh.k_splits = d->k_splits;
  GemmEpi ep{};
  ep.kind = d->epi_kind;
  ep.bias = d->bias;
  ep.act = d->act;
  ep.gate = d->gate; ep.ld_gate = d->ld_gate; ep.gate_scale = d->gate_scale;
  ep.drop_p = d->drop_p; ep.drop_site = d->drop_site;
  ep.fmask_site = d->fmask_site; ep.fmask_site2 = d->fmask_site2; ep.fmask_split = d->fmask_split;
  SDUMC_CHECK_ARG(d->fmask_split == 0 || (d->fmask_site && d->fmask_site2 && d->fmask_split > 0 && d->out_bf16 &&
                                          d->bf16_mode == 1 && d->epi_kind == 0),
                  "sdumc_gemm: fmask_split is defined for the bf16 '+=' output with a frame mask (dH accumulation)");
  ep.out_f32 = d->out_f32; ep.ld_f32 = d->ld_f32; ep.f32_mode = d->f32_mode;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16); ep.ld_bf16 = d->ld_bf16; ep.bf16_mode = d->bf16_mode;
  ep.n_tgt = d->n_tgt;
  for (int i = 0; i < 4; ++i) {
    ep.tgt[i] = static_cast<__nv_bfloat16*>(d->tgt[i]);
    ep.tgt_site[i] = d->tgt_site[i];
  }
  ep.qv = d->qv; ep.q_stride = d->q_stride; ep.nq = d->nq; ep.L = d->L; ep.scores = d->scores;
  ep.key = make_dropkey(d->seed, d->step, d->step_dev);
  GemmOperand A{d->A, d->lda}, B{d->B, d->ldb};
  return launch_gemm(A, B, sh, ep, d->tf32 != 0, d->block_n, d->max_ctas, static_cast<cudaStream_t>(stream));
}

int sdumc_frame_mask(uint64_t seed, uint32_t step, uint32_t site, int64_t rows, int32_t cols, float* out,
                     void* stream) {
  SDUMC_CHECK_ARG(out && rows > 0 && cols > 0, "sdumc_frame_mask: bad arguments");
  DropKey key = make_dropkey(seed, step);
  const long n = rows * ((cols + 31) / 32);
  SDUMC_CUDA(launch_kernel(frame_mask_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, key, site, rows, cols,
                                                                                                 out));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

int sdumc_elem_mask(uint64_t seed, uint32_t step, uint32_t site, int64_t n, float p, float* out, void* stream) {
  SDUMC_CHECK_ARG(out && n > 0 && p >= 0.f && p < 1.f, "sdumc_elem_mask: bad arguments");
  DropKey key = make_dropkey(seed, step);
  SDUMC_CUDA(launch_kernel(elem_mask_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, key, site, n, drop_threshold(p), 1.f / (1.f - p), out));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

#define SDUMC_FWD(name, type, launcher)                                        \
  int name(const type* a, void* stream) {                                     \
    SDUMC_CHECK_ARG(a != nullptr, #name ": null argument block");             \
    return launcher(*a, static_cast<cudaStream_t>(stream));                   \
  }
SDUMC_FWD(sdumc_pool_fwd, sdumc_pool_fwd_args, launch_pool_fwd)
SDUMC_FWD(sdumc_attn_bwd, sdumc_attn_bwd_args, launch_attn_bwd)
SDUMC_FWD(sdumc_act_bwd, sdumc_act_bwd_args, launch_act_bwd)
SDUMC_FWD(sdumc_gate_fwd, sdumc_gate_fwd_args, launch_gate_fwd)
SDUMC_FWD(sdumc_gate_bwd, sdumc_gate_bwd_args, launch_gate_bwd)
SDUMC_FWD(sdumc_weight_fwd, sdumc_weight_fwd_args, launch_weight_fwd)
SDUMC_FWD(sdumc_weight_bwd, sdumc_weight_bwd_args, launch_weight_bwd)
SDUMC_FWD(sdumc_final_fwd, sdumc_final_fwd_args, launch_final_fwd)
SDUMC_FWD(sdumc_final_bwd, sdumc_final_bwd_args, launch_final_bwd)
SDUMC_FWD(sdumc_loss_sums, sdumc_loss_sums_args, launch_loss_sums)
SDUMC_FWD(sdumc_loss_finish, sdumc_loss_finish_args, launch_loss_finish)
SDUMC_FWD(sdumc_rnc, sdumc_rnc_args, launch_rnc)
SDUMC_FWD(sdumc_adam, sdumc_adam_args, launch_adam)
#undef SDUMC_FWD

int sdumc_cast_bf16(const float* src, SDUMC_BF16* dst, int64_t n, void* stream) {
  return launch_cast_bf16(src, dst, n, static_cast<cudaStream_t>(stream));
}
int sdumc_colsum_bf16(const SDUMC_BF16* X, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream) {
  return launch_colsum_bf16(X, ld, rows, cols, out, static_cast<cudaStream_t>(stream));
}
int sdumc_collate_pad(const SDUMC_BF16* packed, const int64_t* row_offset, const int32_t* idx, int32_t b,
                      int32_t Lpad, int32_t D, SDUMC_BF16* out, const int32_t* out_off, void* stream) {
  return launch_collate_pad(packed, reinterpret_cast<const long long*>(row_offset), idx, b, Lpad, D, out, out_off,
                            static_cast<cudaStream_t>(stream));
}
int sdumc_sqdiff_sum(const float* a, const float* b, int64_t n, float* out_sum, void* stream) {
  return launch_sqdiff_sum(a, b, n, out_sum, static_cast<cudaStream_t>(stream));
}
int sdumc_sqdiff_grad(const float* a, const float* b, int64_t n, const float* coef_dev, float* da, float* db_or_null,
                      void* stream) {
  return launch_sqdiff_grad(a, b, n, coef_dev, da, db_or_null, static_cast<cudaStream_t>(stream));
}
uint64_t sdumc_rnc_workspace_bytes(int32_t n, int32_t D) { return rnc_workspace_bytes(n, D, n); }
uint64_t sdumc_rnc_workspace_bytes_rows(int32_t n, int32_t D, int32_t rows) { return rnc_workspace_bytes(n, D, rows); }

int sdumc_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(sdumc_gemm_desc);
    case 1: return (int)sizeof(sdumc_pool_fwd_args);
    case 2: return (int)sizeof(sdumc_attn_bwd_args);
    case 3: return (int)sizeof(sdumc_act_bwd_args);
    case 4: return (int)sizeof(sdumc_gate_fwd_args);
    case 5: return (int)sizeof(sdumc_gate_bwd_args);
    case 6: return (int)sizeof(sdumc_weight_fwd_args);
    case 7: return (int)sizeof(sdumc_weight_bwd_args);
    case 8: return (int)sizeof(sdumc_final_fwd_args);
    case 9: return (int)sizeof(sdumc_final_bwd_args);
    case 10: return (int)sizeof(sdumc_loss_sums_args);
    case 11: return (int)sizeof(sdumc_loss_finish_args);
    case 12: return (int)sizeof(sdumc_rnc_args);
    case 13: return (int)sizeof(sdumc_adam_args);
    default: return -1;
  }
}

}  // extern "C"
