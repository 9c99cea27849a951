// sdumc_b200 — internal names of the C-ABI argument blocks (include/sdumc_b200.h) and the host
// launchers of the non-GEMM kernels.
#pragma once

#include "common.cuh"

namespace sdumc {

using PoolFwdArgs = sdumc_pool_fwd_args;
using AttnBwdArgs = sdumc_attn_bwd_args;
using ActBwdArgs = sdumc_act_bwd_args;
using GateFwdArgs = sdumc_gate_fwd_args;
using GateBwdArgs = sdumc_gate_bwd_args;
using WeightFwdArgs = sdumc_weight_fwd_args;
using WeightBwdArgs = sdumc_weight_bwd_args;
using FinalFwdArgs = sdumc_final_fwd_args;
using FinalBwdArgs = sdumc_final_bwd_args;
using LossSumsArgs = sdumc_loss_sums_args;
using LossFinishArgs = sdumc_loss_finish_args;
using RncArgs = sdumc_rnc_args;
using AdamArgs = sdumc_adam_args;

// frame.cu
int launch_pool_fwd(const PoolFwdArgs& a, cudaStream_t stream);
int launch_attn_bwd(const AttnBwdArgs& a, cudaStream_t stream);
int launch_cast_bf16(const float* src, __nv_bfloat16* dst, long n, cudaStream_t stream);
int launch_colsum_bf16(const __nv_bfloat16* X, long ld, long rows, int cols, float* out, cudaStream_t stream);
int launch_collate_pad(const __nv_bfloat16* packed, const long long* row_offset, const int* idx, int b, int Lpad, int D,
                       __nv_bfloat16* out, const int* out_off, cudaStream_t stream);
// chain.cu
int launch_act_bwd(const ActBwdArgs& a, cudaStream_t stream);
int launch_gate_fwd(const GateFwdArgs& a, cudaStream_t stream);
int launch_gate_bwd(const GateBwdArgs& a, cudaStream_t stream);
int launch_weight_fwd(const WeightFwdArgs& a, cudaStream_t stream);
int launch_weight_bwd(const WeightBwdArgs& a, cudaStream_t stream);
int launch_final_fwd(const FinalFwdArgs& a, cudaStream_t stream);
int launch_final_bwd(const FinalBwdArgs& a, cudaStream_t stream);
// loss.cu
int launch_loss_sums(const LossSumsArgs& a, cudaStream_t stream);
int launch_loss_finish(const LossFinishArgs& a, cudaStream_t stream);
int launch_sqdiff_sum(const float* a, const float* b, long n, float* out_sum, cudaStream_t stream);
int launch_sqdiff_grad(const float* a, const float* b, long n, const float* coef, float* da, float* db_or_null,
                       cudaStream_t stream);
size_t rnc_workspace_bytes(int n, int D, int rows);
int launch_rnc(const RncArgs& a, cudaStream_t stream);
// adam.cu
int launch_adam(const AdamArgs& a, cudaStream_t stream);

}  // namespace sdumc
