// sdumc_b200 — fused Adam over the flat parameter buffer (torch.optim.Adam semantics: L2 weight
// decay folded into the gradient, bias-corrected moments), writing the bf16 shadow copy the
// tensor-core GEMMs read in the same pass.  Reference: main_frame_val_text_missing.py:317,:150.
// HBM-bound: 16 B/param read (p, g, m, v) + 12 B/param written (+2 B shadow).
#include "common.cuh"
#include "kernels.h"

namespace sdumc {

__global__ void __launch_bounds__(256) adam_kernel(AdamArgs a, float step_size, float inv_bc2_sqrt) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const long i4 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= a.n) return;
  const float b1 = a.beta1, b2 = a.beta2;
  if (a.step_dev || a.lr_dev) {  // device-resident step / learning rate (CUDA-graph replay)
    const double t = (double)(a.step_dev ? __ldg(a.step_dev) : a.step);
    const double lr = (double)(a.lr_dev ? __ldg(a.lr_dev) : a.lr);
    step_size = (float)(lr / (1.0 - pow((double)b1, t)));
    inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
  }
  if (i4 + 4 <= a.n) {
    float4 p = *reinterpret_cast<float4*>(a.p + i4);
    const float4 g = *reinterpret_cast<const float4*>(a.g + i4);
    float4 m = *reinterpret_cast<float4*>(a.m + i4);
    float4 v = *reinterpret_cast<float4*>(a.v + i4);
    float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g.x, g.y, g.z, g.w}, mm[4] = {m.x, m.y, m.z, m.w},
          vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * a.grad_scale + a.weight_decay * pp[j];
      mm[j] = b1 * mm[j] + (1.f - b1) * gr;
      vv[j] = b2 * vv[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vv[j]) * inv_bc2_sqrt + a.eps;
      pp[j] -= step_size * mm[j] / denom;
    }
    *reinterpret_cast<float4*>(a.p + i4) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(a.m + i4) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(a.v + i4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (a.p_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(a.p_bf16 + i4) = pk;
    }
  } else {
    for (long i = i4; i < a.n; ++i) {
      const float gr = a.g[i] * a.grad_scale + a.weight_decay * a.p[i];
      const float m = b1 * a.m[i] + (1.f - b1) * gr;
      const float v = b2 * a.v[i] + (1.f - b2) * gr * gr;
      a.m[i] = m;
      a.v[i] = v;
      const float p = a.p[i] - step_size * m / (sqrtf(v) * inv_bc2_sqrt + a.eps);
      a.p[i] = p;
      if (a.p_bf16) a.p_bf16[i] = __float2bfloat16_rn(p);
    }
  }
}

int launch_adam(const AdamArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.p && a.g && a.m && a.v && a.n > 0 && (a.step >= 1 || a.step_dev), "adam: bad arguments");
  SDUMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.p) | reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.m) |
                    reinterpret_cast<uintptr_t>(a.v)) & 15u) == 0,
                  "adam: buffers must be 16-byte aligned");
  const int st = a.step >= 1 ? a.step : 1;
  const double bc1 = 1.0 - pow((double)a.beta1, (double)st);
  const double bc2 = 1.0 - pow((double)a.beta2, (double)st);
  const float step_size = (float)((double)a.lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const long nthreads = (a.n + 3) / 4;
  SDUMC_CUDA(launch_kernel(adam_kernel, dim3((unsigned)((nthreads + 255) / 256)), dim3(256), 0, stream, 1, a, step_size, inv_bc2_sqrt));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sdumc
