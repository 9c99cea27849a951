// sdumc_b200 — frame-level kernels of the pooling-attention family:
//   softmax over the L frames of a sample + weighted pooling (forward), and the row-wise part of
//   the backward pass.  The dense parts (key projection, dZ*W_in, dZ^T*X') are tcgen05 GEMMs
//   (gemm.cuh); these kernels are HBM-bound streaming kernels: one CTA per sample, one warp per
//   frame row, 16-byte loads (8 bf16 columns per lane).
//
// Reference: FRA2UTT_new.forward / Cross_Attention.forward,
//   toolkit/models/wengnet_mosei_mult_views_text_missing.py:56-68, :79-95 (forward);
//   the backward formulas are autograd of those lines.
#include "common.cuh"
#include "kernels.h"

namespace sdumc {

static constexpr int G = 256;  // general_dim of the model (reference :191)

__device__ __forceinline__ void unpack8(const uint4& v, float (&x)[8]) {
  x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xffff0000u);
  x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xffff0000u);
  x[4] = __uint_as_float(v.z << 16); x[5] = __uint_as_float(v.z & 0xffff0000u);
  x[6] = __uint_as_float(v.w << 16); x[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float (&x)[8]) {
  return make_uint4(pack2(x[0], x[1]), pack2(x[2], x[3]), pack2(x[4], x[5]), pack2(x[6], x[7]));
}

// ------------------------------------------------------------------------------------------
// forward: P = softmax_L(alpha * S);  O = P^T X';  out = dropout(O)
// ------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(256) pool_fwd_kernel(PoolFwdArgs a) {
  extern __shared__ float sm[];
  float* Ps = sm;                 // [L][NQ]
  float* red = sm + a.L * NQ;     // [NQ][G]
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L;
  const DropKey key = resolve_key(a.key);
  float* Sg = a.S + (long)b * L * NQ;

  for (int i = tid; i < L * NQ; i += 256) Ps[i] = Sg[i];
  for (int i = tid; i < NQ * G; i += 256) red[i] = 0.f;
  __syncthreads();
  for (int q = warp; q < NQ; q += 8) {
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) m = fmaxf(m, Ps[l * NQ + q]);
    m = warp_max(m);
    float s = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float e = __expf(a.alpha * (Ps[l * NQ + q] - m));
      Ps[l * NQ + q] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int l = lane; l < L; l += 32) Ps[l * NQ + q] *= inv;
  }
  __syncthreads();
  for (int i = tid; i < L * NQ; i += 256) Sg[i] = Ps[i];

  float acc[NQ][8];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
  const __nv_bfloat16* Xb = a.X + (long)b * L * a.ldx + lane * 8;
  int l = warp;
  for (; l + 24 < L; l += 32) {  // 4 rows in flight per warp
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(Xb + (long)(l + 8 * u) * a.ldx));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float x[8];
      unpack8(v[u], x);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float p = Ps[(l + 8 * u) * NQ + q];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(p, x[j], acc[q][j]);
      }
    }
  }
  for (; l < L; l += 8) {
    float x[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(Xb + (long)l * a.ldx)), x);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float p = Ps[l * NQ + q];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(p, x[j], acc[q][j]);
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red[q * G + lane * 8 + j], acc[q][j]);
  __syncthreads();

  const uint32_t thr = drop_threshold(a.drop_p);
  const float scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  for (int i = tid; i < NQ * G; i += 256) {
    const float o = red[i];
    a.O_pre[(long)b * NQ * G + i] = o;
    float y = o;
    if (a.drop_p > 0.f) {
      const uint32_t e = (uint32_t)b * (uint32_t)(NQ * G) + (uint32_t)i;
      y = elem_rand(key, a.site, e) >= thr ? o * scale : 0.f;
    }
    a.out[(long)b * a.out_stride_b + i] = y;
    if (a.out_bf16) a.out_bf16[(long)b * a.out_stride_b + i] = __float2bfloat16_rn(y);
  }
}

int launch_pool_fwd(const PoolFwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.X && a.S && a.O_pre && a.out, "pool_fwd: null pointer");
  SDUMC_CHECK_ARG(a.B > 0 && a.L > 0 && (a.nq == 1 || a.nq == 7), "pool_fwd: bad shape B=%d L=%d nq=%d", a.B, a.L, a.nq);
  SDUMC_CHECK_ARG(a.ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(a.X) & 15u) == 0, "pool_fwd: X must be 16-byte aligned");
  const size_t smem = (size_t)(a.L * a.nq + a.nq * G) * sizeof(float);
  SDUMC_CHECK_ARG(smem <= 200 * 1024, "pool_fwd: L=%d too long for the shared-memory softmax", a.L);
  if (a.nq == 1) {
    static bool done1 = false;
    if (!done1) { SDUMC_CUDA(cudaFuncSetAttribute(pool_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done1 = true; }
    pool_fwd_kernel<1><<<a.B, 256, smem, stream>>>(a);
  } else {
    static bool done7 = false;
    if (!done7) { SDUMC_CUDA(cudaFuncSetAttribute(pool_fwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done7 = true; }
    pool_fwd_kernel<7><<<a.B, 256, smem, stream>>>(a);
  }
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// backward, row-wise part.  Given dOut (gradient of the dropped pooled output):
//   dO = dOut * M_out;  delta_q = <O_q, dO_q>
//   dP_lq = <X'_l, dO_q>;  dS_lq = alpha * P_lq * (dP_lq - delta_q)
//   dZ_l = (sum_q dS_lq Qp_q) * (1 - K_l^2)          -> bf16, feeds the two tcgen05 GEMMs
//   dQp_q = sum_l dS_lq K_l;  db_in = sum_l dZ_l
//   dH_l (+)= (sum_q P_lq dO_q) * M_in                (value path; the GEMM adds dZ W_in)
// ------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(256, 2) attn_bwd_kernel(AttnBwdArgs a) {
  __shared__ float dO_s[NQ][G];
  __shared__ float Qp_s[NQ][G];
  __shared__ float red_q[NQ][G];
  __shared__ float red_b[G];
  __shared__ float delta_s[NQ];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L;
  const uint32_t thr = drop_threshold(a.out_drop_p);
  const float oscale = a.out_drop_p > 0.f ? 1.f / (1.f - a.out_drop_p) : 1.f;
  const DropKey key = resolve_key(a.key);

  for (int i = tid; i < NQ * G; i += 256) {
    float g = a.dOut[(long)b * a.dout_stride_b + i];
    if (a.out_drop_p > 0.f) {
      const uint32_t e = (uint32_t)b * (uint32_t)(NQ * G) + (uint32_t)i;
      g = elem_rand(key, a.out_site, e) >= thr ? g * oscale : 0.f;
    }
    (&dO_s[0][0])[i] = g;
    (&Qp_s[0][0])[i] = a.Qp[(long)b * a.qp_stride_b + i];
    (&red_q[0][0])[i] = 0.f;
  }
  if (tid < G) red_b[tid] = 0.f;
  __syncthreads();
  for (int q = warp; q < NQ; q += 8) {
    float s = 0.f;
    for (int g = lane; g < G; g += 32) s = fmaf(a.O_pre[(long)b * NQ * G + q * G + g], dO_s[q][g], s);
    s = warp_sum(s);
    if (lane == 0) delta_s[q] = s;
  }
  __syncthreads();

  const int g0 = lane * 8;
  float dq_acc[NQ][8];
  float db_acc[8];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) dq_acc[q][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) db_acc[j] = 0.f;

  // software pipeline: the next row's X', K (and dH when accumulating) are in flight while this row computes
  uint4 nx = make_uint4(0, 0, 0, 0), nk = nx, nh = nx;
  if (warp < L) {
    const long row0 = (long)b * L + warp;
    nx = __ldg(reinterpret_cast<const uint4*>(a.X + row0 * a.ldx + g0));
    nk = __ldg(reinterpret_cast<const uint4*>(a.Kt + row0 * a.ldk + g0));
    if (a.dh_mode == 1) nh = *reinterpret_cast<const uint4*>(a.dH + row0 * a.lddh + g0);
  }
  for (int l = warp; l < L; l += 8) {
    const long row = (long)b * L + l;
    float x[8], k[8];
    unpack8(nx, x);
    unpack8(nk, k);
    const uint4 oldh = nh;
    if (l + 8 < L) {
      const long rown = row + 8;
      nx = __ldg(reinterpret_cast<const uint4*>(a.X + rown * a.ldx + g0));
      nk = __ldg(reinterpret_cast<const uint4*>(a.Kt + rown * a.ldk + g0));
      if (a.dh_mode == 1) nh = *reinterpret_cast<const uint4*>(a.dH + rown * a.lddh + g0);
    }
    float dP[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4 d0 = *reinterpret_cast<const float4*>(&dO_s[q][g0]);
      const float4 d1 = *reinterpret_cast<const float4*>(&dO_s[q][g0 + 4]);
      float s = x[0] * d0.x;
      s = fmaf(x[1], d0.y, s); s = fmaf(x[2], d0.z, s); s = fmaf(x[3], d0.w, s);
      s = fmaf(x[4], d1.x, s); s = fmaf(x[5], d1.y, s); s = fmaf(x[6], d1.z, s); s = fmaf(x[7], d1.w, s);
      dP[q] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int q = 0; q < NQ; ++q) dP[q] += __shfl_xor_sync(0xffffffffu, dP[q], o);
    float dS[NQ], Pv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      Pv[q] = __ldg(a.P + row * NQ + q);
      dS[q] = a.alpha * Pv[q] * (dP[q] - delta_s[q]);
    }
    float dK[8], dXv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { dK[j] = 0.f; dXv[j] = 0.f; }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4 q0 = *reinterpret_cast<const float4*>(&Qp_s[q][g0]);
      const float4 q1 = *reinterpret_cast<const float4*>(&Qp_s[q][g0 + 4]);
      const float4 d0 = *reinterpret_cast<const float4*>(&dO_s[q][g0]);
      const float4 d1 = *reinterpret_cast<const float4*>(&dO_s[q][g0 + 4]);
      const float qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dK[j] = fmaf(dS[q], qq[j], dK[j]);
        dXv[j] = fmaf(Pv[q], dd[j], dXv[j]);
        dq_acc[q][j] = fmaf(dS[q], k[j], dq_acc[q][j]);
      }
    }
    float dZ[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dZ[j] = dK[j] * (1.f - k[j] * k[j]);
      db_acc[j] += dZ[j];
    }
    *reinterpret_cast<uint4*>(a.dZ + row * a.lddz + g0) = pack8(dZ);

    if (a.fmask_site) {
      const U4 w = frame_mask_words(key, a.fmask_site, (uint32_t)row, (uint32_t)(g0 >> 7));
      const int wsel = (g0 >> 5) & 3;
      const uint32_t bits = (wsel == 0 ? w.x : (wsel == 1 ? w.y : (wsel == 2 ? w.z : w.w))) >> (g0 & 31);
#pragma unroll
      for (int j = 0; j < 8; ++j) dXv[j] = ((bits >> j) & 1u) ? 2.f * dXv[j] : 0.f;
    }
    uint4* hp = reinterpret_cast<uint4*>(a.dH + row * a.lddh + g0);
    if (a.dh_mode == 1) {
      float old[8];
      unpack8(oldh, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) dXv[j] += old[j];
    }
    *hp = pack8(dXv);
  }

#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red_q[q][g0 + j], dq_acc[q][j]);
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&red_b[g0 + j], db_acc[j]);
  __syncthreads();
  for (int i = tid; i < NQ * G; i += 256) {
    const float v = (&red_q[0][0])[i];
    if (a.qp_stride_b == 0) atomicAdd(a.dQp + i, v);            // shared context vector: sum over the batch
    else a.dQp[(long)b * a.dqp_stride_b + i] = v;
  }
  if (tid < G) atomicAdd(a.db + tid, red_b[tid]);
}

int launch_attn_bwd(const AttnBwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.X && a.Kt && a.P && a.dOut && a.O_pre && a.Qp && a.dZ && a.dH && a.dQp && a.db,
                  "attn_bwd: null pointer");
  SDUMC_CHECK_ARG(a.B > 0 && a.L > 0 && (a.nq == 1 || a.nq == 7), "attn_bwd: bad shape");
  SDUMC_CHECK_ARG(a.ldx % 8 == 0 && a.ldk % 8 == 0 && a.lddz % 8 == 0 && a.lddh % 8 == 0, "attn_bwd: ld %% 8");
  if (a.nq == 1) attn_bwd_kernel<1><<<a.B, 256, 0, stream>>>(a);
  else           attn_bwd_kernel<7><<<a.B, 256, 0, stream>>>(a);
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// small streaming helpers
// ------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
    *reinterpret_cast<uint4*>(dst + i) =
        make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
  } else {
    for (long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}
int launch_cast_bf16(const float* src, __nv_bfloat16* dst, long n, cudaStream_t stream) {
  SDUMC_CHECK_ARG(src && dst && n > 0, "cast_bf16: bad arguments");
  SDUMC_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0,
                  "cast_bf16: pointers must be 16-byte aligned");
  const long nthreads = (n + 7) / 8;
  cast_f32_bf16_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, stream>>>(src, dst, n);
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// column sums of a bf16 matrix [rows, 256] -> atomicAdd into out[256]   (bias gradient of the in-projection)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long ld, long rows,
                                                           float* __restrict__ out) {
  __shared__ float red[G];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < G) red[tid] = 0.f;
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long r = (long)blockIdx.x * 8 + warp; r < rows; r += (long)gridDim.x * 8) {
    float x[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(X + r * ld + lane * 8)), x);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += x[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&red[lane * 8 + j], acc[j]);
  __syncthreads();
  if (tid < G) atomicAdd(out + tid, red[tid]);
}
int launch_colsum_bf16(const __nv_bfloat16* X, long ld, long rows, float* out, cudaStream_t stream) {
  SDUMC_CHECK_ARG(X && out && rows > 0 && ld % 8 == 0, "colsum_bf16: bad arguments");
  long blocks = (rows + 63) / 64;
  if (blocks > 148 * 8) blocks = 148 * 8;
  colsum_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(X, ld, rows, out);
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sdumc
