// sdumc_b200 — utterance-level glue kernels between the tcgen05 GEMMs of the MLP chain:
// raw modality gate + partial fusions (reference model file :301-320), gate-weighted sum of the
// cross-attended features (:346-349), query gate + fused feature + regression head (:352-364),
// and the ReLU/dropout backward that turns dY into the bf16 dZ the weight-gradient GEMMs consume.
// Rows are utterances (R = passes x batch); everything here is a few MB: latency-, not
// bandwidth-bound, one warp per utterance row.
#include "common.cuh"
#include "kernels.h"

namespace sdumc {


// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_bwd_kernel(ActBwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  // 64 column-quads x 4 row lanes; the bias gradient is reduced inside the block before the atomics
  // (one atomic per column per block: same-address atomics serialise in L2).
  __shared__ float red[4][256];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int cb = 0; cb < a.cols; cb += 256) {
    const int c = cb + tx * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < a.cols) {
      for (int r = blockIdx.x * 4 + ty; r < a.rows; r += gridDim.x * 4) {
        const float4 dy = *reinterpret_cast<const float4*>(a.dY + (long)r * a.ld_dy + c);
        float z[4] = {dy.x, dy.y, dy.z, dy.w};
        if (a.dY2) {
          const float4 e = *reinterpret_cast<const float4*>(a.dY2 + (long)r * a.ld_dy2 + c);
          z[0] += e.x; z[1] += e.y; z[2] += e.z; z[3] += e.w;
        }
        if (a.Y) {
          const float4 y = *reinterpret_cast<const float4*>(a.Y + (long)r * a.ld_y + c);
          z[0] = y.x > 0.f ? z[0] * a.scale : 0.f;
          z[1] = y.y > 0.f ? z[1] * a.scale : 0.f;
          z[2] = y.z > 0.f ? z[2] * a.scale : 0.f;
          z[3] = y.w > 0.f ? z[3] * a.scale : 0.f;
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(z[0], z[1]), hi = __floats2bfloat162_rn(z[2], z[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(a.dZ + (long)r * a.ld_dz + c) = pk;
        acc[0] += z[0]; acc[1] += z[1]; acc[2] += z[2]; acc[3] += z[3];
      }
    }
    if (a.db) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = acc[j];
      __syncthreads();
      const int col = cb + threadIdx.x;
      if (col < a.cols)
        atomicAdd(a.db + col, (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]));
    }
  }
}
int launch_act_bwd(const ActBwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.dY && a.dZ && a.rows > 0 && a.cols > 0 && a.cols % 4 == 0, "act_bwd: bad arguments");
  SDUMC_CHECK_ARG(a.ld_dy % 4 == 0 && a.ld_dz % 4 == 0 && (!a.Y || a.ld_y % 4 == 0) && (!a.dY2 || a.ld_dy2 % 4 == 0),
                  "act_bwd: ld %% 4");
  int blocks = (a.rows + 63) / 64;
  if (blocks > 148) blocks = 148;
  SDUMC_CUDA(launch_kernel(act_bwd_kernel, dim3(blocks), dim3(256), 0, stream, 1, a));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// gate forward: warp per utterance row; a lane owns 8 columns of every 256-column block (G = 256: one block)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8(const float* p, float (&x)[8]) {
  const float4 u = *reinterpret_cast<const float4*>(p), v = *reinterpret_cast<const float4*>(p + 4);
  x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
}
__device__ __forceinline__ void st8(float* p, const float (&x)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}

__global__ void __launch_bounds__(256) gate_fwd_kernel(GateFwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= a.R) return;
  const int G = a.G;
  float s[3] = {0.f, 0.f, 0.f};
  for (int c0 = lane * 8; c0 < G; c0 += 256) {
    float x[8];
    ld8(a.a2 + (long)r * a.ld_a2 + c0, x);
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[m] = fmaf(x[j], __ldg(a.Wg + m * G + c0 + j), s[m]);
  }
  float g[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) g[m] = warp_sum(s[m]) + __ldg(a.bg + m);
  if (lane < 4) a.g[(long)r * 4 + lane] = lane == 0 ? g[0] : (lane == 1 ? g[1] : (lane == 2 ? g[2] : 0.f));
  for (int c0 = lane * 8; c0 < G; c0 += 256) {
    float h[3][8];
#pragma unroll
    for (int m = 0; m < 3; ++m) ld8(a.h + (long)r * a.ld_h + m * G + c0, h[m]);
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float pa = g[0] * h[0][j], pt = g[1] * h[1][j], pv = g[2] * h[2][j];
      // same association order as torch.matmul over the stacked modalities: (a + t) + v
      o[0][j] = (pa + pt) + pv;  // fused
      o[1][j] = pa + pt;         // audio+text
      o[2][j] = pt + pv;         // text+video
      o[3][j] = pa + pv;         // audio+video
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) st8(a.qin + (long)i * a.qin_stride + (long)r * G + c0, o[i]);
  }
}
int launch_gate_fwd(const GateFwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.a2 && a.Wg && a.bg && a.h && a.g && a.qin && a.R > 0, "gate_fwd: bad arguments");
  SDUMC_CHECK_ARG(a.G > 0 && a.G % 256 == 0 && a.G <= 1024, "gate: general_dim %d unsupported", a.G);
  SDUMC_CUDA(launch_kernel(gate_fwd_kernel, dim3((a.R + 7) / 8), dim3(256), 0, stream, 1, a));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) gate_bwd_kernel(GateBwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  __shared__ float sW[3 * 1024];
  __shared__ float sb[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.G;
  for (int i = threadIdx.x; i < 3 * G; i += 256) sW[i] = 0.f;
  if (threadIdx.x < 4) sb[threadIdx.x] = 0.f;
  __syncthreads();
  const int r = blockIdx.x * 8 + warp;
  if (r < a.R) {
    const float g0 = a.g[(long)r * 4 + 0], g1 = a.g[(long)r * 4 + 1], g2 = a.g[(long)r * 4 + 2];
    float dg[3] = {0.f, 0.f, 0.f};
    // pass 1: dh += g * (sum of the consumers' gradients), dg = <h, ...> over all G columns
    for (int c0 = lane * 8; c0 < G; c0 += 256) {
      float d[4][8], h[3][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) ld8(a.dqin + (long)i * a.dqin_stride + (long)r * G + c0, d[i]);
#pragma unroll
      for (int m = 0; m < 3; ++m) ld8(a.h + (long)r * a.ld_h + m * G + c0, h[m]);
      float dh[3][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ua = d[0][j] + d[1][j] + d[3][j];  // consumers of g_a h_a: fused, at, av
        const float ut = d[0][j] + d[1][j] + d[2][j];  // fused, at, tv
        const float uv = d[0][j] + d[2][j] + d[3][j];  // fused, tv, av
        dh[0][j] = g0 * ua; dh[1][j] = g1 * ut; dh[2][j] = g2 * uv;
        dg[0] = fmaf(h[0][j], ua, dg[0]);
        dg[1] = fmaf(h[1][j], ut, dg[1]);
        dg[2] = fmaf(h[2][j], uv, dg[2]);
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        float* dst = a.dh + (long)r * a.ld_dh + m * G + c0;
        float old[8];
        ld8(dst, old);
#pragma unroll
        for (int j = 0; j < 8; ++j) old[j] += dh[m][j];
        st8(dst, old);
      }
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      dg[m] = warp_sum(dg[m]);
      if (a.dg_extra) dg[m] += a.dg_extra[(long)r * 4 + m];
    }
    // pass 2: through fc_att
    for (int c0 = lane * 8; c0 < G; c0 += 256) {
      float x[8], da[8];
      ld8(a.a2 + (long)r * a.ld_a2 + c0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        da[j] = dg[0] * __ldg(a.Wg + c0 + j) + dg[1] * __ldg(a.Wg + G + c0 + j) + dg[2] * __ldg(a.Wg + 2 * G + c0 + j);
#pragma unroll
        for (int m = 0; m < 3; ++m) atomicAdd(&sW[m * G + c0 + j], dg[m] * x[j]);
      }
      st8(a.da2 + (long)r * a.ld_da2 + c0, da);
    }
    if (lane < 3) atomicAdd(&sb[lane], lane == 0 ? dg[0] : (lane == 1 ? dg[1] : dg[2]));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * G; i += 256) atomicAdd(a.dWg + i, sW[i]);
  if (threadIdx.x < 3) atomicAdd(a.dbg + threadIdx.x, sb[threadIdx.x]);
}
int launch_gate_bwd(const GateBwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.dqin && a.g && a.h && a.a2 && a.Wg && a.dh && a.da2 && a.dWg && a.dbg && a.R > 0,
                  "gate_bwd: bad arguments");
  SDUMC_CHECK_ARG(a.G > 0 && a.G % 256 == 0 && a.G <= 1024, "gate: general_dim %d unsupported", a.G);
  SDUMC_CUDA(launch_kernel(gate_bwd_kernel, dim3((a.R + 7) / 8), dim3(256), 0, stream, 1, a));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// W[r,q,c] = sum_m g[r,m] c_m[r,q,c]     (7*128 = 896 values per row)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_fwd_kernel(WeightFwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const long i4 = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i4 >= (long)a.R * 896) return;
  const long r = i4 / 896;
  const float g0 = a.g[r * 4], g1 = a.g[r * 4 + 1], g2 = a.g[r * 4 + 2];
  const float4 x = *reinterpret_cast<const float4*>(a.c[0] + i4);
  const float4 y = *reinterpret_cast<const float4*>(a.c[1] + i4);
  const float4 z = *reinterpret_cast<const float4*>(a.c[2] + i4);
  float4 o;
  o.x = (g0 * x.x + g1 * y.x) + g2 * z.x;
  o.y = (g0 * x.y + g1 * y.y) + g2 * z.y;
  o.z = (g0 * x.z + g1 * y.z) + g2 * z.z;
  o.w = (g0 * x.w + g1 * y.w) + g2 * z.w;
  *reinterpret_cast<float4*>(a.W + i4) = o;
}
int launch_weight_fwd(const WeightFwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.c[0] && a.c[1] && a.c[2] && a.g && a.W && a.R > 0, "weight_fwd: bad arguments");
  const long n4 = (long)a.R * 896 / 4;
  SDUMC_CUDA(launch_kernel(weight_fwd_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, stream, 1, a));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) weight_bwd_kernel(WeightBwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= a.R) return;
  const float g[3] = {a.g[(long)r * 4], a.g[(long)r * 4 + 1], a.g[(long)r * 4 + 2]};
  float dg[3] = {0.f, 0.f, 0.f};
  for (int i = lane * 4; i < 896; i += 128) {
    const long off = (long)r * 896 + i;
    const float4 d = *reinterpret_cast<const float4*>(a.dW + off);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const float4 c = *reinterpret_cast<const float4*>(a.c[m] + off);
      dg[m] += d.x * c.x + d.y * c.y + d.z * c.z + d.w * c.w;
      float4 o = make_float4(g[m] * d.x, g[m] * d.y, g[m] * d.z, g[m] * d.w);
      if (a.dc_extra[m]) {
        const float4 e = *reinterpret_cast<const float4*>(a.dc_extra[m] + off);
        o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
      }
      *reinterpret_cast<float4*>(a.dc[m] + off) = o;
    }
  }
#pragma unroll
  for (int m = 0; m < 3; ++m) dg[m] = warp_sum(dg[m]);
  if (lane < 4) a.dg[(long)r * 4 + lane] = lane == 0 ? dg[0] : (lane == 1 ? dg[1] : (lane == 2 ? dg[2] : 0.f));
}
int launch_weight_bwd(const WeightBwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.dW && a.c[0] && a.c[1] && a.c[2] && a.g && a.dc[0] && a.dc[1] && a.dc[2] && a.dg && a.R > 0,
                  "weight_bwd: bad arguments");
  SDUMC_CUDA(launch_kernel(weight_bwd_kernel, dim3((a.R + 7) / 8), dim3(256), 0, stream, 1, a));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// r = cross_fc_att(x2) [7]; f = sum_q r_q W_q [128]; vals = fc_out_v(f).  Warp per row, lane owns 4 columns.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) final_fwd_kernel(FinalFwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= a.R) return;
  const int c0 = lane * 4;
  const float4 x = *reinterpret_cast<const float4*>(a.x2 + (long)r * a.ld_x2 + c0);
  float rq[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(a.Wr + q * 128 + c0));
    rq[q] = warp_sum(x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w) + __ldg(a.br + q);
  }
  {
    float mine = 0.f;
#pragma unroll
    for (int q = 0; q < 7; ++q) mine = (lane == q) ? rq[q] : mine;
    if (lane < 8) a.r[(long)r * 8 + lane] = mine;
  }
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const float4 w = *reinterpret_cast<const float4*>(a.W + (long)r * 896 + q * 128 + c0);
    f.x = fmaf(rq[q], w.x, f.x); f.y = fmaf(rq[q], w.y, f.y); f.z = fmaf(rq[q], w.z, f.z); f.w = fmaf(rq[q], w.w, f.w);
  }
  *reinterpret_cast<float4*>(a.f + (long)r * 128 + c0) = f;
  const float4 wv = __ldg(reinterpret_cast<const float4*>(a.Wv + c0));
  const float v = warp_sum(f.x * wv.x + f.y * wv.y + f.z * wv.z + f.w * wv.w) + __ldg(a.bv);
  if (lane == 0) a.vals[r] = v;
}
int launch_final_fwd(const FinalFwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.x2 && a.Wr && a.br && a.W && a.Wv && a.bv && a.r && a.f && a.vals && a.R > 0,
                  "final_fwd: bad arguments");
  SDUMC_CUDA(launch_kernel(final_fwd_kernel, dim3((a.R + 7) / 8), dim3(256), 0, stream, 1, a));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) final_bwd_kernel(FinalBwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  __shared__ float sWr[7][128];
  __shared__ float sWv[128];
  __shared__ float sbr[8];
  __shared__ float sbv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 7 * 128; i += 256) (&sWr[0][0])[i] = 0.f;
  if (threadIdx.x < 128) sWv[threadIdx.x] = 0.f;
  if (threadIdx.x < 8) sbr[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) sbv = 0.f;
  __syncthreads();
  const int r = blockIdx.x * 8 + warp;
  if (r < a.R) {
    const int c0 = lane * 4;
    const float dv = a.dvals ? a.dvals[r] : 0.f;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(a.Wv + c0));
    float4 df = make_float4(dv * wv.x, dv * wv.y, dv * wv.z, dv * wv.w);
    if (a.df_ext) {
      const float4 e = *reinterpret_cast<const float4*>(a.df_ext + (long)r * 128 + c0);
      df.x += e.x; df.y += e.y; df.z += e.z; df.w += e.w;
    }
    const float4 f = *reinterpret_cast<const float4*>(a.f + (long)r * 128 + c0);
    atomicAdd(&sWv[c0], dv * f.x); atomicAdd(&sWv[c0 + 1], dv * f.y);
    atomicAdd(&sWv[c0 + 2], dv * f.z); atomicAdd(&sWv[c0 + 3], dv * f.w);
    if (lane == 0) atomicAdd(&sbv, dv);
    const float4 x = *reinterpret_cast<const float4*>(a.x2 + (long)r * a.ld_x2 + c0);
    float4 dx = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(a.W + (long)r * 896 + q * 128 + c0);
      const float dr = warp_sum(df.x * w.x + df.y * w.y + df.z * w.z + df.w * w.w);
      const float rq = a.r[(long)r * 8 + q];
      *reinterpret_cast<float4*>(a.dWc + (long)r * 896 + q * 128 + c0) = make_float4(rq * df.x, rq * df.y, rq * df.z, rq * df.w);
      const float4 wr = __ldg(reinterpret_cast<const float4*>(a.Wr + q * 128 + c0));
      dx.x = fmaf(dr, wr.x, dx.x); dx.y = fmaf(dr, wr.y, dx.y); dx.z = fmaf(dr, wr.z, dx.z); dx.w = fmaf(dr, wr.w, dx.w);
      atomicAdd(&sWr[q][c0], dr * x.x); atomicAdd(&sWr[q][c0 + 1], dr * x.y);
      atomicAdd(&sWr[q][c0 + 2], dr * x.z); atomicAdd(&sWr[q][c0 + 3], dr * x.w);
      if (lane == 0) atomicAdd(&sbr[q], dr);
    }
    *reinterpret_cast<float4*>(a.dx2 + (long)r * a.ld_dx2 + c0) = dx;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 7 * 128; i += 256) atomicAdd(a.dWr + i, (&sWr[0][0])[i]);
  if (threadIdx.x < 128) atomicAdd(a.dWv + threadIdx.x, sWv[threadIdx.x]);
  if (threadIdx.x < 7) atomicAdd(a.dbr + threadIdx.x, sbr[threadIdx.x]);
  if (threadIdx.x == 0) atomicAdd(a.dbv, sbv);
}
int launch_final_bwd(const FinalBwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.x2 && a.Wr && a.W && a.r && a.f && a.Wv && a.dWc && a.dx2 && a.dWr && a.dbr && a.dWv && a.dbv &&
                      a.R > 0,
                  "final_bwd: bad arguments");
  SDUMC_CUDA(launch_kernel(final_bwd_kernel, dim3((a.R + 7) / 8), dim3(256), 0, stream, 1, a));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sdumc
