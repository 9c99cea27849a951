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
  float* Ps = sm;                                   // [L][NQ]
  float* red = sm + ((a.L * NQ + 3) & ~3);          // [8 warps][NQ][G], 16-byte aligned
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L;
  const DropKey key = resolve_key(a.key);
  float* Sg = a.S + (long)b * L * NQ;

  for (int i = tid; i < L * NQ; i += 256) Ps[i] = Sg[i];
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
  // per-warp partials -> shared memory [8][NQ][G], summed by the writer loop below (no atomics)
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float* dst = red + (warp * NQ + q) * G + lane * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
  }
  __syncthreads();

  const uint32_t thr = drop_threshold(a.drop_p);
  const float scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  for (int i = tid; i < NQ * G; i += 256) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) o += red[w * NQ * G + i];
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
  const size_t smem = (size_t)(((a.L * a.nq + 3) & ~3) + 8 * a.nq * G) * sizeof(float);
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
// One CTA per sample, one warp per frame row, a lane owns 8 of the 256 columns.  dO and Qp slices live in
// registers (no shared-memory traffic in the row loop), the sample's probabilities in shared memory, the
// next row's X'/K/dH loads are in flight while a row computes, and the NQ per-row dot products are
// reduced with a halving butterfly (9 shuffles + NQ broadcasts instead of 5 * NQ).
// ------------------------------------------------------------------------------------------
// One CTA per sample.  The sample's X', K (and dH when accumulating) rows are contiguous in memory, so they
// are streamed through a kStages-deep shared-memory ring of 16-row stages filled by 1-D bulk copies
// (cp.async.bulk + mbarrier): the bytes in flight (~100 KB per SM) no longer depend on registers.
// A stage is consumed by all 8 warps, two rows each; a lane owns 8 of the 256 columns.
constexpr int kBwdStages = 4;
constexpr int kBwdRows = 16;                       // rows per stage
constexpr int kBwdTile = kBwdRows * G * 2;         // bytes of one bf16 [16,256] tile

template <int NQ>
__global__ void __launch_bounds__(256, 1) attn_bwd_kernel(AttnBwdArgs a) {
  constexpr int RB = 2;  // rows per warp per stage (independent dependency chains -> ILP)
  extern __shared__ __align__(128) unsigned char dyn[];
  // layout: ring [kStages][3 tiles: X', K, dH_old] | P_s [L][8] | part [8 warps][NQ][G]
  unsigned char* ring = dyn;
  float* P_s = reinterpret_cast<float*>(dyn + kBwdStages * 3 * kBwdTile);
  float* part = P_s + ((a.L * 8 + 3) & ~3);
  __shared__ float dO_s[NQ][G];
  __shared__ float Qp_s[NQ][G];
  __shared__ float red_b[8][G];
  __shared__ float delta_s[8];
  __shared__ __align__(8) uint64_t full_bar[kBwdStages];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L;
  const int n_iter = (L + kBwdRows - 1) / kBwdRows;
  const bool rmw = a.dh_mode == 1;
  const uint32_t thr = drop_threshold(a.out_drop_p);
  const float oscale = a.out_drop_p > 0.f ? 1.f / (1.f - a.out_drop_p) : 1.f;
  const DropKey key = resolve_key(a.key);
  const __nv_bfloat16* Xb = a.X + (long)b * L * G;    // host guarantees ldx == ldk == lddh == 256
  const __nv_bfloat16* Kb = a.Kt + (long)b * L * G;
  const __nv_bfloat16* Hb = a.dH + (long)b * L * G;

  auto issue_stage = [&](int it) {                    // thread 0 only
    const int slot = it % kBwdStages;
    const int rows = min(kBwdRows, L - it * kBwdRows);
    const uint32_t bytes = (uint32_t)rows * G * 2;
    unsigned char* dst = ring + slot * 3 * kBwdTile;
    mbar_expect_tx(&full_bar[slot], bytes * (rmw ? 3u : 2u));
    bulk_load(dst, Xb + (long)it * kBwdRows * G, bytes, &full_bar[slot]);
    bulk_load(dst + kBwdTile, Kb + (long)it * kBwdRows * G, bytes, &full_bar[slot]);
    if (rmw) bulk_load(dst + 2 * kBwdTile, Hb + (long)it * kBwdRows * G, bytes, &full_bar[slot]);
  };
  if (tid == 0) {
    for (int i = 0; i < kBwdStages; ++i) mbar_init(&full_bar[i], 1);
    fence_mbar_init();
    for (int i = 0; i < kBwdStages && i < n_iter; ++i) issue_stage(i);
  }

  for (int i = tid; i < NQ * G; i += 256) {
    float g = a.dOut[(long)b * a.dout_stride_b + i];
    if (a.out_drop_p > 0.f) {
      const uint32_t e = (uint32_t)b * (uint32_t)(NQ * G) + (uint32_t)i;
      g = elem_rand(key, a.out_site, e) >= thr ? g * oscale : 0.f;
    }
    (&dO_s[0][0])[i] = g;
    (&Qp_s[0][0])[i] = __ldg(a.Qp + (long)b * a.qp_stride_b + i);
  }
  for (int i = tid; i < L * 8; i += 256) {
    const int l = i >> 3, q = i & 7;
    P_s[i] = q < NQ ? __ldg(a.P + ((long)b * L + l) * NQ + q) : 0.f;
  }
  if (tid < 8) delta_s[tid] = 0.f;
  __syncthreads();
  for (int q = warp; q < NQ; q += 8) {
    float s = 0.f;
    for (int g = lane; g < G; g += 32) s = fmaf(a.O_pre[(long)b * NQ * G + q * G + g], dO_s[q][g], s);
    s = warp_sum(s);
    if (lane == 0) delta_s[q] = s;
  }
  __syncthreads();

  const int g0 = lane * 8;
  float dO[NQ][8], dq_acc[NQ][8], db_acc[8];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 d0 = *reinterpret_cast<const float4*>(&dO_s[q][g0]);
    const float4 d1 = *reinterpret_cast<const float4*>(&dO_s[q][g0 + 4]);
    dO[q][0] = d0.x; dO[q][1] = d0.y; dO[q][2] = d0.z; dO[q][3] = d0.w;
    dO[q][4] = d1.x; dO[q][5] = d1.y; dO[q][6] = d1.z; dO[q][7] = d1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) dq_acc[q][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) db_acc[j] = 0.f;
  // which of the (up to 8) per-row sums this lane owns after the butterfly
  const int own = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const float my_delta = delta_s[own];

  for (int it = 0; it < n_iter; ++it) {
    const int slot = it % kBwdStages;
    mbar_wait(&full_bar[slot], (uint32_t)((it / kBwdStages) & 1));
    const unsigned char* st = ring + slot * 3 * kBwdTile;
    const int l0 = it * kBwdRows + warp * RB;
    float x[RB][8], k[RB][8];
    uint4 oldh[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int lr = warp * RB + r;   // row inside the stage (rows past L hold stale data; their stores are skipped)
      unpack8(*reinterpret_cast<const uint4*>(st + (lr * G + g0) * 2), x[r]);
      unpack8(*reinterpret_cast<const uint4*>(st + kBwdTile + (lr * G + g0) * 2), k[r]);
      oldh[r] = *reinterpret_cast<const uint4*>(st + 2 * kBwdTile + (lr * G + g0) * 2);
    }
    // partial dot products of this lane's 8 columns, for both rows
    float dS[RB][NQ];
    float Pv[RB][8];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int l = min(l0 + r, L - 1);
      const float4 p0 = *reinterpret_cast<const float4*>(&P_s[l * 8]);
      const float4 p1 = *reinterpret_cast<const float4*>(&P_s[l * 8 + 4]);
      Pv[r][0] = p0.x; Pv[r][1] = p0.y; Pv[r][2] = p0.z; Pv[r][3] = p0.w;
      Pv[r][4] = p1.x; Pv[r][5] = p1.y; Pv[r][6] = p1.z; Pv[r][7] = p1.w;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < NQ) {
          float s = x[r][0] * dO[q][0];
#pragma unroll
          for (int j = 1; j < 8; ++j) s = fmaf(x[r][j], dO[q][j], s);
          v[q] = s;
        } else {
          v[q] = 0.f;
        }
      }
      if (NQ == 1) {
        const float dP = warp_sum(v[0]);
        dS[r][0] = a.alpha * Pv[r][0] * (dP - delta_s[0]);
      } else {
        // halving butterfly: after 4+2+1+2 shuffles lane `own` holds the full sum of index `own`
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        float w[4], u[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float keep = b4 ? v[4 + i] : v[i], send = b4 ? v[i] : v[4 + i];
          w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float keep = b3 ? w[2 + i] : w[i], send = b3 ? w[i] : w[2 + i];
          u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        float t = (b2 ? u[1] : u[0]) + __shfl_xor_sync(0xffffffffu, b2 ? u[0] : u[1], 4);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        const float mine = a.alpha * P_s[l * 8 + own] * (t - my_delta);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
          dS[r][q] = __shfl_sync(0xffffffffu, mine, ((q & 4) ? 16 : 0) + ((q & 2) ? 8 : 0) + ((q & 1) ? 4 : 0));
      }
    }
    // dK, value path, dQp: the Qp slice comes from shared memory once per row pair
    float dK[RB][8], dXv[RB][8];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) { dK[r][j] = 0.f; dXv[r][j] = 0.f; }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4 q0 = *reinterpret_cast<const float4*>(&Qp_s[q][g0]);
      const float4 q1 = *reinterpret_cast<const float4*>(&Qp_s[q][g0 + 4]);
      const float qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const bool live = (l0 + r) < L;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dK[r][j] = fmaf(dS[r][q], qq[j], dK[r][j]);
          dXv[r][j] = fmaf(Pv[r][q], dO[q][j], dXv[r][j]);
          if (live) dq_acc[q][j] = fmaf(dS[r][q], k[r][j], dq_acc[q][j]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int l = l0 + r;
      if (l >= L) break;
      const long row = (long)b * L + l;
      float dZ[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dZ[j] = dK[r][j] * (1.f - k[r][j] * k[r][j]);
        db_acc[j] += dZ[j];
      }
      *reinterpret_cast<uint4*>(a.dZ + row * G + g0) = pack8(dZ);
      if (a.fmask_site) {
        const U4 wm = frame_mask_words(key, a.fmask_site, (uint32_t)row, (uint32_t)(g0 >> 7));
        const int wsel = (g0 >> 5) & 3;
        const uint32_t bits = (wsel == 0 ? wm.x : (wsel == 1 ? wm.y : (wsel == 2 ? wm.z : wm.w))) >> (g0 & 31);
#pragma unroll
        for (int j = 0; j < 8; ++j) dXv[r][j] = ((bits >> j) & 1u) ? 2.f * dXv[r][j] : 0.f;
      }
      if (rmw) {
        float old[8];
        unpack8(oldh[r], old);
#pragma unroll
        for (int j = 0; j < 8; ++j) dXv[r][j] += old[j];
      }
      *reinterpret_cast<uint4*>(a.dH + row * G + g0) = pack8(dXv[r]);
    }
    // every warp is done with this slot: refill it with the stage kStages ahead
    __syncthreads();
    if (tid == 0 && it + kBwdStages < n_iter) issue_stage(it + kBwdStages);
  }

  // per-warp partials -> shared memory, summed without atomics
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float* dst = part + (warp * NQ + q) * G + g0;
    *reinterpret_cast<float4*>(dst) = make_float4(dq_acc[q][0], dq_acc[q][1], dq_acc[q][2], dq_acc[q][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(dq_acc[q][4], dq_acc[q][5], dq_acc[q][6], dq_acc[q][7]);
  }
  *reinterpret_cast<float4*>(&red_b[warp][g0]) = make_float4(db_acc[0], db_acc[1], db_acc[2], db_acc[3]);
  *reinterpret_cast<float4*>(&red_b[warp][g0 + 4]) = make_float4(db_acc[4], db_acc[5], db_acc[6], db_acc[7]);
  __syncthreads();
  for (int i = tid; i < NQ * G; i += 256) {
    float val = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) val += part[w * NQ * G + i];
    if (a.qp_stride_b == 0) atomicAdd(a.dQp + i, val);            // shared context vector: sum over the batch
    else a.dQp[(long)b * a.dqp_stride_b + i] = val;
  }
  if (tid < G) {
    float sdb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sdb += red_b[w][tid];
    atomicAdd(a.db + tid, sdb);
  }
}

static size_t attn_bwd_smem(int L, int nq) {
  return (size_t)kBwdStages * 3 * kBwdTile + (size_t)((L * 8 + 3) & ~3) * 4 + (size_t)8 * nq * G * 4;
}

int launch_attn_bwd(const AttnBwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.X && a.Kt && a.P && a.dOut && a.O_pre && a.Qp && a.dZ && a.dH && a.dQp && a.db,
                  "attn_bwd: null pointer");
  SDUMC_CHECK_ARG(a.B > 0 && a.L > 0 && (a.nq == 1 || a.nq == 7), "attn_bwd: bad shape");
  SDUMC_CHECK_ARG(a.ldx == G && a.ldk == G && a.lddz == G && a.lddh == G,
                  "attn_bwd: frame tensors must be dense [B*L,256] (bulk-copy staging)");
  SDUMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.X) | reinterpret_cast<uintptr_t>(a.Kt) | reinterpret_cast<uintptr_t>(a.dH) |
                    reinterpret_cast<uintptr_t>(a.dZ)) & 15u) == 0, "attn_bwd: frame tensors must be 16-byte aligned");
  const size_t smem = attn_bwd_smem(a.L, a.nq);
  // dynamic + static shared memory must stay under 227 KB (static: 10 KB for nq = 1, 23 KB for nq = 7)
  constexpr size_t kMaxDyn = 200 * 1024;
  SDUMC_CHECK_ARG(smem <= kMaxDyn, "attn_bwd: L=%d too long for the shared-memory probability cache", a.L);
  static bool attr_done = false;
  if (!attr_done) {
    SDUMC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn));
    SDUMC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn));
    attr_done = true;
  }
  if (a.nq == 1) attn_bwd_kernel<1><<<a.B, 256, smem, stream>>>(a);
  else           attn_bwd_kernel<7><<<a.B, 256, smem, stream>>>(a);
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
