// sdumc_b200 — frame-level kernels of the pooling-attention family:
//   softmax over the L frames of a sample + weighted pooling (forward), and the row-wise part of
//   the backward pass.  The dense parts (key projection, dZ*W_in, dZ^T*X') are tcgen05 GEMMs
//   (gemm.cuh); these kernels are HBM-bound streaming kernels: one CTA per sample, one warp per
//   frame row, 16-byte loads (8 bf16 columns per lane).
//
// Reference: FRA2UTT_new.forward / Cross_Attention.forward,
//   toolkit/models/wengnet_mosei_mult_views_text_missing.py:56-68, :79-95 (forward);
//   the backward formulas are autograd of those lines.
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

namespace sdumc {

// general_dim of the model: 256 in the reference (:191); 1024 is BASELINE config 4 ("hidden 1024" stress).  A stage of
// the frame rings always holds 16 K-rows' worth of bytes per warp-row-slab, so the tile shapes scale with G:
//   G = 256 : 64 frame rows per stage = 4 slabs of 16 rows x (2 | 4) column groups of warps
//   G = 1024: 16 frame rows per stage = 1 slab            x (8 | 16) column groups
// and every warp keeps the same fragment loops: 128 columns (forward) / 64 columns (backward) of its slab.
template <int G>
struct FrameCfg {
  static_assert(G == 256 || G == 1024, "general_dim 256 (reference) or 1024 (stress configuration)");
  static constexpr int kRows = 64 * 256 / G;           // frame rows per stage
  static constexpr int kSlabs = kRows / 16;            // 16-row slabs per stage
  static constexpr int kPitch = G + 8;                 // bf16 elements per padded shared-memory row
  static constexpr int kTile = kRows * kPitch * 2;     // bytes of one padded bf16 [kRows, G] tile
  static constexpr int kFwdColWarps = 8 / kSlabs;      // pool_fwd: 8 warps, 128 columns each
  static constexpr int kBwdColWarps = 16 / kSlabs;     // attn_bwd: 16 warps, 64 columns each
};

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
// forward: S = K Qp^T (optional);  P = softmax_L(alpha * S);  O = P^T X';  out = dropout(O)
//
// One CTA (8 warps) per sample.  The sample's K rows (when the scores are computed here) and then its X'
// rows stream through a 2-deep ring of 64-row stages (per-row bulk copies into 528-byte padded rows).
// Both products have one side only NQ (<= 8) wide, so they run on the warp-level tensor-core path with
// the queries padded to 8 - the same fragments as attn_bwd:
//   S   [16 rows x 8 q]   = K[16 x 256] * Qp^T         (per warp: its 128-column half, summed in the pair)
//   O^T [256 x 8 q]      += X'^T[256 x 16 rows] * P    A: ldmatrix.trans of the X' tile, B: movmatrix(P)
// When Kt is NULL the scores are taken from S (written by the key-projection GEMM epilogue, the
// inference path that never materialises K).  P is written back to S for the backward pass.
// ------------------------------------------------------------------------------------------
constexpr int kFwdStages = 2;
constexpr int kFwdThreads = 256;

template <int NQ, int G>
__global__ void __launch_bounds__(kFwdThreads, 2) pool_fwd_kernel(PoolFwdArgs a) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  using FC = FrameCfg<G>;
  constexpr int kFwdRows = FC::kRows, kFwdPitch = FC::kPitch, kFwdTile = FC::kTile, kSlabs = FC::kSlabs;
  constexpr int kCW = FC::kFwdColWarps, kWCols = G / kCW;   // 128 columns per warp
  extern __shared__ __align__(128) unsigned char dyn[];
  // layout: ring [2][tile] | QpB hi, lo [2][8][264] bf16 | S_s [L][8] f32
  // (queries and probabilities enter the tensor-core products as bf16 hi + lo pairs: ~16 mantissa bits for
  //  twice the - negligible - mma work; the frame operands K and X' are bf16 in memory anyway)
  unsigned char* ring = dyn;
  __nv_bfloat16* QpB = reinterpret_cast<__nv_bfloat16*>(dyn + kFwdStages * kFwdTile);
  __nv_bfloat16* QpL = QpB + 8 * kFwdPitch;
  float* S_s = reinterpret_cast<float*>(QpL + 8 * kFwdPitch);
  __shared__ float4 sp_x[kFwdThreads];
  __shared__ __align__(8) uint64_t full_bar[kFwdStages];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform (uniform-register bulk copies)
  const int gid = lane >> 2, tq = lane & 3;
  const int rw = warp % kSlabs, cw = warp / kSlabs, hc = cw * kWCols;   // row slab / column group of this warp
  // dense layout: sample b owns rows [b*L, (b+1)*L).  Varlen (row_off): only its T_b valid frames exist, packed at
  // rows [row_off[b], row_off[b+1]); the a.L - T_b padded frames of the reference's batch enter in closed form below.
  long row0 = (long)b * a.L;
  int L = a.L;
  if (a.row_off) {
    row0 = __ldg(a.row_off + b);
    L = __ldg(a.row_off + b + 1) - (int)row0;
  }
  const int n_pad = a.L - L;
  __shared__ float ppad_s[8];                                // probability mass of the padded frames per query
  const int n_iter = (L + kFwdRows - 1) / kFwdRows;
  const bool has_k = a.Kt != nullptr;
  const int n_total = has_k ? 2 * n_iter : n_iter;
  const DropKey key = resolve_key(a.key);
  const __nv_bfloat16* Xb = a.X + row0 * G;                  // host guarantees dense [rows,G]
  const __nv_bfloat16* Kb = has_k ? a.Kt + row0 * G : nullptr;
  float* Sg = a.S + row0 * NQ;

  auto issue_stage = [&](int j) {                           // lane 0 of warp w: rows w*rpw .. w*rpw + rpw - 1
    constexpr int rpw = kFwdRows / 8;
    const int slot = j % kFwdStages;
    const bool kphase = has_k && j < n_iter;
    const int it = kphase ? j : j - (has_k ? n_iter : 0);
    const int rows = min(kFwdRows, L - it * kFwdRows);
    const __nv_bfloat16* src = (kphase ? Kb : Xb) + (long)it * kFwdRows * G;
    unsigned char* dst = ring + slot * kFwdTile;
    if (lane == 0) {
      if (warp == 0) mbar_expect_tx(&full_bar[slot], (uint32_t)rows * G * 2);
      const int r1 = min(rows, warp * rpw + rpw);
      for (int r = warp * rpw; r < r1; ++r) bulk_load(dst + r * kFwdPitch * 2, src + (long)r * G, G * 2, &full_bar[slot]);
    }
  };
  if (tid == 0) {
    for (int i = 0; i < kFwdStages; ++i) mbar_init(&full_bar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  for (int i = 0; i < kFwdStages && i < n_total; ++i) issue_stage(i);

  // prologue: queries -> bf16 [8][264] (padding queries zero), output staging buffer, and - when the scores
  // are an input - the scores themselves
  if (has_k) {
    constexpr int kV4 = NQ * G / 4;
    const float4* Qg = reinterpret_cast<const float4*>(a.Qp + (long)b * a.qp_stride_b);
    for (int i = tid; i < 8 * G / 4; i += kFwdThreads) {
      const float4 v = i < kV4 ? __ldg(Qg + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int q = (i * 4) / G, g = (i * 4) & (G - 1);
      const uint32_t h0 = pack2(v.x, v.y), h1 = pack2(v.z, v.w);
      *reinterpret_cast<uint2*>(QpB + q * kFwdPitch + g) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(QpL + q * kFwdPitch + g) =
          make_uint2(pack2(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xffff0000u)),
                     pack2(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xffff0000u)));
    }
  } else {
    for (int i0 = 0; i0 < L * NQ; i0 += 8 * kFwdThreads) {
      float sv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int i = i0 + j * kFwdThreads + tid; sv[j] = i < L * NQ ? Sg[i] : 0.f; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * kFwdThreads + tid;
        if (i < L * NQ) S_s[(i / NQ) * 8 + (i % NQ)] = sv[j];
      }
    }
  }
  __syncthreads();

  float acc[8][4];       // O^T fragments: [m-tile of 16 columns][(col gid, q 2tq) (col gid, q 2tq+1) (col gid+8, ...)]
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  for (int j = 0; j < n_total; ++j) {
    const int slot = j % kFwdStages;
    const bool kphase = has_k && j < n_iter;
    const int it = kphase ? j : j - (has_k ? n_iter : 0);
    if (!kphase && it == 0) {
      // ---- softmax over the L frames, one warp per query; P -> S_s (padding queries: 0) and -> global ----
      // (every score of the sample is in S_s: the K phase ended with a __syncthreads, or the prologue did)
      for (int q = warp; q < 8; q += kFwdThreads / 32) {
        if (q < NQ) {
          // score of a padded frame of this sample: s_pad = k_pad . Qp_q (the same for all its padded frames)
          float sp = 0.f;
          if (n_pad > 0) {
            const float* Qg = a.Qp + (long)b * a.qp_stride_b + (long)q * G;
            for (int g = lane; g < G; g += 32) sp = fmaf(__bfloat162float(a.Kpad[g]), __ldg(Qg + g), sp);
            sp = warp_sum(sp);
          }
          float m = n_pad > 0 ? sp : -INFINITY;
          for (int l = lane; l < L; l += 32) m = fmaxf(m, S_s[l * 8 + q]);
          m = warp_max(m);
          float ssum = 0.f;
          for (int l = lane; l < L; l += 32) {
            const float e = __expf(a.alpha * (S_s[l * 8 + q] - m));
            S_s[l * 8 + q] = e;
            ssum += e;
          }
          ssum = warp_sum(ssum);
          const float epad = n_pad > 0 ? (float)n_pad * __expf(a.alpha * (sp - m)) : 0.f;
          const float inv = 1.f / (ssum + epad);
          for (int l = lane; l < L; l += 32) S_s[l * 8 + q] *= inv;
          if (lane == 0) ppad_s[q] = epad * inv;
        } else {
          for (int l = lane; l < L; l += 32) S_s[l * 8 + q] = 0.f;
          if (lane == 0) ppad_s[q] = 0.f;
        }
      }
      __syncthreads();
      for (int i = tid; i < L * NQ; i += kFwdThreads) Sg[i] = S_s[(i / NQ) * 8 + (i % NQ)];
    }
    mbar_wait(&full_bar[slot], (uint32_t)((j / kFwdStages) & 1));
    __nv_bfloat16* Ts = reinterpret_cast<__nv_bfloat16*>(ring + slot * kFwdTile) + rw * 16 * kFwdPitch + hc;
    const int l0 = it * kFwdRows + rw * 16;
    const int valid = min(16, L - l0);
    if (valid > 0) {
      if (kphase) {
        // scores of this warp's 16 rows: partial over its 128 columns, then add the other column groups' partials
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
        const __nv_bfloat16* arow = Ts + ((lane & 7) + ((lane >> 3) & 1) * 8) * kFwdPitch + (lane >> 4) * 8;
        const __nv_bfloat16* brow = QpB + gid * kFwdPitch + hc + 2 * tq;
#pragma unroll
        for (int kk = 0; kk < kWCols / 16; ++kk) {
          uint32_t af[4];
          ldsm_x4(af, arow + kk * 16);
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(brow + kk * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(brow + kk * 16 + 8);
          mma_16816(sc, af, b0, b1);
          const uint32_t c0 = *reinterpret_cast<const uint32_t*>(brow + 8 * kFwdPitch + kk * 16);
          const uint32_t c1 = *reinterpret_cast<const uint32_t*>(brow + 8 * kFwdPitch + kk * 16 + 8);
          mma_16816(sc, af, c0, c1);
        }
        sp_x[tid] = make_float4(sc[0], sc[1], sc[2], sc[3]);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + rw), "n"(kCW * 32) : "memory");
        float4 o = sp_x[rw * 32 + lane];                           // column groups summed in a fixed order
#pragma unroll
        for (int c = 1; c < kCW; ++c) {
          const float4 x = sp_x[(c * kSlabs + rw) * 32 + lane];
          o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
        }
        // column group 0 writes rows gid, group 1 rows gid+8 (rows past L hold garbage: skipped)
        if (cw == 0) { if (gid < valid) *reinterpret_cast<float2*>(&S_s[(l0 + gid) * 8 + 2 * tq]) = make_float2(o.x, o.y); }
        else if (cw == 1 && gid + 8 < valid) *reinterpret_cast<float2*>(&S_s[(l0 + gid + 8) * 8 + 2 * tq]) = make_float2(o.z, o.w);
      } else {
        if (valid < 16) {                             // rows past L: zero so 0 * garbage cannot produce NaN
          for (int i = lane; i < 16 * (kWCols / 8); i += 32) {
            const int r = i / (kWCols / 8), c = (i % (kWCols / 8)) * 8;
            if (r >= valid) *reinterpret_cast<uint4*>(Ts + r * kFwdPitch + c) = make_uint4(0, 0, 0, 0);
          }
          __syncwarp();
        }
        const bool v0 = gid < valid, v1 = gid + 8 < valid;
        const float2 p0 = v0 ? *reinterpret_cast<const float2*>(&S_s[(l0 + gid) * 8 + 2 * tq]) : make_float2(0.f, 0.f);
        const float2 p1 = v1 ? *reinterpret_cast<const float2*>(&S_s[(l0 + gid + 8) * 8 + 2 * tq]) : make_float2(0.f, 0.f);
        const uint32_t h0 = pack2(p0.x, p0.y), h1 = pack2(p1.x, p1.y);
        const uint32_t l0p = pack2(p0.x - __uint_as_float(h0 << 16), p0.y - __uint_as_float(h0 & 0xffff0000u));
        const uint32_t l1p = pack2(p1.x - __uint_as_float(h1 << 16), p1.y - __uint_as_float(h1 & 0xffff0000u));
        const uint32_t bt0 = movmatrix_trans(h0), bt1 = movmatrix_trans(h1);
        const uint32_t bl0 = movmatrix_trans(l0p), bl1 = movmatrix_trans(l1p);
        const __nv_bfloat16* arow = Ts + ((lane & 7) + (lane >> 4) * 8) * kFwdPitch + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
          uint32_t af[4];
          ldsm_x4_trans(af, arow + mt * 16);
          mma_16816(acc[mt], af, bt0, bt1);
          mma_16816(acc[mt], af, bl0, bl1);
        }
      }
    }
    __syncthreads();                                  // slot consumed by every warp (and S_s complete after the K phase)
    if (j + kFwdStages < n_total) issue_stage(j + kFwdStages);
  }

  // O: fragments of the four row-slab warps -> per-slab partials in the (now idle) ring -> summed in a fixed
  // order (bitwise reproducible, unlike shared-memory atomics) -> global (+ output dropout)
  float* part = reinterpret_cast<float*>(ring);            // [kSlabs][8 q][G]: 32 KB of the 66 KB ring
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    const int c = hc + mt * 16 + gid, q = 2 * tq;
    float* pr = part + rw * 8 * G;
    pr[q * G + c] = acc[mt][0];       pr[q * G + c + 8] = acc[mt][2];
    pr[(q + 1) * G + c] = acc[mt][1]; pr[(q + 1) * G + c + 8] = acc[mt][3];
  }
  __syncthreads();
  const uint32_t thr = drop_threshold(a.drop_p);
  const float scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  for (int i = tid; i < NQ * G; i += kFwdThreads) {
    float o;
    if constexpr (kSlabs == 4) o = (part[i] + part[8 * G + i]) + (part[2 * 8 * G + i] + part[3 * 8 * G + i]);
    else o = part[i];
    if (n_pad > 0) o = fmaf(ppad_s[i / G], __bfloat162float(a.Hpad[i % G]), o);   // the padded frames' pooled share
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
  if (a.row_off)
    SDUMC_CHECK_ARG(a.Hpad && a.Kpad && a.Qp && a.drop_p == 0.f,
                    "pool_fwd: the varlen layout needs Hpad, Kpad and Qp and is defined for eval mode only (no dropout)");
  const int G = a.G > 0 ? a.G : 256;
  SDUMC_CHECK_ARG(G == 256 || G == 1024, "pool_fwd: general_dim %d unsupported (256 or 1024)", G);
  SDUMC_CHECK_ARG(a.ldx == G && (reinterpret_cast<uintptr_t>(a.X) & 15u) == 0,
                  "pool_fwd: X must be dense [B*L,G] and 16-byte aligned (bulk-copy staging)");
  if (a.Kt) {
    SDUMC_CHECK_ARG(a.ldk == G && (reinterpret_cast<uintptr_t>(a.Kt) & 15u) == 0,
                    "pool_fwd: Kt must be dense [B*L,G] and 16-byte aligned");
    SDUMC_CHECK_ARG(a.Qp && (reinterpret_cast<uintptr_t>(a.Qp) & 15u) == 0 && a.qp_stride_b % 4 == 0,
                    "pool_fwd: Qp is required with Kt (16-byte aligned, stride a multiple of 4)");
  }
  const size_t tile = G == 256 ? FrameCfg<256>::kTile : FrameCfg<1024>::kTile;
  const size_t smem = (size_t)kFwdStages * tile + (size_t)2 * 8 * (G + 8) * 2 + (size_t)a.L * 8 * 4;
  constexpr size_t kMaxDyn = 216 * 1024;
  SDUMC_CHECK_ARG(smem <= kMaxDyn, "pool_fwd: L=%d too long for the shared-memory softmax", a.L);
  static bool attr_done[kMaxDevices] = {false};
  const int dev = current_device();
  if (!attr_done[dev]) {
    SDUMC_CUDA(cudaFuncSetAttribute(pool_fwd_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn));
    SDUMC_CUDA(cudaFuncSetAttribute(pool_fwd_kernel<7, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn));
    SDUMC_CUDA(cudaFuncSetAttribute(pool_fwd_kernel<1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn));
    SDUMC_CUDA(cudaFuncSetAttribute(pool_fwd_kernel<7, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn));
    attr_done[dev] = true;
  }
  if (G == 256) {
    if (a.nq == 1) SDUMC_CUDA(launch_kernel(pool_fwd_kernel<1, 256>, dim3(a.B), dim3(kFwdThreads), smem, stream, 1, a));
    else           SDUMC_CUDA(launch_kernel(pool_fwd_kernel<7, 256>, dim3(a.B), dim3(kFwdThreads), smem, stream, 1, a));
  } else {
    if (a.nq == 1) SDUMC_CUDA(launch_kernel(pool_fwd_kernel<1, 1024>, dim3(a.B), dim3(kFwdThreads), smem, stream, 1, a));
    else           SDUMC_CUDA(launch_kernel(pool_fwd_kernel<7, 1024>, dim3(a.B), dim3(kFwdThreads), smem, stream, 1, a));
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
// The four products with the NQ (<= 8) queries are skinny matrix products (one side 8 wide), ~7K MACs per
// frame row: far too many for fp32 SIMT at HBM speed and >90% padding for a tcgen05 tile, so they run on
// the warp-level tensor-core path (mma.sync m16n8k16 / m16n8k8, bf16 in, fp32 accumulate) with the
// queries padded to 8:
//   dP  [16 rows x 8 q]    = X'[16 x 256] * dO^T          A: ldmatrix from the X' tile
//   dQp^T [256 x 8 q]     += K^T[256 x 16 rows] * dS      A: ldmatrix.trans from the K tile, B: movmatrix(dS)
//   dK  [16 x 256]         = dS[16 x 8] * Qp              A: the dP accumulator fragment, reused as operand
//   dXv [16 x 256]         = P [16 x 8] * dO
// 16 compute warps + 1 DMA warp per CTA; a stage is 64 frame rows of X' and K in a 2-deep shared-memory ring.  The
// frame tensors are addressed through [B, L, G] tensor maps: a stage = 2 * G/64 SWIZZLE_128B boxes of 64 columns x one
// stage of rows of ONE sample (rows past L: zero-filled on load, clipped on store), conflict-free for ldmatrix
// without padding (chunk ^ (row & 7)).  Compute warp w works on rows 16*(w&3).. of the stage and on column quarter
// (w>>2) = box (w>>2) of every product (4 warps per scheduler hide the dependent-chain latency that bounded the
// 8-warp version); the only exchange inside a quartet is the 16x8 partial dP (each quarter reduces over its own 64
// columns).  dZ and the value-path gradient are written back into the boxes in place; a warp releases the slot with
// one mbarrier arrive, and the DMA warp - which issues every copy of the CTA - sends the boxes out with tensor stores
// (the "+= old dH" of the accumulate mode is a tensor reduce-add performed at L2, so the old gradient is never read
// by the SM), waits until they have read shared memory and refills the slot with the unit two ahead.
constexpr int kBwdStages = 2;
constexpr int kBwdThreads = 512;

template <int NQ, int G>
__global__ void __launch_bounds__(kBwdThreads + 32, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmH, AttnBwdArgs a) {
  using FC = FrameCfg<G>;
  constexpr int kBwdRows = FC::kRows;      // rows per stage (16 per slab)
  constexpr int kBwdPitch = FC::kPitch;    // bf16 elements per padded row of the small dO operand copy
  constexpr int kBoxes = G / 64;           // 64-column boxes per tile: one per column group of warps
  constexpr int kBox = kBwdRows * 128;     // bytes of one SWIZZLE_128B box [kBwdRows rows x 64 columns]
  constexpr int kBwdTile = kBoxes * kBox;  // bytes of one bf16 [kBwdRows, G] tile (32 KB)
  constexpr int kSlabs = FC::kSlabs, kCW = FC::kBwdColWarps;   // row slabs / column groups (64 columns per warp)
  static_assert(kCW == kBoxes, "one column group of warps per box");
  extern __shared__ unsigned char dyn_raw[];
  // layout: ring [2 stages][X' tile, K tile] (1024-byte aligned: swizzle atoms) | dO_b [8][G+8] | dOT [G][8] | QpT [G][8] |
  //         dqp_s [8][G] f32 | P_s [L][8] f32
  unsigned char* dyn = dyn_raw + ((1024u - (smem_u32(dyn_raw) & 1023u)) & 1023u);
  unsigned char* ring = dyn;
  __nv_bfloat16* dO_b = reinterpret_cast<__nv_bfloat16*>(dyn + kBwdStages * 2 * kBwdTile);
  __nv_bfloat16* dOT = dO_b + 8 * kBwdPitch;
  __nv_bfloat16* QpT = dOT + G * 8;
  float* dqp_s = reinterpret_cast<float*>(QpT + G * 8);
  float* P_s = dqp_s + 8 * G;
  __shared__ float delta_s[8];
  __shared__ __align__(8) uint64_t full_bar[kBwdStages];   // TMA -> compute warps: the stage has landed
  __shared__ __align__(8) uint64_t done_bar[kBwdStages];   // compute warps (16 arrivals) -> DMA warp: the stage may leave
  auto compute_sync = [] { asm volatile("bar.sync 9, %0;" ::"n"(kBwdThreads) : "memory"); };   // the 16 compute warps
  const int tid = threadIdx.x, lane = tid & 31;
  // the broadcast makes the warp index provably warp-uniform, so bulk-copy addresses derived from it live in
  // uniform registers (no per-lane serialisation loop around UBLKCP)
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int gid = lane >> 2, tq = lane & 3;           // fragment coordinates: row group / column pair
  const int rw = warp % kSlabs, qc = warp / kSlabs;     // row slab of the stage / column group
  const int hc = qc * 64;                              // first column of this warp's 64-column group
  // partial dP exchange inside a slab.  The warps of a slab are only loosely coupled (one barrier per stage), so the
  // buffer is double-buffered by stage parity; at G = 1024 (32 KB more of per-sample operands) there is no room for
  // the second copy and a second slab barrier after the reads takes its place
  constexpr int kDpBufs = G == 256 ? 2 : 1;
  __shared__ float4 dp_xx[kDpBufs][kBwdThreads];
  const int L = a.L;
  const int n_iter = (L + kBwdRows - 1) / kBwdRows;
  const bool rmw = a.dh_mode == 1;
  const uint32_t thr = drop_threshold(a.out_drop_p);
  const float oscale = a.out_drop_p > 0.f ? 1.f / (1.f - a.out_drop_p) : 1.f;
  const DropKey key = resolve_key(a.key);
  // persistent CTAs over (sample, 64-row stage) units: a contiguous, equal share per CTA, so neither the grid
  // (waves of whole samples) nor the per-sample prologue quantises the work; the ring runs across samples
  const int n_units = a.B * n_iter;
  const int u_begin = (int)(((long)n_units * blockIdx.x) / gridDim.x);
  const int my_units = (int)(((long)n_units * (blockIdx.x + 1)) / gridDim.x) - u_begin;

  // Stage k = unit u_begin + k: 2 * kBoxes TMA boxes (X' boxes, then K boxes) of [kBwdRows rows x 64 columns], each
  // landing 128B-swizzled (conflict-free ldmatrix without padding).
  constexpr int kNB = 2 * kBoxes;
  auto issue_box = [&](int k, int i) {
    const int u = u_begin + k;
    const int ub = u / n_iter, it = u - ub * n_iter;
    const int slot = k % kBwdStages;
    unsigned char* dst = ring + slot * 2 * kBwdTile + i * kBox;
    const int c = i < kBoxes ? i : i - kBoxes;
    tma_load_3d(dst, i < kBoxes ? &tmX : &tmK, c * 64, it * kBwdRows, ub, &full_bar[slot]);
  };
  auto issue_stage = [&](int k) {                     // one thread (the DMA warp's lane 0)
    mbar_expect_tx(&full_bar[k % kBwdStages], 2u * kBwdTile);   // rows past L are zero-filled, their bytes count
#pragma unroll 1
    for (int i = 0; i < kNB; ++i) issue_box(k, i);
  };
  if (tid == 0) {
    for (int i = 0; i < kBwdStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&done_bar[i], kBwdThreads / 32); }
    fence_mbar_init();
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmH);
  }
  if (tid < kBwdThreads) for (int i = tid; i < 8 * G; i += kBwdThreads) dqp_s[i] = 0.f;
  __syncthreads();
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();

  if (warp == kBwdThreads / 32) {
    // ===================== DMA warp: every tensor copy of the CTA =====================
    // loads of unit k into slot k % 2 (2 * kBoxes boxes, one expect_tx); when all 16 compute warps have released the
    // slot (done_bar), its boxes leave through tensor stores (dZ: store; dH: store, or reduce-add at L2 in accumulate
    // mode; rows past L are clipped by the tensor map) and, once those have read shared memory, the slot is refilled
    // with the unit two ahead.  The compute warps never wait for a store.
    if (lane == 0) {
      for (int i = 0; i < kBwdStages && i < my_units; ++i) issue_stage(i);
      for (int k = 0; k < my_units; ++k) {
        const int u = u_begin + k;
        const int ub = u / n_iter, it = u - ub * n_iter;
        const int slot = k % kBwdStages;
        mbar_wait(&done_bar[slot], (uint32_t)((k / kBwdStages) & 1));
        const unsigned char* st = ring + slot * 2 * kBwdTile;
#pragma unroll 1
        for (int i = 0; i < kNB; ++i) {
          const unsigned char* src = st + i * kBox;
          if (i < kBoxes) {
            if (rmw) tma_reduce_add_3d(&tmH, i * 64, it * kBwdRows, ub, src);
            else     tma_store_3d(&tmH, i * 64, it * kBwdRows, ub, src);
          } else {
            tma_store_3d(&tmZ, (i - kBoxes) * 64, it * kBwdRows, ub, src);
          }
        }
        bulk_commit();
        if (k + kBwdStages < my_units) {
          bulk_wait_read();
          issue_stage(k + kBwdStages);
        }
      }
      bulk_wait_all();
    }
    return;
  }

  // per-sample constants: masked dO (bf16, both layouts), Qp^T (bf16), probabilities, delta.
  // All global loads of a batch are issued before the first use: with one CTA per SM this prologue is pure
  // latency, and a load-use-load-use loop costs one DRAM round trip per iteration.  (The stage ring keeps
  // loading meanwhile.)  Called by all threads after a __syncthreads.
  constexpr int kV4 = NQ * G / 4;                     // float4 groups of dO / Qp / O_pre
  constexpr int kPer = (kV4 + kBwdThreads - 1) / kBwdThreads;
  float dl0 = 0.f, dl1 = 0.f;
  auto load_sample = [&](int b) {
  float dpart[kPer];                                  // this thread's partial <O_pre, dO> per group (one query each)
  {
    const float4* dOg = reinterpret_cast<const float4*>(a.dOut + (long)b * a.dout_stride_b);
    const float4* Qg = reinterpret_cast<const float4*>(a.Qp + (long)b * a.qp_stride_b);
    const float4* Og = reinterpret_cast<const float4*>(a.O_pre + (long)b * NQ * G);
    float4 dv[kPer], qv[kPer], ov[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = tid + j * kBwdThreads;
      if (i < kV4) { dv[j] = __ldg(dOg + i); qv[j] = __ldg(Qg + i); ov[j] = __ldg(Og + i); }
      else { dv[j] = qv[j] = ov[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    }
    // probabilities: [L, NQ] contiguous per sample -> P_s [L][8]
    const float* Pg = a.P + (long)b * L * NQ;
    for (int i0 = 0; i0 < L * NQ; i0 += 8 * kBwdThreads) {
      float pv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int i = i0 + j * kBwdThreads + tid; pv[j] = i < L * NQ ? __ldg(Pg + i) : 0.f; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * kBwdThreads + tid;
        if (i < L * NQ) P_s[(i / NQ) * 8 + (i % NQ)] = pv[j];
      }
    }
    if (NQ < 8) for (int i = tid; i < L * (8 - NQ); i += kBwdThreads) P_s[(i / (8 - NQ)) * 8 + NQ + i % (8 - NQ)] = 0.f;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = tid + j * kBwdThreads;
      dpart[j] = 0.f;
      if (i < kV4) {
        const int q = (i * 4) / G, g = (i * 4) & (G - 1);
        float d[4] = {dv[j].x, dv[j].y, dv[j].z, dv[j].w};
        if (a.out_drop_p > 0.f) {
          // four consecutive elements share one Philox counter (elem_rand: counter e >> 2, word e & 3); stacked passes
          // (split_b): the second problem draws from its own site with its own sample index
          const bool second = a.split_b > 0 && b >= a.split_b;
          const uint32_t e = (uint32_t)(second ? b - a.split_b : b) * (uint32_t)(NQ * G) + (uint32_t)(i * 4);
          const U4 r = philox4x32_10(e >> 2, 0x5D0Cu, second ? a.out_site2 : a.out_site, key.step, key.seed_lo, key.seed_hi);
          d[0] = r.x >= thr ? d[0] * oscale : 0.f; d[1] = r.y >= thr ? d[1] * oscale : 0.f;
          d[2] = r.z >= thr ? d[2] * oscale : 0.f; d[3] = r.w >= thr ? d[3] * oscale : 0.f;
        }
        const float qq[4] = {qv[j].x, qv[j].y, qv[j].z, qv[j].w};
        const float oo[4] = {ov[j].x, ov[j].y, ov[j].z, ov[j].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __nv_bfloat16 dvb = __float2bfloat16(d[c]);
          dO_b[q * kBwdPitch + g + c] = dvb;
          dOT[(g + c) * 8 + q] = dvb;
          QpT[(g + c) * 8 + q] = __float2bfloat16(qq[c]);
          // delta uses the same bf16-rounded dO as the dP product, so sum_l P_l (dP_l - delta) = 0 holds exactly
          dpart[j] = fmaf(oo[c], __bfloat162float(dvb), dpart[j]);
        }
      }
    }
  }
  // zero the padding queries
  for (int i = tid; i < (8 - NQ) * G; i += kBwdThreads) {
    const int q = NQ + i / G, g = i % G;
    dO_b[q * kBwdPitch + g] = __float2bfloat16(0.f);
    dOT[g * 8 + q] = __float2bfloat16(0.f);
    QpT[g * 8 + q] = __float2bfloat16(0.f);
  }
  if (tid < 8) delta_s[tid] = 0.f;
  compute_sync();
  // a float4 group lies inside one query row (64 groups per query): warp-reduce, then one atomic per warp and query
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = tid + j * kBwdThreads;           // warp-uniform q: 32 consecutive groups never straddle a row of G/4
    const float s = warp_sum(dpart[j]);
    if (lane == 0 && i < kV4) atomicAdd(&delta_s[(i * 4) / G], s);
  }
  compute_sync();
  dl0 = delta_s[2 * tq];
  dl1 = delta_s[2 * tq + 1];
  };

  float dq_acc[4][4];    // dQp^T fragments: [m-tile of 16 columns][(col gid, q 2tq) (col gid, q 2tq+1) (col gid+8, ...)]
  float db_acc[8][2];    // column sums of dZ over this thread's rows, per 8-column tile
#pragma unroll
  for (int i = 0; i < 4; ++i) { dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f; }
#pragma unroll
  for (int i = 0; i < 8; ++i) { db_acc[i][0] = db_acc[i][1] = 0.f; }

  // dQp of the finished sample: fragments of the four row-slab warps -> shared memory -> global atomics
  // (a sample may be shared with the neighbouring CTAs; the host zero-fills dQp)
  auto flush_dqp = [&](int b) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int c = hc + mt * 16 + gid, q = 2 * tq;
      if (q < NQ)     { atomicAdd(&dqp_s[q * G + c], dq_acc[mt][0]);       atomicAdd(&dqp_s[q * G + c + 8], dq_acc[mt][2]); }
      if (q + 1 < NQ) { atomicAdd(&dqp_s[(q + 1) * G + c], dq_acc[mt][1]); atomicAdd(&dqp_s[(q + 1) * G + c + 8], dq_acc[mt][3]); }
      dq_acc[mt][0] = dq_acc[mt][1] = dq_acc[mt][2] = dq_acc[mt][3] = 0.f;
    }
    compute_sync();
    float* dst = a.dQp + (a.qp_stride_b == 0 ? 0 : (long)b * a.dqp_stride_b);
    for (int i = tid; i < NQ * G; i += kBwdThreads) {
      atomicAdd(dst + i, dqp_s[i]);
      dqp_s[i] = 0.f;
    }
    compute_sync();
  };

  int b = -1;
  for (int k = 0; k < my_units; ++k) {
    const int u = u_begin + k;
    const int ub = u / n_iter, it = u - ub * n_iter;
    if (ub != b) {                                    // uniform over the compute warps; the barriers separate the samples
      if (b >= 0) {
        // per-sample queries: the finished sample's dQp leaves now; a query shared by all samples (FRA2UTT context
        // vector) keeps accumulating in the fragments and leaves once, at the end
        if (a.qp_stride_b != 0) flush_dqp(b);
        else compute_sync();                          // every warp is done with the old sample's shared-memory operands
      }
      b = ub;
      load_sample(b);
    }
    const int slot = k % kBwdStages;
    mbar_wait(&full_bar[slot], (uint32_t)((k / kBwdStages) & 1));
    unsigned char* st = ring + slot * 2 * kBwdTile;
    // this warp's 16 rows x 64 columns: rows rw*16.. of box qc.  Byte address of (row r, 16-byte chunk c) inside a box:
    // r * 128 + ((c ^ (r & 7)) << 4); every row this lane touches has r & 7 == lane & 7 (ldmatrix) or gid (fragments)
    unsigned char* Xs = st + qc * kBox + rw * 16 * 128;
    unsigned char* Ks = st + kBwdTile + qc * kBox + rw * 16 * 128;
    const int l0 = it * kBwdRows + rw * 16;           // first frame of this warp's slab
    const int valid = min(16, L - l0);                // may be <= 0 for a trailing warp (rows past L arrive as zeros)
    if (valid > 0) {
      // (1) dP = X' * dO^T: this warp's 64 columns, then add the partials of the other column groups
      float dP[4] = {0.f, 0.f, 0.f, 0.f};
      {
        float4* dp_x = dp_xx[kDpBufs == 2 ? (k & 1) : 0];
        const unsigned char* arow = Xs + ((lane & 7) + ((lane >> 3) & 1) * 8) * 128;
        const int achunk = lane >> 4, sw = lane & 7;
        const __nv_bfloat16* brow = dO_b + gid * kBwdPitch + hc + 2 * tq;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t af[4];
          ldsm_x4(af, arow + (((kk * 2 + achunk) ^ sw) << 4));
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(brow + kk * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(brow + kk * 16 + 8);
          mma_16816(dP, af, b0, b1);
        }
        dp_x[tid] = make_float4(dP[0], dP[1], dP[2], dP[3]);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + rw), "n"(kCW * 32) : "memory");
        // every column group adds the partials in the same order: identical dS in every warp of the slab
        if constexpr (kCW == 4) {
          const float4 o0 = dp_x[(rw << 5) + lane], o1 = dp_x[128 + (rw << 5) + lane];
          const float4 o2 = dp_x[256 + (rw << 5) + lane], o3 = dp_x[384 + (rw << 5) + lane];
          dP[0] = (o0.x + o1.x) + (o2.x + o3.x); dP[1] = (o0.y + o1.y) + (o2.y + o3.y);
          dP[2] = (o0.z + o1.z) + (o2.z + o3.z); dP[3] = (o0.w + o1.w) + (o2.w + o3.w);
        } else {
          float4 o = dp_x[rw * 32 + lane];
#pragma unroll
          for (int c = 1; c < kCW; ++c) {
            const float4 x = dp_x[(c * kSlabs + rw) * 32 + lane];
            o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
          }
          dP[0] = o.x; dP[1] = o.y; dP[2] = o.z; dP[3] = o.w;
        }
        if constexpr (kDpBufs == 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + rw), "n"(kCW * 32) : "memory");
      }
      // (2) dS = alpha * P * (dP - delta) on the accumulator layout: rows gid / gid+8, queries 2tq / 2tq+1
      const bool v0 = gid < valid, v1 = gid + 8 < valid;
      const float2 p0 = v0 ? *reinterpret_cast<const float2*>(&P_s[(l0 + gid) * 8 + 2 * tq]) : make_float2(0.f, 0.f);
      const float2 p1 = v1 ? *reinterpret_cast<const float2*>(&P_s[(l0 + gid + 8) * 8 + 2 * tq]) : make_float2(0.f, 0.f);
      const uint32_t dS_lo = pack2(v0 ? a.alpha * p0.x * (dP[0] - dl0) : 0.f, v0 ? a.alpha * p0.y * (dP[1] - dl1) : 0.f);
      const uint32_t dS_hi = pack2(v1 ? a.alpha * p1.x * (dP[2] - dl0) : 0.f, v1 ? a.alpha * p1.y * (dP[3] - dl1) : 0.f);
      const uint32_t P_lo = pack2(p0.x, p0.y), P_hi = pack2(p1.x, p1.y);
      // (4) dQp^T += K^T * dS  (before the K tile is overwritten by dZ)
      {
        const uint32_t bt0 = movmatrix_trans(dS_lo), bt1 = movmatrix_trans(dS_hi);
        const unsigned char* arow = Ks + ((lane & 7) + (lane >> 4) * 8) * 128;
        const int achunk = (lane >> 3) & 1, sw = lane & 7;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          uint32_t af[4];
          ldsm_x4_trans(af, arow + (((mt * 2 + achunk) ^ sw) << 4));
          mma_16816(dq_acc[mt], af, bt0, bt1);
        }
      }
      // frame-mask words of this thread's two rows (the 128-column block holding this warp's quarter)
      // (one Philox call per lane: lane i draws the words of slab row i & 15, the owners of rows gid / gid + 8 fetch
      // them by shuffle - the four lanes of a row group used to draw the same two rows each)
      U4 ma, mb;
      if (a.fmask_site) {
        const bool second = a.split_b > 0 && b >= a.split_b;
        const uint32_t r0 = (uint32_t)((long)(second ? b - a.split_b : b) * L + l0);
        const U4 mine = frame_mask_words(key, second ? a.fmask_site2 : a.fmask_site, r0 + (uint32_t)(lane & 15),
                                         (uint32_t)(hc >> 7));
        ma.x = __shfl_sync(0xffffffffu, mine.x, gid);     mb.x = __shfl_sync(0xffffffffu, mine.x, gid + 8);
        ma.y = __shfl_sync(0xffffffffu, mine.y, gid);     mb.y = __shfl_sync(0xffffffffu, mine.y, gid + 8);
        ma.z = __shfl_sync(0xffffffffu, mine.z, gid);     mb.z = __shfl_sync(0xffffffffu, mine.z, gid + 8);
        ma.w = __shfl_sync(0xffffffffu, mine.w, gid);     mb.w = __shfl_sync(0xffffffffu, mine.w, gid + 8);
      }
      // (3) dK = dS * Qp, dXv = P * dO, eight columns at a time; dZ and the masked dXv replace K and X' in place
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int off = gid * 128 + ((nt ^ gid) << 4) + tq * 4;   // columns nt*8 + 2tq, +1 of row gid (row gid + 8: + 1024)
        const uint32_t bq = *reinterpret_cast<const uint32_t*>(QpT + (hc + nt * 8 + gid) * 8 + 2 * tq);
        const uint32_t bo = *reinterpret_cast<const uint32_t*>(dOT + (hc + nt * 8 + gid) * 8 + 2 * tq);
        float dK[4], dX[4];
        mma_1688(dK, dS_lo, dS_hi, bq);
        mma_1688(dX, P_lo, P_hi, bo);
        uint32_t* k0p = reinterpret_cast<uint32_t*>(Ks + off);
        uint32_t* k1p = reinterpret_cast<uint32_t*>(Ks + off + 8 * 128);
        const uint32_t kv0 = *k0p, kv1 = *k1p;
        const float k00 = __uint_as_float(kv0 << 16), k01 = __uint_as_float(kv0 & 0xffff0000u);
        const float k10 = __uint_as_float(kv1 << 16), k11 = __uint_as_float(kv1 & 0xffff0000u);
        const float z00 = dK[0] * (1.f - k00 * k00), z01 = dK[1] * (1.f - k01 * k01);
        const float z10 = dK[2] * (1.f - k10 * k10), z11 = dK[3] * (1.f - k11 * k11);
        db_acc[nt][0] += z00 + z10;
        db_acc[nt][1] += z01 + z11;
        *k0p = pack2(z00, z01);
        *k1p = pack2(z10, z11);
        if (a.fmask_site) {
          const int w = (((hc & 127) >> 5) + (nt >> 2)) & 3;   // 32-column word inside the 128-column block
          const uint32_t wa = (w == 0 ? ma.x : (w == 1 ? ma.y : (w == 2 ? ma.z : ma.w))) >> ((nt & 3) * 8 + 2 * tq);
          const uint32_t wb = (w == 0 ? mb.x : (w == 1 ? mb.y : (w == 2 ? mb.z : mb.w))) >> ((nt & 3) * 8 + 2 * tq);
          dX[0] = (wa & 1u) ? 2.f * dX[0] : 0.f; dX[1] = (wa & 2u) ? 2.f * dX[1] : 0.f;
          dX[2] = (wb & 1u) ? 2.f * dX[2] : 0.f; dX[3] = (wb & 2u) ? 2.f * dX[3] : 0.f;
        }
        *reinterpret_cast<uint32_t*>(Xs + off) = pack2(dX[0], dX[1]);
        *reinterpret_cast<uint32_t*>(Xs + off + 8 * 128) = pack2(dX[2], dX[3]);
      }
    }
    // the warp's part of the stage is final in shared memory: release the slot to the DMA warp and go on
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(&done_bar[slot]);
  }
  if (b >= 0) flush_dqp(b);
  // db: reduce over the 8 row groups of the warp, then over the warps through shared memory
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = db_acc[nt][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      db_acc[nt][j] = v;
    }
  }
  float* db_s = dqp_s;                                // the dQp staging buffer is idle now (flushed and re-zeroed)
  for (int i = tid; i < G; i += kBwdThreads) db_s[i] = 0.f;
  compute_sync();
  if (gid == 0) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      atomicAdd(&db_s[hc + nt * 8 + 2 * tq], db_acc[nt][0]);
      atomicAdd(&db_s[hc + nt * 8 + 2 * tq + 1], db_acc[nt][1]);
    }
  }
  compute_sync();
  for (int i = tid; i < G; i += kBwdThreads) atomicAdd(a.db + i, db_s[i]);
}

static size_t attn_bwd_smem(int L, int G) {
  const size_t tile = 64 * 256 * 2;   // bf16 [rows per stage, G] = 32 KB for both widths (unpadded: swizzled boxes)
  return 1024 /* alignment of the swizzle atoms */ + (size_t)kBwdStages * 2 * tile + (size_t)8 * (G + 8) * 2 +
         (size_t)2 * G * 8 * 2 + (size_t)8 * G * 4 + (size_t)L * 8 * 4;
}

int launch_attn_bwd(const AttnBwdArgs& a, cudaStream_t stream) {
  SDUMC_CHECK_ARG(a.X && a.Kt && a.P && a.dOut && a.O_pre && a.Qp && a.dZ && a.dH && a.dQp && a.db,
                  "attn_bwd: null pointer");
  SDUMC_CHECK_ARG(a.B > 0 && a.L > 0 && (a.nq == 1 || a.nq == 7), "attn_bwd: bad shape");
  const int G = a.G > 0 ? a.G : 256;
  SDUMC_CHECK_ARG(G == 256 || G == 1024, "attn_bwd: general_dim %d unsupported (256 or 1024)", G);
  SDUMC_CHECK_ARG(a.ldx == G && a.ldk == G && a.lddz == G && a.lddh == G,
                  "attn_bwd: frame tensors must be dense [B*L,G] (tensor-map staging)");
  SDUMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.X) | reinterpret_cast<uintptr_t>(a.Kt) | reinterpret_cast<uintptr_t>(a.dH) |
                    reinterpret_cast<uintptr_t>(a.dZ)) & 15u) == 0, "attn_bwd: frame tensors must be 16-byte aligned");
  SDUMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.dOut) | reinterpret_cast<uintptr_t>(a.Qp) | reinterpret_cast<uintptr_t>(a.O_pre)) & 15u) == 0 &&
                  a.dout_stride_b % 4 == 0 && a.qp_stride_b % 4 == 0,
                  "attn_bwd: dOut / Qp / O_pre must be 16-byte aligned with strides that are multiples of 4");
  const size_t smem = attn_bwd_smem(a.L, G);
  // + 16.5 KB (G = 256: dP exchange x2, barriers) / 8.3 KB static stays under the 227 KB per-CTA limit
  const size_t kMaxDyn = (G == 256 ? 210 : 218) * 1024;
  SDUMC_CHECK_ARG(smem <= kMaxDyn, "attn_bwd: L=%d too long for the shared-memory probability cache", a.L);
  static bool attr_done[kMaxDevices] = {false};
  const int dev = current_device();
  if (!attr_done[dev]) {
    SDUMC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    SDUMC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<7, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    SDUMC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024));
    SDUMC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<7, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024));
    attr_done[dev] = true;
  }
  const int rows_per_stage = G == 256 ? FrameCfg<256>::kRows : FrameCfg<1024>::kRows;
  const long n_units = (long)a.B * ((a.L + rows_per_stage - 1) / rows_per_stage);
  CUtensorMap tx, tk, tz, th;     // [B, L, G] views: boxes of 64 columns x one stage of rows of one sample
  SDUMC_TRY(get_frame_tmap(a.X, G, a.L, a.B, rows_per_stage, &tx));
  SDUMC_TRY(get_frame_tmap(a.Kt, G, a.L, a.B, rows_per_stage, &tk));
  SDUMC_TRY(get_frame_tmap(a.dZ, G, a.L, a.B, rows_per_stage, &tz));
  SDUMC_TRY(get_frame_tmap(a.dH, G, a.L, a.B, rows_per_stage, &th));
  int grid = (int)std::min<long>(n_units, num_sms());   // one resident CTA per SM (shared-memory bound)
  if (a.max_ctas > 0) grid = std::min(grid, a.max_ctas);
  if (G == 256) {
    if (a.nq == 1) SDUMC_CUDA(launch_kernel(attn_bwd_kernel<1, 256>, dim3(grid), dim3(kBwdThreads + 32), smem, stream, 1, tx, tk, tz, th, a));
    else           SDUMC_CUDA(launch_kernel(attn_bwd_kernel<7, 256>, dim3(grid), dim3(kBwdThreads + 32), smem, stream, 1, tx, tk, tz, th, a));
  } else {
    if (a.nq == 1) SDUMC_CUDA(launch_kernel(attn_bwd_kernel<1, 1024>, dim3(grid), dim3(kBwdThreads + 32), smem, stream, 1, tx, tk, tz, th, a));
    else           SDUMC_CUDA(launch_kernel(attn_bwd_kernel<7, 1024>, dim3(grid), dim3(kBwdThreads + 32), smem, stream, 1, tx, tk, tz, th, a));
  }
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// small streaming helpers
// ------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
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
  SDUMC_CUDA(launch_kernel(cast_f32_bf16_kernel, dim3((unsigned)((nthreads + 255) / 256)), dim3(256), 0, stream, 1, src, dst, n));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// collate of the device-resident feature store: gather + right-zero-pad (read_data.py:223-248).
// One warp per output row, 16-byte accesses; HBM-bound copy (2 * D bytes read + 2 * D written per row).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) collate_pad_kernel(const __nv_bfloat16* __restrict__ packed,
                                                           const long long* __restrict__ row_offset,
                                                           const int* __restrict__ idx, int b, int Lpad, int D,
                                                           __nv_bfloat16* __restrict__ out,
                                                           const int* __restrict__ out_off) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long)b * Lpad) return;
  const int i = (int)(row / Lpad), l = (int)(row - (long)i * Lpad);
  const int u = __ldg(idx + i);
  const long long r0 = __ldg(row_offset + u), r1 = __ldg(row_offset + u + 1);
  const bool live = l < (int)(r1 - r0);
  if (out_off && !live) return;                       // packed output: padding rows do not exist
  const uint4* src = reinterpret_cast<const uint4*>(packed + (r0 + l) * (long long)D);
  uint4* dst = reinterpret_cast<uint4*>(out + (out_off ? (long)(__ldg(out_off + i) + l) : row) * D);
  for (int c = threadIdx.x & 31; c < D / 8; c += 32) dst[c] = live ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
}
int launch_collate_pad(const __nv_bfloat16* packed, const long long* row_offset, const int* idx, int b, int Lpad, int D,
                       __nv_bfloat16* out, const int* out_off, cudaStream_t stream) {
  SDUMC_CHECK_ARG(packed && row_offset && idx && out, "collate_pad: null pointer");
  SDUMC_CHECK_ARG(b > 0 && Lpad > 0 && D > 0 && D % 8 == 0, "collate_pad: bad shape b=%d Lpad=%d D=%d", b, Lpad, D);
  SDUMC_CHECK_ARG(((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                  "collate_pad: pointers must be 16-byte aligned");
  const long rows = (long)b * Lpad;
  SDUMC_CUDA(launch_kernel(collate_pad_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, stream, 1, packed, row_offset, idx, b, Lpad, D, out, out_off));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

// column sums of a bf16 matrix [rows, cols] -> atomicAdd into out[cols]   (bias gradients: the in-projections',
// and those of the MLP layers whose dZ comes straight out of a GEMM epilogue); cols % 8 == 0, cols <= 1024
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long ld, long rows, int cols,
                                                           float* __restrict__ out) {
  pdl_wait();                // predecessors complete + visible (see common.cuh: PDL)
  pdl_launch_dependents();
  __shared__ float red[1024];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < cols; i += 256) red[i] = 0.f;
  __syncthreads();
  for (int cb = 0; cb < cols; cb += 256) {
    const int c = cb + lane * 8;
    if (c >= cols) continue;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long r = (long)blockIdx.x * 8 + warp; r < rows; r += (long)gridDim.x * 8) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(X + r * ld + c)), x);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += x[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red[c + j], acc[j]);
  }
  __syncthreads();
  for (int i = tid; i < cols; i += 256) atomicAdd(out + i, red[i]);
}
int launch_colsum_bf16(const __nv_bfloat16* X, long ld, long rows, int cols, float* out, cudaStream_t stream) {
  SDUMC_CHECK_ARG(X && out && rows > 0 && ld % 8 == 0 && cols > 0 && cols % 8 == 0 && cols <= 1024 &&
                      (reinterpret_cast<uintptr_t>(X) & 15u) == 0,
                  "colsum_bf16: bad arguments (cols %% 8 == 0, cols <= 1024, ld %% 8 == 0, 16-byte aligned)");
  long blocks = (rows + 63) / 64;
  if (blocks > (long)num_sms() * 8) blocks = (long)num_sms() * 8;
  SDUMC_CUDA(launch_kernel(colsum_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, 1, X, ld, rows, cols, out));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sdumc
