// sdumc_b200 — low-latency GEMM for the utterance-level chain.
//
// The utterance chain of the model (modality MLPs, gate, query MLPs, cross MLPs, heads: reference
// toolkit/models/wengnet_mosei_mult_views_text_missing.py:293-368 and their autograd) is ~60 dependent
// [R x 256] x [256 x 256]-sized products per step (R = 2B or 14B rows).  They are latency-bound: the TMA /
// tcgen05 pipeline of gemm.cuh has ~5 us of fixed cost per launch (tensor-map fetch, TMEM allocation, three
// barrier hand-offs before the first store), which the CUPTI timeline of the step showed as ~0.8 ms of
// nearly idle GPU.  This kernel serves the same operator (same epilogue semantics, same operand rounding)
// with the warp-level tensor-core path: cp.async double buffering straight into padded shared memory,
// mma.sync (tf32 m16n8k8 / bf16 m16n8k16), epilogue from registers.  64 x 64 output tile per CTA, 4 warps.
//
//   C[M,N] = epilogue(A[M,K] * op(B)),  A row-major (K contiguous),
//   B stored [N,K] (nn.Linear weight: forward, tf32 or bf16) or [K,N] (dX = dZ W: bf16).
#include <type_traits>

#include "gemm.cuh"

namespace sdumc {

namespace {

constexpr int kTM = 64, kTN = 64, kTK = 32;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t sz = pred ? 16u : 0u;   // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kN>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

// D[16x8] += A[16x8] * B[8x8], tf32 operands (the low 13 mantissa bits of the fp32 registers are ignored, like
// tcgen05 kind::tf32 reading fp32 shared memory), fp32 accumulators
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool kTF32, bool kBMN>
struct Smem {
  // pitches chosen so that every fragment access below is bank-conflict free
  static constexpr int kAP = kTF32 ? 36 : 40;                    // elements per A row (32 + pad)
  static constexpr int kBP = kBMN ? 72 : (kTF32 ? 36 : 40);      // per B row: [n][k] rows of 32, or [k][n] rows of 64
  static constexpr int kElem = kTF32 ? 4 : 2;
  static constexpr int kABytes = kTM * kAP * kElem;
  static constexpr int kBBytes = (kBMN ? kTK : kTN) * kBP * kElem;
};

template <bool kTF32, bool kBMN>
__global__ void __launch_bounds__(128) small_gemm_kernel(const void* __restrict__ Ap, long lda, const void* __restrict__ Bp,
                                                         long ldb, int M, int N, int K, const GemmEpi ep) {
  using S = Smem<kTF32, kBMN>;
  using T = typename std::conditional<kTF32, float, __nv_bfloat16>::type;
  __shared__ __align__(16) unsigned char sA[2][S::kABytes];
  __shared__ __align__(16) unsigned char sB[2][S::kBBytes];
  const T* A = static_cast<const T*>(Ap);
  const T* B = static_cast<const T*>(Bp);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;          // this warp's 32 x 32 sub-tile
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * kTM, n0 = blockIdx.y * kTN;
  constexpr int kVec = 16 / S::kElem;                              // elements per 16-byte copy

  auto load_chunk = [&](int buf, int k0) {
    T* as = reinterpret_cast<T*>(sA[buf]);
    T* bs = reinterpret_cast<T*>(sB[buf]);
    constexpr int kAVecRow = kTK / kVec;                           // 16-byte copies per A row
    for (int x = tid; x < kTM * kAVecRow; x += 128) {
      const int r = x / kAVecRow, c = (x % kAVecRow) * kVec;
      cp_async16(as + r * S::kAP + c, A + (long)min(m0 + r, M - 1) * lda + k0 + c, m0 + r < M);
    }
    if constexpr (!kBMN) {                                         // B[n][k]
      for (int x = tid; x < kTN * kAVecRow; x += 128) {
        const int r = x / kAVecRow, c = (x % kAVecRow) * kVec;
        cp_async16(bs + r * S::kBP + c, B + (long)min(n0 + r, N - 1) * ldb + k0 + c, n0 + r < N);
      }
    } else {                                                       // B[k][n]
      constexpr int kBVecRow = kTN / kVec;
      for (int x = tid; x < kTK * kBVecRow; x += 128) {
        const int r = x / kBVecRow, c = (x % kBVecRow) * kVec;
        cp_async16(bs + r * S::kBP + c, B + (long)(k0 + r) * ldb + min(n0 + c, N - 8), n0 + c < N);   // N % 8 == 0 (host)
      }
    }
    cp_async_commit();
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  const int nk = K / kTK;                                          // K % 32 == 0 (host)
  load_chunk(0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    if (kc + 1 < nk) {
      load_chunk((kc + 1) & 1, (kc + 1) * kTK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const T* as = reinterpret_cast<const T*>(sA[kc & 1]);
    const T* bs = reinterpret_cast<const T*>(sB[kc & 1]);
    if constexpr (kTF32) {
      const uint32_t* a32 = reinterpret_cast<const uint32_t*>(as);
      const uint32_t* b32 = reinterpret_cast<const uint32_t*>(bs);
#pragma unroll
      for (int ks = 0; ks < kTK; ks += 8) {
        uint32_t af[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t* p = a32 + (wm + i * 16 + g) * S::kAP + ks + t;
          af[i][0] = p[0]; af[i][1] = p[8 * S::kAP]; af[i][2] = p[4]; af[i][3] = p[8 * S::kAP + 4];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t* q = b32 + (wn + j * 8 + g) * S::kBP + ks + t;
          const uint32_t b0 = q[0], b1 = q[4];
          mma_tf32(acc[0][j], af[0], b0, b1);
          mma_tf32(acc[1][j], af[1], b0, b1);
        }
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < kTK; ks += 16) {
        uint32_t af[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const __nv_bfloat16* p = as + (wm + i * 16 + g) * S::kAP + ks + 2 * t;
          af[i][0] = *reinterpret_cast<const uint32_t*>(p);
          af[i][1] = *reinterpret_cast<const uint32_t*>(p + 8 * S::kAP);
          af[i][2] = *reinterpret_cast<const uint32_t*>(p + 8);
          af[i][3] = *reinterpret_cast<const uint32_t*>(p + 8 * S::kAP + 8);
        }
        if constexpr (!kBMN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat16* q = bs + (wn + j * 8 + g) * S::kBP + ks + 2 * t;
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(q), b1 = *reinterpret_cast<const uint32_t*>(q + 8);
            mma_16816(acc[0][j], af[0], b0, b1);
            mma_16816(acc[1][j], af[1], b0, b1);
          }
        } else {
          // B[k][n]: transposing ldmatrix; one x4 covers k 0..15 of two 8-column tiles
#pragma unroll
          for (int jp = 0; jp < 2; ++jp) {
            uint32_t bf[4];
            const __nv_bfloat16* q = bs + (ks + (lane & 7) + ((lane >> 3) & 1) * 8) * S::kBP + wn + jp * 16 + (lane >> 4) * 8;
            ldsm_x4_trans(bf, q);
            mma_16816(acc[0][2 * jp], af[0], bf[0], bf[1]);
            mma_16816(acc[1][2 * jp], af[1], bf[0], bf[1]);
            mma_16816(acc[0][2 * jp + 1], af[0], bf[2], bf[3]);
            mma_16816(acc[1][2 * jp + 1], af[1], bf[2], bf[3]);
          }
        }
      }
    }
    __syncthreads();                                               // the buffer is refilled two iterations later
  }

  // ---- epilogue: bias -> activation -> ReLU/dropout gate -> element dropout -> fp32 / bf16 outputs (gemm.cuh order) ----
  const uint32_t drop_thr = drop_threshold(ep.drop_p);
  const float drop_scale = ep.drop_p > 0.f ? 1.f / (1.f - ep.drop_p) : 1.f;
  const DropKey key = resolve_key(ep.key);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {                                  // accumulator rows g / g + 8
      const int r = m0 + wm + i * 16 + g + h * 8;
      if (r >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn + j * 8 + 2 * t;
        if (n >= N) continue;                                      // N is even: the pair (n, n + 1) is in or out together
        float v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
        if (ep.bias) { v0 += __ldg(ep.bias + n); v1 += __ldg(ep.bias + n + 1); }
        if (ep.act == ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        else if (ep.act == ACT_TANH) { v0 = tanh_fast(v0); v1 = tanh_fast(v1); }
        if (ep.gate) {
          const float2 gt = *reinterpret_cast<const float2*>(ep.gate + (long)r * ep.ld_gate + n);
          v0 = gt.x > 0.f ? v0 * ep.gate_scale : 0.f;
          v1 = gt.y > 0.f ? v1 * ep.gate_scale : 0.f;
        }
        if (ep.drop_p > 0.f) {
          // four consecutive elements share one Philox counter (e >> 2); this thread holds the lower or upper pair
          const uint32_t e = (uint32_t)r * (uint32_t)N + (uint32_t)n;
          const U4 rw = philox4x32_10(e >> 2, 0x5D0Cu, ep.drop_site, key.step, key.seed_lo, key.seed_hi);
          const uint32_t w0 = (e & 2u) ? rw.z : rw.x, w1 = (e & 2u) ? rw.w : rw.y;
          v0 = w0 >= drop_thr ? v0 * drop_scale : 0.f;
          v1 = w1 >= drop_thr ? v1 * drop_scale : 0.f;
        }
        if (ep.out_f32) {
          float* dst = ep.out_f32 + (long)r * ep.ld_f32 + n;
          if (ep.f32_mode == OUT_ATOMIC) {
            asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(dst), "f"(v0), "f"(v1) : "memory");
          } else {
            float2 o = make_float2(v0, v1);
            if (ep.f32_mode == OUT_ADD) {
              const float2 old = *reinterpret_cast<const float2*>(dst);
              o.x += old.x; o.y += old.y;
            }
            *reinterpret_cast<float2*>(dst) = o;
          }
        }
        if (ep.out_bf16) {
          uint32_t* dst = reinterpret_cast<uint32_t*>(ep.out_bf16 + (long)r * ep.ld_bf16 + n);
          uint32_t w = pack_bf16x2(v0, v1);
          if (ep.bf16_mode == OUT_ADD) w = add_bf16x2(w, *dst);
          *dst = w;
        }
      }
    }
  }
}

}  // namespace

bool small_gemm_eligible(const GemmShape& sh, const GemmEpi& ep, bool tf32, long lda, long ldb) {
  const int elem = tf32 ? 4 : 2;
  if (ep.kind != EPI_GENERIC || sh.a_mn || sh.k_splits > 1 || ep.fmask_site != 0) return false;
  if (tf32 && sh.b_mn) return false;
  if (sh.M > 16384 || sh.K % kTK != 0 || sh.N % 8 != 0) return false;
  if ((lda * elem) % 16 != 0 || (ldb * elem) % 16 != 0) return false;
  if (ep.drop_p > 0.f && sh.N % 4 != 0) return false;
  if (ep.out_f32 && (ep.ld_f32 % 2 != 0 || (reinterpret_cast<uintptr_t>(ep.out_f32) & 7u) != 0)) return false;
  if (ep.out_bf16 && (ep.ld_bf16 % 2 != 0 || (reinterpret_cast<uintptr_t>(ep.out_bf16) & 3u) != 0)) return false;
  if (ep.gate && (ep.ld_gate % 2 != 0 || (reinterpret_cast<uintptr_t>(ep.gate) & 7u) != 0)) return false;
  return true;
}

int launch_small_gemm(const GemmOperand& A, const GemmOperand& B, const GemmShape& sh, const GemmEpi& ep, bool tf32,
                      cudaStream_t stream) {
  SDUMC_CHECK_ARG(A.ptr && B.ptr && (reinterpret_cast<uintptr_t>(A.ptr) & 15u) == 0 && (reinterpret_cast<uintptr_t>(B.ptr) & 15u) == 0,
                  "gemm: operands must be 16-byte aligned");
  const dim3 grid((sh.M + kTM - 1) / kTM, (sh.N + kTN - 1) / kTN);
  if (tf32)          small_gemm_kernel<true, false><<<grid, 128, 0, stream>>>(A.ptr, A.ld, B.ptr, B.ld, sh.M, sh.N, sh.K, ep);
  else if (sh.b_mn)  small_gemm_kernel<false, true><<<grid, 128, 0, stream>>>(A.ptr, A.ld, B.ptr, B.ld, sh.M, sh.N, sh.K, ep);
  else               small_gemm_kernel<false, false><<<grid, 128, 0, stream>>>(A.ptr, A.ld, B.ptr, B.ld, sh.M, sh.N, sh.K, ep);
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sdumc
