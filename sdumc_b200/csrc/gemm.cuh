// sdumc_b200 — the one tcgen05 GEMM of the library (sm_100a).
//
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N]      fp32 accumulation in TMEM
//
// * operands bf16 (kind::f16) or fp32 read as tf32 (kind::tf32), staged in shared memory by TMA
//   (SWIZZLE_128B boxes) through a kStages-deep mbarrier ring;
// * either operand may be K-major (reduction index contiguous: the nn.Linear forward case) or
//   MN-major (reduction index strided: the dX = dY*W and dW = dY^T*X cases) — only the TMA box
//   and the UMMA descriptor differ;
// * persistent CTAs, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
//   warp 2 = TMEM allocator, warps 4..7 / 8..11 = two epilogue groups (TMEM -> registers -> 256-bit
//   row stores, software-pipelined over 32-column chunks) that alternate tiles; accumulators double-buffered in
//   TMEM (2 x kBlockN columns) so the epilogue of tile i overlaps the MMAs of tiles i+1, i+2;
// * split-K over the reduction for the weight-gradient shapes (fp32 atomics into the grad buffer).
//
// Epilogues (fused, selected at run time):
//   generic : bias, ReLU/tanh, ReLU+dropout backward gate, element dropout, frame-mask multiply,
//             fp32 store / += / atomicAdd, bf16 store / += (old values prefetched one chunk ahead)
//   in-proj : bias, plain bf16 H plus up to 4 independently dropped copies (the X' of
//             FRA2UTT_new / Cross_Attention of both passes; reference
//             toolkit/models/wengnet_mosei_mult_views_text_missing.py:57,81)
//   key-proj: bias + tanh + per-sample query dot products -> attention scores
//             (reference :60-61 and :82-88), optional bf16 K for the backward pass
#pragma once

#include "common.cuh"

namespace sdumc {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };
enum : int { OUT_STORE = 0, OUT_ADD = 1, OUT_ATOMIC = 2 };
enum : int { EPI_GENERIC = 0, EPI_INPROJ = 1, EPI_KEYPROJ = 2 };
// compile-time specialisations of the epilogue (one compact chunk loop per kind: the all-in-one body thrashed
// the instruction cache); KIND_RMW = generic restricted to 'bf16 += acc * frame_mask' (the dH accumulation)
enum : int { KIND_GENERIC = 0, KIND_INPROJ = 1, KIND_KEYPROJ = 2, KIND_RMW = 3 };

struct GemmEpi {
  int kind;  // EPI_*
  const float* bias;
  int act;
  const float* gate;  // fp32 [M, N] (ld_gate): v *= gate_scale * (gate > 0)
  long ld_gate;
  float gate_scale;
  float drop_p;  // element dropout after activation, index e = r * N + n
  uint32_t drop_site;
  uint32_t fmask_site;  // != 0: v *= 2 * framebit(site, r, n)
  uint32_t fmask_site2; // rows >= fmask_split (> 0): framebit(site2, r - fmask_split, n)
  long fmask_split;
  float* out_f32;
  long ld_f32;
  int f32_mode;
  __nv_bfloat16* out_bf16;
  long ld_bf16;
  int bf16_mode;
  // EPI_INPROJ
  int n_tgt;
  __nv_bfloat16* tgt[4];
  uint32_t tgt_site[4];
  // EPI_KEYPROJ
  const float* qv;  // [n_samples, nq, N] (q_stride = nq * N) or one shared [nq, N] (q_stride = 0)
  long q_stride;
  int nq;
  int L;          // frames per sample: sample = row / L
  float* scores;  // [M, nq]
  DropKey key;
};

struct GemmShape {
  int M, N, K;
  int a_mn, b_mn;  // 0 = K-major, 1 = MN-major
  int k_splits;
  int b_res;       // 1: the whole B operand (one N tile, K <= kStages k-blocks) stays resident in shared memory
};


__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  const float lo = __uint_as_float(a << 16) + __uint_as_float(b << 16);
  const float hi = __uint_as_float(a & 0xffff0000u) + __uint_as_float(b & 0xffff0000u);
  return pack_bf16x2(lo, hi);
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a lane's 32 bf16 of one row chunk = two full 32-byte sectors
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* p, uint32_t* w) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p)
               : "memory");
}
// store (or accumulate onto prefetched old values) this lane's 32 consecutive bf16 of its own row
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, uint32_t (&w)[16], const uint32_t* oldv) {
  if (oldv) {
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = add_bf16x2(w[j], oldv[j]);
  }
  st_global_256(dst, &w[0]);
  st_global_256(dst + 16, &w[8]);
}

template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

template <int kBlockN>
struct GemmCfg {
  static constexpr int kBlockM = 128;
  static constexpr int kRowBytes = 128;                      // one swizzle row
  static constexpr int kABytes = kBlockM * kRowBytes;        // 16 KB
  static constexpr int kBBytes = kBlockN * kRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kBlockN == 256) ? 4 : (kBlockN == 128 ? 6 : 8);
  static constexpr int kMaxStages = 8;                       // barrier slots (the resident-B ring may be deeper than kStages)
  // ring area: kStages full (A + B) stages; for N = 256 one extra A tile, so that a resident 256 x 256 weight (128 KB)
  // leaves a 5-deep A ring: the K = 256 frame GEMMs are bound by DRAM latency x ring depth (one 128-row tile = 4
  // k-blocks = the whole 4-stage ring: the loads of tile i+1 could only start as tile i's MMAs retired)
  static constexpr int kRingBytes = kStages * kStageBytes + (kBlockN == 256 ? kABytes : 0);
  static constexpr int kTmemCols = (2 * kBlockN < 32) ? 32 : 2 * kBlockN;
  static constexpr int kEpiWarps = 8;                        // two groups of 4, alternating tiles
  static constexpr int kVecBytes = 2 /*groups*/ * 2 /*bias, ctx*/ * kBlockN * 4;
  static constexpr int kSmemBytes = kRingBytes + kVecBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kThreads = 384;
};

// kPair: the kernel runs as clusters of two CTAs on one TPC (cta_group::2).  One UMMA of M = 256 covers the pair's two
// 128-row tiles; each CTA stages its own A tile but only HALF of the B tile (split along N), so a B stage costs half
// the shared memory and the ring (or the A-only ring behind a resident weight) gets deeper: 6 instead of 4 full stages,
// 8 instead of 5 A tiles behind a resident 256 x 256 weight - the frame-level GEMMs are bound by DRAM latency x ring
// depth, not by the MMA.  The leader CTA (cluster rank 0) issues every MMA; TMA loads of both CTAs count their bytes on
// the leader's barriers; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; the epilogue warps
// of both CTAs report "accumulator drained" to the leader.  B is also fetched once per pair instead of once per CTA.
template <int kBlockN, bool kTF32, int kKind, bool kPair = false>
__global__ void __launch_bounds__(384, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmShape sh, const GemmEpi ep) {
  using Cfg = GemmCfg<kBlockN>;
  static_assert(!kPair || (!kTF32 && kBlockN == 256), "CTA pairs: bf16, BLOCK_N = 256");
  constexpr int kBCta = kPair ? Cfg::kBBytes / 2 : Cfg::kBBytes;      // bytes of B per k-block held by one CTA
  constexpr int kNCta = kPair ? kBlockN / 2 : kBlockN;               // columns of the N tile staged by one CTA
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  constexpr int kElem = kTF32 ? 4 : 2;
  constexpr int kBlockK = 128 / kElem;   // elements (K-major) == k rows (MN-major) per stage
  constexpr int kUmmaK = 32 / kElem;     // 16 (bf16) / 8 (tf32)
  constexpr int kPanel = 128 / kElem;    // MN elements per 128-byte swizzle row
  constexpr uint32_t kFmt = kTF32 ? 2u : 1u;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  float* vec_s = reinterpret_cast<float*>(smem + Cfg::kRingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes + Cfg::kVecBytes);
  uint64_t* full_bar = bars;                          // [kMaxStages]
  uint64_t* empty_bar = bars + Cfg::kMaxStages;       // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kMaxStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* bres_bar = tempty_bar + 3;             // resident-B mode: B loaded once per CTA
  // resident-B mode (frame-level GEMMs with a 256x256 weight): B occupies the first nkb * kBBytes of the stage
  // area, the ring behind it carries A only.  Without it every 128-row tile re-fetches the weight from L2
  // (2x the A traffic) and the L2->SM fabric, not HBM, bounds the main loop.
  const bool b_res = sh.b_res != 0;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // M tiles per scheduling unit: a pair works on two neighbouring M tiles (CTA rank r takes tile 2 * unit + r; with
  // an odd tile count the last pair's second tile is out of range: TMA zero-fills it, the epilogue stores nothing)
  const int m_tiles = kPair ? ((sh.M + 127) / 128 + 1) / 2 : (sh.M + 127) / 128;
  const int n_tiles = (sh.N + kBlockN - 1) / kBlockN;
  const int nkb = (sh.K + kBlockK - 1) / kBlockK;
  const int num_tiles = m_tiles * n_tiles * sh.k_splits;
  const int cta0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // scheduling unit index / count
  const int nctas = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // ring depth: full stages, or - with B resident - as many A tiles as fit behind it
  const int nst = b_res ? min(Cfg::kMaxStages, (Cfg::kRingBytes - nkb * kBCta) / Cfg::kABytes)
                        : min(Cfg::kMaxStages, Cfg::kRingBytes / (Cfg::kABytes + kBCta));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kPair ? 8 : 4);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    mbar_init(bres_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kPair) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (descriptor prefetch, barrier init, TMEM allocation) may overlap the tail of the previous
  // kernel; nothing below touches global memory before the previous grid is complete
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) reg_dealloc<40>();   // warpgroup 0 (producer, MMA issuer, TMEM allocator, spare) keeps 40 registers
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint8_t* ring_base = stage_base + (b_res ? nkb * kBCta : 0);
      const int stage_stride = b_res ? Cfg::kABytes : Cfg::kABytes + kBCta;
      // loads of a pair count on the leader's barrier; only the leader posts the expected byte count (of both CTAs)
      auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
        if constexpr (kPair) tma_load_2d_pair(dst, tm, c0, c1, bar);
        else tma_load_2d(dst, tm, c0, c1, bar);
      };
      constexpr uint32_t kTxMul = kPair ? 2u : 1u;
      const int n_half = (int)crank * kNCta;            // first column of this CTA's share of the N tile
      if (b_res && cta0 < num_tiles) {
        if (leader) mbar_expect_tx(bres_bar, kTxMul * (uint32_t)(nkb * kBCta));
        for (int kb = 0; kb < nkb; ++kb) {
          uint8_t* sb = stage_base + kb * kBCta;
          if (!sh.b_mn) {
            load(sb, &tmB, kb * kBlockK, n_half, bres_bar);
          } else {
#pragma unroll
            for (int p = 0; p < kNCta / kPanel; ++p)
              load(sb + p * (kBlockK * 128), &tmB, n_half + p * kPanel, kb * kBlockK, bres_bar);
          }
        }
      }
      for (int t = cta0; t < num_tiles; t += nctas) {
        const int ks = t / (m_tiles * n_tiles);
        const int mn = t - ks * (m_tiles * n_tiles);
        const int mu = mn / n_tiles, n_blk = mn - mu * n_tiles;
        const int m_blk = kPair ? 2 * mu + (int)crank : mu;
        const int kb0 = (int)(((long)ks * nkb) / sh.k_splits);
        const int kb1 = (int)(((long)(ks + 1) * nkb) / sh.k_splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = ring_base + stage * stage_stride;
          uint8_t* sb = sa + Cfg::kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], kTxMul * (uint32_t)(b_res ? Cfg::kABytes : Cfg::kABytes + kBCta));
          if (!sh.a_mn) {
            load(sa, &tmA, kb * kBlockK, m_blk * 128, &full_bar[stage]);
          } else {
#pragma unroll
            for (int p = 0; p < 128 / kPanel; ++p)
              load(sa + p * (kBlockK * 128), &tmA, m_blk * 128 + p * kPanel, kb * kBlockK, &full_bar[stage]);
          }
          if (b_res) {
            // B is resident
          } else if (!sh.b_mn) {
            load(sb, &tmB, kb * kBlockK, n_blk * kBlockN + n_half, &full_bar[stage]);
          } else {
#pragma unroll
            for (int p = 0; p < kNCta / kPanel; ++p)
              load(sb + p * (kBlockK * 128), &tmB, n_blk * kBlockN + n_half + p * kPanel, kb * kBlockK, &full_bar[stage]);
          }
          if (++stage == nst) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (one thread; the leader CTA of a pair) =====================
    const uint32_t idesc = make_idesc(kFmt, (uint32_t)sh.a_mn, (uint32_t)sh.b_mn, (uint32_t)kBlockN, kPair ? 256u : 128u);
    const uint32_t mn_lbo = (uint32_t)(kBlockK * 128);
    const uint32_t mn_sbo = 1024u;
    const uint32_t a_lbo = sh.a_mn ? mn_lbo : 16u, a_sbo = sh.a_mn ? mn_sbo : 1024u;
    const uint32_t b_lbo = sh.b_mn ? mn_lbo : 16u, b_sbo = sh.b_mn ? mn_sbo : 1024u;
    // descriptor start-address advance per UMMA_K step, in 16-byte units
    const uint32_t a_kstep = sh.a_mn ? (uint32_t)(kUmmaK * 128) >> 4 : 32u >> 4;
    const uint32_t b_kstep = sh.b_mn ? (uint32_t)(kUmmaK * 128) >> 4 : 32u >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t ring_u32 = smem_u32(stage_base) + (b_res ? (uint32_t)(nkb * kBCta) : 0u);
    const uint32_t stage_stride = b_res ? (uint32_t)Cfg::kABytes : (uint32_t)(Cfg::kABytes + kBCta);
    auto commit = [&](uint64_t* bar) {
      if constexpr (kPair) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    if (b_res && cta0 < num_tiles) mbar_wait(bres_bar, 0u);
    for (int t = cta0; t < num_tiles; t += nctas) {
      const int ks = t / (m_tiles * n_tiles);
      const int kb0 = (int)(((long)ks * nkb) / sh.k_splits);
      const int kb1 = (int)(((long)(ks + 1) * nkb) / sh.k_splits);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kBlockN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ring_u32 + (uint32_t)stage * stage_stride;
          const uint32_t sb = b_res ? smem_u32(stage_base) + (uint32_t)(kb * kBCta) : sa + Cfg::kABytes;
          const uint64_t da = make_smem_desc(sa, a_lbo, a_sbo);
          const uint64_t db = make_smem_desc(sb, b_lbo, b_sbo);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            if constexpr (kPair)
              umma_ss_pair(tmem_d, da + (uint64_t)(k * a_kstep), db + (uint64_t)(k * b_kstep), idesc,
                           (kb > kb0 || k > 0) ? 1u : 0u);
            else
              umma_ss<kTF32>(tmem_d, da + (uint64_t)(k * a_kstep), db + (uint64_t)(k * b_kstep), idesc,
                             (kb > kb0 || k > 0) ? 1u : 0u);
          }
          commit(&empty_bar[stage]);                 // smem slot free (in both CTAs of a pair) once these MMAs retire
          if (kb == kb1 - 1) commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == nst) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: two groups of 4 warps, alternating tiles =====================
    reg_alloc<224>();                 // registers released by the producer / MMA warpgroup
    const int ew = warp & 3;          // TMEM lane quarter this warp may touch
    const int grp = (warp - 4) >> 2;  // group g drains accumulator stage g (tiles with local index % 2 == g)
    const int gtid = threadIdx.x - 128 - grp * 128;
    float* bias_s = vec_s + grp * 2 * kBlockN;       // this tile's bias slice
    float* ctx_s = bias_s + kBlockN;                 // shared query vector (key-projection, nq == 1)
    uint32_t acc_phase = 0;
    const uint32_t drop_thr = drop_threshold(ep.drop_p);
    const float drop_scale = ep.drop_p > 0.f ? 1.f / (1.f - ep.drop_p) : 1.f;
    const DropKey key = resolve_key(ep.key);
    const bool rmw = kKind == KIND_RMW || (kKind == KIND_GENERIC && ep.out_bf16 && ep.bf16_mode == OUT_ADD);
    const bool ctx_shared = kKind == KIND_KEYPROJ && ep.q_stride == 0 && ep.nq == 1;
    constexpr int NC = kBlockN / 32;
    // per-column vectors (bias, shared query) of an N tile -> this group's shared-memory copy (group-local named
    // barrier, 128 threads).  With a single N tile (every frame-level GEMM) they are staged ONCE per CTA: the global
    // load + barrier used to sit on every tile's epilogue, which paces these kernels (profiles/r2_experiments.md).
    auto stage_vecs = [&](int n_blk) {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // the previous tile's readers are done
      for (int j = gtid; j < kBlockN; j += 128) {
        const int n = n_blk * kBlockN + j;
        bias_s[j] = (ep.bias && n < sh.N) ? __ldg(ep.bias + n) : 0.f;
        if (ctx_shared) ctx_s[j] = n < sh.N ? __ldg(ep.qv + n) : 0.f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    };
    if (n_tiles == 1) stage_vecs(0);
    int li = 0;
    for (int t = cta0; t < num_tiles; t += nctas, ++li) {
      if ((li & 1) != grp) continue;
      const int acc = grp;
      const int mn = t % (m_tiles * n_tiles);
      const int mu = mn / n_tiles, n_blk = mn - mu * n_tiles;
      const int m_blk = kPair ? 2 * mu + (int)crank : mu;
      if (n_tiles > 1) stage_vecs(n_blk);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int r0 = m_blk * 128 + ew * 32;  // first row of this warp's slab
      const int r = r0 + lane;
      const bool row_ok = r < sh.M;
      const bool f32_vec = ((reinterpret_cast<uintptr_t>(ep.out_f32) | (uintptr_t)(ep.ld_f32 * 4)) & 15u) == 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * kBlockN);

      float sc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* qrow = ep.qv;
      if (kKind == KIND_KEYPROJ && row_ok) qrow = ep.qv + (long)(r / ep.L) * ep.q_stride;
      U4 tw[4];

      // issue the TMEM load of chunk c (and, for read-modify-write outputs, the loads of the old values)
      auto issue = [&](int c, uint32_t (&buf)[32], uint32_t (&oldv)[16]) {
        tmem_ld32(taddr + (uint32_t)(c * 32), buf);
        if (rmw) {
          const int n0 = n_blk * kBlockN + c * 32;
          if (row_ok && n0 < sh.N) {
            const __nv_bfloat16* src = ep.out_bf16 + (long)r * ep.ld_bf16 + n0;
            ld_global_256(src, &oldv[0]);
            ld_global_256(src + 16, &oldv[8]);
          }
        }
      };

      auto process = [&](int c, const uint32_t (&acc_r)[32], const uint32_t (&oldv)[16]) {
        const int n0 = n_blk * kBlockN + c * 32;
        if (n0 >= sh.N) return;  // warp-uniform
        const bool full = (n0 + 32 <= sh.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
          v[j] = __uint_as_float(acc_r[j]) + b4.x;
          v[j + 1] = __uint_as_float(acc_r[j + 1]) + b4.y;
          v[j + 2] = __uint_as_float(acc_r[j + 2]) + b4.z;
          v[j + 3] = __uint_as_float(acc_r[j + 3]) + b4.w;
        }
        if constexpr (kKind == KIND_KEYPROJ) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tanh_fast(v[j]);
        } else if constexpr (kKind == KIND_GENERIC) {
          if (ep.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (ep.act == ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tanh_fast(v[j]);
          }
        }

        if constexpr (kKind == KIND_INPROJ) {
          uint32_t w[16];
          if (ep.out_bf16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            if (row_ok) store_row_bf16(ep.out_bf16 + (long)r * ep.ld_bf16 + n0, w, nullptr);
          }
          // dropped copies: word (n0 >> 5) & 3 of Philox(row, n0 >> 7, site, step)
          if ((c & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < ep.n_tgt) tw[i] = frame_mask_words(key, ep.tgt_site[i], (uint32_t)r, (uint32_t)(n0 >> 7));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i >= ep.n_tgt) break;
            const int wsel = (n0 >> 5) & 3;
            const uint32_t bits = wsel == 0 ? tw[i].x : (wsel == 1 ? tw[i].y : (wsel == 2 ? tw[i].z : tw[i].w));
#pragma unroll
            for (int j = 0; j < 16; ++j)
              w[j] = pack_bf16x2(((bits >> (2 * j)) & 1u) ? 2.f * v[2 * j] : 0.f,
                                 ((bits >> (2 * j + 1)) & 1u) ? 2.f * v[2 * j + 1] : 0.f);
            if (row_ok) store_row_bf16(ep.tgt[i] + (long)r * ep.ld_bf16 + n0, w, nullptr);
          }
          return;
        }

        if constexpr (kKind == KIND_KEYPROJ) {
          if (ep.out_bf16) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            if (row_ok) store_row_bf16(ep.out_bf16 + (long)r * ep.ld_bf16 + n0, w, nullptr);
            // the backward pass reads the bf16 K; score with the same rounded values
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[2 * j] = __uint_as_float(w[j] << 16);
              v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
            }
          }
          if (ctx_shared) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(ctx_s + c * 32 + j);
              s0 = fmaf(v[j], w4.x, s0);
              s1 = fmaf(v[j + 1], w4.y, s1);
              s2 = fmaf(v[j + 2], w4.z, s2);
              s3 = fmaf(v[j + 3], w4.w, s3);
            }
            sc[0] += (s0 + s1) + (s2 + s3);
          } else if (ep.nq == 7) {
            // column groups outer, queries inner: 7 independent accumulators, no serial FMA chain per query
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 w4[7];
#pragma unroll
              for (int q = 0; q < 7; ++q)
                w4[q] = __ldg(reinterpret_cast<const float4*>(qrow + (long)q * sh.N + n0) + j);
#pragma unroll
              for (int q = 0; q < 7; ++q) {
                sc[q] = fmaf(v[4 * j], w4[q].x, sc[q]);
                sc[q] = fmaf(v[4 * j + 1], w4[q].y, sc[q]);
                sc[q] = fmaf(v[4 * j + 2], w4[q].z, sc[q]);
                sc[q] = fmaf(v[4 * j + 3], w4[q].w, sc[q]);
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 7; ++q) {
              if (q >= ep.nq) break;
              const float4* qp = reinterpret_cast<const float4*>(qrow + (long)q * sh.N + n0);
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 w4 = __ldg(qp + j);
                s0 = fmaf(v[4 * j], w4.x, s0);
                s1 = fmaf(v[4 * j + 1], w4.y, s1);
                s2 = fmaf(v[4 * j + 2], w4.z, s2);
                s3 = fmaf(v[4 * j + 3], w4.w, s3);
              }
              sc[q] += (s0 + s1) + (s2 + s3);
            }
          }
          return;
        }

        if constexpr (kKind == KIND_RMW) {   // dH += acc * frame mask
          if (ep.fmask_site) {
            const bool second = ep.fmask_split > 0 && r >= ep.fmask_split;   // stacked passes: own site, own row count
            const U4 w4 = frame_mask_words(key, second ? ep.fmask_site2 : ep.fmask_site,
                                           (uint32_t)(second ? r - ep.fmask_split : r), (uint32_t)(n0 >> 7));
            const int wsel = (n0 >> 5) & 3;
            const uint32_t bits = wsel == 0 ? w4.x : (wsel == 1 ? w4.y : (wsel == 2 ? w4.z : w4.w));
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? 2.f * v[j] : 0.f;
          }
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          if (row_ok) store_row_bf16(ep.out_bf16 + (long)r * ep.ld_bf16 + n0, w, oldv);
          return;
        }

        // ---- generic ----
        if constexpr (kKind == KIND_GENERIC) {
        if (ep.gate && row_ok) {
          const float* g = ep.gate + (long)r * ep.ld_gate + n0;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + j));
              v[j] = g4.x > 0.f ? v[j] * ep.gate_scale : 0.f;
              v[j + 1] = g4.y > 0.f ? v[j + 1] * ep.gate_scale : 0.f;
              v[j + 2] = g4.z > 0.f ? v[j + 2] * ep.gate_scale : 0.f;
              v[j + 3] = g4.w > 0.f ? v[j + 3] * ep.gate_scale : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < sh.N) v[j] = (__ldg(g + j) > 0.f) ? v[j] * ep.gate_scale : 0.f;
          }
        }
        if (ep.drop_p > 0.f) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            // e = r * N + n0 + j is a multiple of 4 whenever N % 4 == 0 (checked on the host)
            const uint32_t e = (uint32_t)r * (uint32_t)sh.N + (uint32_t)(n0 + j);
            const U4 rw = philox4x32_10(e >> 2, 0x5D0Cu, ep.drop_site, key.step, key.seed_lo, key.seed_hi);
            v[j] = rw.x >= drop_thr ? v[j] * drop_scale : 0.f;
            v[j + 1] = rw.y >= drop_thr ? v[j + 1] * drop_scale : 0.f;
            v[j + 2] = rw.z >= drop_thr ? v[j + 2] * drop_scale : 0.f;
            v[j + 3] = rw.w >= drop_thr ? v[j + 3] * drop_scale : 0.f;
          }
        }
        if (ep.fmask_site) {
          const bool second = ep.fmask_split > 0 && r >= ep.fmask_split;
          const U4 w4 = frame_mask_words(key, second ? ep.fmask_site2 : ep.fmask_site,
                                         (uint32_t)(second ? r - ep.fmask_split : r), (uint32_t)(n0 >> 7));
          const int wsel = (n0 >> 5) & 3;
          const uint32_t bits = wsel == 0 ? w4.x : (wsel == 1 ? w4.y : (wsel == 2 ? w4.z : w4.w));
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? 2.f * v[j] : 0.f;
        }
        if (ep.out_f32 && row_ok) {
          float* dst = ep.out_f32 + (long)r * ep.ld_f32 + n0;
          if (ep.f32_mode == OUT_ATOMIC) {
            if (full && f32_vec) {
              // split-K reduction: 16-byte vector reds (REDG.F32x4), a quarter of the L2 atomic operations
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || n0 + j < sh.N) atomicAdd(dst + j, v[j]);
            }
          } else if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (ep.f32_mode == OUT_ADD) {
                const float4 old = *reinterpret_cast<const float4*>(dst + j);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *reinterpret_cast<float4*>(dst + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < sh.N) dst[j] = (ep.f32_mode == OUT_ADD) ? dst[j] + v[j] : v[j];
          }
        }
        if (ep.out_bf16) {  // host guarantees N % 32 == 0 for bf16 outputs
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          if (row_ok) store_row_bf16(ep.out_bf16 + (long)r * ep.ld_bf16 + n0, w, rmw ? oldv : nullptr);
        }
        }  // KIND_GENERIC
      };

      // software pipeline over the 32-column chunks: the TMEM load (and RMW prefetch) of chunk c+1 is in
      // flight while chunk c is processed
      uint32_t bufA[32], bufB[32];
      uint32_t oldA[16], oldB[16];
      issue(0, bufA, oldA);
#pragma unroll 1
      for (int c = 0; c < NC; c += 2) {
        tmem_ld_wait();
        issue(c + 1, bufB, oldB);
        process(c, bufA, oldA);
        tmem_ld_wait();
        if (c + 2 < NC) issue(c + 2, bufA, oldA);
        process(c + 1, bufB, oldB);
      }
      if (kKind == KIND_KEYPROJ && row_ok) {
        float* srow = ep.scores + (long)r * ep.nq;
#pragma unroll
        for (int q = 0; q < 7; ++q)
          if (q < ep.nq) srow[q] = sc[q];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_leader(&tempty_bar[acc]);   // the leader's MMA thread waits for both CTAs
        else mbar_arrive(&tempty_bar[acc]);
      }
      acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the peer may still signal it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// host side (gemm.cu)
struct GemmOperand {
  const void* ptr;
  long ld;  // leading dimension in elements of the stored (row-major) matrix
};
// Launch C = op(A) op(B). `tf32` selects fp32 operands (else bf16). Returns 0 or an error code.
int launch_gemm(const GemmOperand& A, const GemmOperand& B, const GemmShape& shape, const GemmEpi& epi, bool tf32,
                int block_n /*0 = auto*/, int max_ctas /*0 = all SMs*/, cudaStream_t stream);
}  // namespace sdumc
