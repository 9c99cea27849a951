// sdumc_b200 — host side of the tcgen05 GEMM: TMA tensor-map construction (cached) and launch.
#include "gemm.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace sdumc {

// ---------------------------------------------------------------------------------
// error channel
// ---------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

// ---------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  long ld;
  long inner, outer;
  int box_inner, box_outer, elem;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && ld == o.ld && inner == o.inner && outer == o.outer && box_inner == o.box_inner &&
           box_outer == o.box_outer && elem == o.elem;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](size_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.ld); mix((size_t)k.inner); mix((size_t)k.outer);
    mix((size_t)k.box_inner); mix((size_t)k.box_outer); mix((size_t)k.elem);
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

// 2-D row-major matrix [outer, inner] with leading dimension ld (elements), box [box_outer, box_inner].
static int get_tmap(const void* ptr, long ld, long inner, long outer, int box_inner, int box_outer, int elem,
                    CUtensorMap* out) {
  SDUMC_CHECK_ARG(ptr != nullptr, "gemm: null operand");
  SDUMC_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "gemm: operand %p not 16-byte aligned", ptr);
  SDUMC_CHECK_ARG((ld * elem) % 16 == 0, "gemm: row pitch %ld bytes not a multiple of 16", ld * elem);
  SDUMC_CHECK_ARG(inner > 0 && outer > 0 && ld >= inner, "gemm: bad operand extent inner=%ld outer=%ld ld=%ld",
                  inner, outer, ld);
  TmapKey key{ptr, ld, inner, outer, box_inner, box_outer, elem};
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(SDUMC_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld * elem)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = fn(&tm, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SDUMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%ld outer=%ld ld=%ld box=%dx%d",
                     (int)r, inner, outer, ld, box_inner, box_outer);
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    if (g_tmaps.size() > 4096) g_tmaps.clear();
    g_tmaps.emplace(key, tm);
  }
  *out = tm;
  return 0;
}

int get_frame_tmap(const void* ptr, int G, int L, int B, int box_rows, CUtensorMap* out) {
  SDUMC_CHECK_ARG(ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "frame tensor %p not 16-byte aligned", ptr);
  SDUMC_CHECK_ARG(G % 64 == 0 && L > 0 && B > 0 && box_rows > 0 && box_rows <= 256, "frame tensor map: bad shape");
  TmapKey key{ptr, -(long)B, G, L, 64, box_rows, 2};     // ld < 0 marks the 3-D maps in the shared cache
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(SDUMC_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t gdim[3] = {(cuuint64_t)G, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)G * 2, (cuuint64_t)L * G * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap tm;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SDUMC_ERR_CUDA, "cuTensorMapEncodeTiled (frames) failed (%d) G=%d L=%d B=%d box_rows=%d", (int)r, G, L,
                     B, box_rows);
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    if (g_tmaps.size() > 4096) g_tmaps.clear();
    g_tmaps.emplace(key, tm);
  }
  *out = tm;
  return 0;
}

void clear_tmap_cache() {
  std::lock_guard<std::mutex> g(g_tmap_mu);
  g_tmaps.clear();
}

// per-device caches (a process may drive several GPUs): SM counts and "attribute already set" flags are indexed
// by the current device; racing first calls write the same values
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SDUMC_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
int num_sms() {
  static int sms[kMaxDevices] = {0};
  const int dev = current_device();
  if (sms[dev] == 0) {
    int n = 0;
    sms[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return sms[dev];
}

template <int kBlockN, bool kTF32, int kKind, bool kPair = false>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& sh, const GemmEpi& ep,
                       int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<kBlockN>;
  static bool attr_done[kMaxDevices] = {false};
  auto kern = gemm_tcgen05_kernel<kBlockN, kTF32, kKind, kPair>;
  const int dev = current_device();
  if (!attr_done[dev]) {
    SDUMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done[dev] = true;
  }
  // pairs: clusters of two CTAs (one TPC) for the cta_group::2 MMAs
  SDUMC_CUDA(launch_kernel(kern, dim3((unsigned)grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, kPair ? 2 : 1, ta, tb,
                           sh, ep));
  SDUMC_CUDA(cudaGetLastError());
  return 0;
}

template <int kKind>
static int launch_n256_bf16(bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& sh,
                            const GemmEpi& ep, int grid, cudaStream_t stream) {
  if (pair) return launch_inst<256, false, kKind, true>(ta, tb, sh, ep, grid, stream);
  return launch_inst<256, false, kKind, false>(ta, tb, sh, ep, grid, stream);
}

int launch_gemm(const GemmOperand& A, const GemmOperand& B, const GemmShape& shape, const GemmEpi& epi, bool tf32,
                int block_n, int max_ctas, cudaStream_t stream) {
  GemmShape sh = shape;
  SDUMC_CHECK_ARG(sh.M > 0 && sh.N > 0 && sh.K > 0, "gemm: empty problem M=%d N=%d K=%d", sh.M, sh.N, sh.K);
  const int elem = tf32 ? 4 : 2;
  const int block_k = 128 / elem;
  const int panel = 128 / elem;
  const int nkb = (sh.K + block_k - 1) / block_k;
  if (sh.k_splits < 1) sh.k_splits = 1;
  if (sh.k_splits > nkb) sh.k_splits = nkb;
  const int m_tiles = (sh.M + 127) / 128;
  const int sms = num_sms();
  if (block_n == 0) {
    if (epi.kind == EPI_KEYPROJ || epi.kind == EPI_INPROJ) {   // specialised epilogues own a whole 256-wide row
      block_n = 256;
    } else {
      const long t256 = (long)m_tiles * ((sh.N + 255) / 256) * sh.k_splits;
      const long t128 = (long)m_tiles * ((sh.N + 127) / 128) * sh.k_splits;
      if (sh.N > 128 && t256 >= sms / 2) block_n = 256;
      else if (sh.N > 64 && t128 >= sms / 2) block_n = 128;
      else block_n = 64;
    }
  }
  SDUMC_CHECK_ARG(block_n == 64 || block_n == 128 || block_n == 256, "gemm: block_n %d unsupported", block_n);
  if (epi.kind == EPI_KEYPROJ) {
    SDUMC_CHECK_ARG(sh.N == 256 && block_n == 256, "gemm: key-proj epilogue needs N == 256");
    SDUMC_CHECK_ARG(epi.nq >= 1 && epi.nq <= 7 && epi.L >= 1 && epi.qv && epi.scores, "gemm: bad key-proj epilogue");
    SDUMC_CHECK_ARG(epi.act == ACT_TANH, "gemm: key-proj epilogue expects tanh");
  }
  if (epi.kind == EPI_INPROJ) {
    SDUMC_CHECK_ARG(sh.N % 32 == 0 && epi.n_tgt >= 0 && epi.n_tgt <= 4, "gemm: bad in-proj epilogue");
    SDUMC_CHECK_ARG(epi.ld_bf16 % 16 == 0, "gemm: in-proj epilogue needs ld %% 16 == 0");
    for (int i = 0; i < epi.n_tgt; ++i)
      SDUMC_CHECK_ARG(epi.tgt[i] && (reinterpret_cast<uintptr_t>(epi.tgt[i]) & 31u) == 0, "gemm: in-proj target %d misaligned", i);
  }
  if (epi.kind == EPI_KEYPROJ && epi.out_bf16) SDUMC_CHECK_ARG(epi.ld_bf16 % 8 == 0, "gemm: key-proj K ld %% 8");
  if (epi.out_f32 && epi.f32_mode != OUT_ATOMIC)
    SDUMC_CHECK_ARG(epi.ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(epi.out_f32) & 15u) == 0,
                    "gemm: fp32 output must be 16-byte aligned with ld %% 4 == 0");
  if (epi.drop_p > 0.f) SDUMC_CHECK_ARG(sh.N % 4 == 0, "gemm: element dropout needs N %% 4 == 0");
  if (epi.out_bf16)
    SDUMC_CHECK_ARG(sh.N % 32 == 0 && epi.ld_bf16 % 16 == 0 && (reinterpret_cast<uintptr_t>(epi.out_bf16) & 31u) == 0,
                    "gemm: bf16 output needs N %% 32 == 0, ld %% 16 == 0 and a 32-byte aligned base");
  if (epi.bias) SDUMC_CHECK_ARG((reinterpret_cast<uintptr_t>(epi.bias) & 15u) == 0, "gemm: bias must be 16-byte aligned");
  if (epi.gate) SDUMC_CHECK_ARG(epi.ld_gate % 4 == 0 && (reinterpret_cast<uintptr_t>(epi.gate) & 15u) == 0,
                                "gemm: gate must be 16-byte aligned with ld %% 4 == 0");
  if (sh.k_splits > 1)
    SDUMC_CHECK_ARG(epi.kind == EPI_GENERIC && epi.f32_mode == OUT_ATOMIC && !epi.out_bf16 && !epi.bias &&
                        epi.act == ACT_NONE,
                    "gemm: split-K requires a plain atomic fp32 epilogue");

  const long n_tiles = (sh.N + block_n - 1) / block_n;
  // CTA pairs (cta_group::2) for the large bf16 N-tile-256 problems: a deeper ring per byte of shared memory and B
  // fetched once per pair.  Needs at least one full wave of work; `max_ctas` < 0 forces single-CTA mode (A/B timing).
  const bool pair = !tf32 && block_n == 256 && m_tiles >= 2 && max_ctas >= 0 && (sms % 2 == 0) &&
                    (long)m_tiles * n_tiles * sh.k_splits >= sms;
  if (max_ctas < 0) max_ctas = 0;
  CUtensorMap ta, tb;
  if (!sh.a_mn) SDUMC_TRY(get_tmap(A.ptr, A.ld, sh.K, sh.M, block_k, 128, elem, &ta));
  else          SDUMC_TRY(get_tmap(A.ptr, A.ld, sh.M, sh.K, panel, block_k, elem, &ta));
  if (!sh.b_mn) SDUMC_TRY(get_tmap(B.ptr, B.ld, sh.K, sh.N, block_k, pair ? block_n / 2 : block_n, elem, &tb));
  else          SDUMC_TRY(get_tmap(B.ptr, B.ld, sh.N, sh.K, panel, block_k, elem, &tb));

  // scheduling units: tiles, or - for pairs - two neighbouring M tiles
  long tiles = (long)(pair ? (m_tiles + 1) / 2 : m_tiles) * n_tiles * sh.k_splits;
  const int slots = pair ? sms / 2 : sms;
  int grid = (int)(tiles < slots ? tiles : slots);
  if (max_ctas > 0 && grid > (pair ? max_ctas / 2 : max_ctas)) grid = std::max(1, pair ? max_ctas / 2 : max_ctas);
  // resident B: one N tile whose whole K extent fits beside an A-only ring, and at least two tiles per CTA to
  // amortise it
  const int stages = block_n == 256 ? 4 : (block_n == 128 ? 6 : 8);
  sh.b_res = (n_tiles == 1 && sh.k_splits == 1 && nkb <= stages && tiles >= 2L * grid) ? 1 : 0;
  if (pair) grid *= 2;

  if (tf32) {
    SDUMC_CHECK_ARG(epi.kind == EPI_GENERIC, "gemm: tf32 operands support the generic epilogue only");
    if (block_n == 256) return launch_inst<256, true, KIND_GENERIC>(ta, tb, sh, epi, grid, stream);
    if (block_n == 128) return launch_inst<128, true, KIND_GENERIC>(ta, tb, sh, epi, grid, stream);
    return launch_inst<64, true, KIND_GENERIC>(ta, tb, sh, epi, grid, stream);
  }
  if (epi.kind == EPI_INPROJ) {
    SDUMC_CHECK_ARG(block_n == 256, "gemm: in-proj epilogue is instantiated for block_n 256 (N = 256)");
    return launch_n256_bf16<KIND_INPROJ>(pair, ta, tb, sh, epi, grid, stream);
  }
  if (epi.kind == EPI_KEYPROJ) return launch_n256_bf16<KIND_KEYPROJ>(pair, ta, tb, sh, epi, grid, stream);
  // 'bf16 += acc * frame mask' without any other option: the dH accumulation of the attention backward
  const bool pure_rmw = block_n == 256 && epi.out_bf16 && epi.bf16_mode == OUT_ADD && !epi.out_f32 && !epi.bias &&
                        epi.act == ACT_NONE && !epi.gate && epi.drop_p == 0.f;
  if (pure_rmw) return launch_n256_bf16<KIND_RMW>(pair, ta, tb, sh, epi, grid, stream);
  if (block_n == 256) return launch_n256_bf16<KIND_GENERIC>(pair, ta, tb, sh, epi, grid, stream);
  if (block_n == 128) return launch_inst<128, false, KIND_GENERIC>(ta, tb, sh, epi, grid, stream);
  return launch_inst<64, false, KIND_GENERIC>(ta, tb, sh, epi, grid, stream);
}

}  // namespace sdumc
