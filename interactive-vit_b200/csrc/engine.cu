// libvitb200.so — host side of the B200-native ViT forward engine and its C ABI (include/vitb200.h).
//
// The engine owns: bf16 weights, an fp32 residual token stream x [B, N, d], bf16 GEMM operands, fp32
// attention-map buffers, one CUDA stream.  A forward is the kernel sequence of torchvision's
// VisionTransformer.forward (vision_transformer.py:289-306; EncoderBlock 110-119), i.e. what the reference
// would execute inside Model.compute (main/context.py:79-88) for a ViT plugin:
//   patchify -> GEMM(+bias +pos, scatter behind class token) -> cls rows
//   L x [ GEMM qkv (LN1 folded) -> fused attention (+maps) -> GEMM out_proj (+x) -> GEMM fc1 (LN2 folded, GELU) -> GEMM fc2 (+x) ]
//   LN(class rows) -> GEMM head -> logits;   rollout over the head-averaged maps.
//
// LayerNorm is folded into the GEMMs.  LN(x) W^T + b = rstd * (x W'^T - mean * colsum) + b' with W' = gamma o W,
// colsum[n] = sum_k W'[n, k], b' = b + W beta (fold_ln_weight_kernel, once after the weights are loaded).  Every GEMM
// that updates the fp32 residual stream x (patch embedding, out_proj, fc2) also writes the bf16 copy xb of its
// result -- the A operand of the next GEMM -- and per row the partial sums (sum, sum of squares) of each 32-column
// chunk into fixed slots of `stats`; the consuming GEMM's epilogue reduces them to mean / rstd and applies the
// affine correction per element.  This removes the 24 LayerNorm launches per forward (232 MB of HBM traffic each).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/vitb200.h"
#include "attention.cuh"
#include "attention_pp.cuh"
#include "attention_long.cuh"
#include "attention_precise.cuh"
#include "gemm.cuh"
#include "patch_embed.cuh"
#include "peer.cuh"
#include "rowwise.cuh"

namespace vitb200 {

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU_TRY(expr)                                                                                      \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return fail(VITB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define VT_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != VITB200_OK) return _s; \
  } while (0)

// ------------------------------------------------------------------------------------------ TMA descriptors
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
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

// 2-D bf16 row-major tensor [rows, cols] (cols contiguous, row pitch ld elements), box = box_cols x box_rows,
// 128-byte swizzle (box_cols * 2 bytes must be 128), out-of-bounds elements read as zero.
static int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (box_cols * 2 != 128 || box_rows == 0 || box_rows > 256)
    return fail(VITB200_ERR_INVALID, "bad TMA box %u x %u", box_rows, box_cols);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(VITB200_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p, pitch %llu B)", base,
                (unsigned long long)(ld * 2));
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VITB200_OK;
}

// Generic tiled tensor map: `rank` dimensions, fastest first; strides in bytes for dimensions 1 .. rank-1 (multiples of 16);
// swizzle chosen by the caller (the box's innermost extent in bytes must not exceed the swizzle span).
static int make_tmap_nd(CUtensorMap* tm, CUtensorMapDataType dtype, const void* base, int rank, const cuuint64_t* dims,
                        const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(VITB200_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p)", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (strides_bytes[i] % 16 != 0) return fail(VITB200_ERR_INVALID, "TMA stride %d (%llu B) must be a multiple of 16", i, (unsigned long long)strides_bytes[i]);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled (%d-D) failed with CUresult %d", rank, (int)r);
  return VITB200_OK;
}

// 3-D bf16 tensor [batch][rows][cols] (cols contiguous, row pitch ld elements, image pitch rows * ld), box =
// box_cols x box_rows x 1, 128-byte swizzle.  Used for stores that must clip at the end of each image.
static int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (box_cols * 2 != 128 || box_rows == 0 || box_rows > 256)
    return fail(VITB200_ERR_INVALID, "bad TMA box %u x %u", box_rows, box_cols);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(VITB200_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p, pitch %llu B)", base,
                (unsigned long long)(ld * 2));
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * 2, rows * ld * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
  return VITB200_OK;
}

// 3-D fp32 tensor [batch][rows][cols], box = 32 cols (128 B) x box_rows x 1, 128-byte swizzle (attention-map stores).
static int make_tmap_f32_3d(CUtensorMap* tm, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                            uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 4) % 16 != 0 || box_rows == 0 || box_rows > 256)
    return fail(VITB200_ERR_INVALID, "fp32 TMA operand must be 16-byte aligned (base %p, pitch %llu B)", base,
                (unsigned long long)(ld * 4));
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * 4, rows * ld * 4};
  cuuint32_t box[3] = {32, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VITB200_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 3-D) failed with CUresult %d", (int)r);
  return VITB200_OK;
}

// ------------------------------------------------------------------------------------------ launch helpers
// Per-DEVICE launch state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current device only, and the
// SM count / co-resident cluster count are device properties: a process may hold engines on several GPUs
// (VitEngine(cfg, device=k)), so none of this may live in function-local statics.  One mutex guards the table; it is
// taken once per launch (tens of nanoseconds next to a kernel launch).
struct DeviceCtx {
  int sms = 0;
  std::map<const void*, int> func_smem;    // kernel -> dynamic smem limit configured on this device
  std::map<const void*, int> gemm_slots;   // GEMM instantiation -> persistent scheduler slots on this device
  float2* op_stats = nullptr;              // scratch of the single-kernel attention entry point (parity tests)
  void* op_qkv = nullptr;                  // ... and its fp16-V copy of the input
  size_t op_qkv_cap = 0;
  void* op_parts = nullptr;                // ... and the head-average scratch of its head-split small launches
  size_t op_parts_cap = 0;
  size_t op_stats_cap = 0;
};
static std::mutex g_dev_mu;
static std::map<int, DeviceCtx> g_dev;

// call with g_dev_mu held
static DeviceCtx& dev_ctx_locked() {
  int dev = 0;
  cudaGetDevice(&dev);
  DeviceCtx& c = g_dev[dev];
  if (c.sms == 0) {
    cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
    if (c.sms <= 0) c.sms = 148;
  }
  return c;
}

static int device_sms() {
  std::lock_guard<std::mutex> lock(g_dev_mu);
  return dev_ctx_locked().sms;
}

// Programmatic dependent launch (ptx.cuh grid_dep_wait): kernels of the forward are launched so that their prologue may
// overlap the tail of the previous kernel of the stream.  VITB200_PDL=0 launches them fully serialised (A/B switch;
// read once).  Only kernels that execute grid_dep_wait() before touching global memory may go through launch_pdl.
static bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("VITB200_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_if(bool allow, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = (allow && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl_if(true, kern, grid, block, smem, st, static_cast<Args&&>(args)...);
}

// Raise the dynamic shared-memory limit of `kern` on the CURRENT device (once per device and kernel).
static int ensure_func_smem(const void* kern, int bytes) {
  std::lock_guard<std::mutex> lock(g_dev_mu);
  DeviceCtx& c = dev_ctx_locked();
  auto it = c.func_smem.find(kern);
  if (it != c.func_smem.end() && it->second >= bytes) return VITB200_OK;
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  c.func_smem[kern] = bytes;
  return VITB200_OK;
}

// Host mirror of GemmWork::num_units (gemm.cuh).
static int gemm_units(int M, int N, int tile_m, int bn, int pairs) {
  const int m_tiles = (M + tile_m - 1) / tile_m, n_tiles = (N + bn - 1) / bn;
  if (pairs == 1) return m_tiles * n_tiles;
  return (m_tiles / 2) * n_tiles + ((m_tiles & 1) ? (n_tiles + 1) / 2 : 0);
}

template <int BN, int kPair, int kPairs, bool kGelu, bool kOutF32, bool kResid, bool kRemap, bool kLnIn = false,
          bool kPrefetch = false>
static int launch_gemm_t(const CUtensorMap* maps, GemmShape sh, const GemmEpilogue& ep, cudaStream_t st) {
  const CUtensorMap &ta = maps[0], &tw = maps[1], &ta_lo = maps[2], &tw_lo = maps[3];
  using C = gemm_cfg::Cfg<BN, kPair, gemm_cfg::epi_warps(kGelu), kPrefetch>;
  auto kern = gemm_bf16_kernel<BN, kPair, kPairs, kGelu, kOutF32, kResid, kRemap, kLnIn, kPrefetch>;
  constexpr int kClusterCtas = kPair * kPairs;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClusterCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // persistent: one CTA / pair / 4-CTA cluster per scheduler slot; the slots are what the device can co-schedule
  // (4-CTA clusters do not tile every GPC: 148 SMs hold fewer than 37 of them)
  VT_TRY(ensure_func_smem(reinterpret_cast<const void*>(kern), C::kSmemBytes));
  int slots = 0;
  {
    std::lock_guard<std::mutex> lock(g_dev_mu);
    DeviceCtx& dc = dev_ctx_locked();
    auto it = dc.gemm_slots.find(reinterpret_cast<const void*>(kern));
    if (it != dc.gemm_slots.end()) {
      slots = it->second;
    } else {
      slots = dc.sms / kClusterCtas;
      if (kClusterCtas > 2) {
        cfg.gridDim = dim3(dc.sms / kClusterCtas * kClusterCtas);
        int n = 0;
        CU_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
        if (n <= 0) return fail(VITB200_ERR_CUDA, "gemm: no %d-CTA cluster fits on this device", kClusterCtas);
        if (n < slots) slots = n;
        const char* v = getenv("VITB200_GEMM_SLOTS");
        if (v && atoi(v) > 0 && atoi(v) < slots) slots = atoi(v);
      }
      dc.gemm_slots[reinterpret_cast<const void*>(kern)] = slots;
    }
  }
  const int units = gemm_units(sh.M, sh.N, gemm_cfg::BM * kPair, BN, kPairs);
  cfg.gridDim = dim3((units < slots ? units : slots) * kClusterCtas);
  CU_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tw, ta_lo, tw_lo, sh, ep));
  return VITB200_OK;
}

// VITB200_GEMM_PAIR: 1 = single-CTA baseline, 2 (default) = CTA pairs, 4 = two pairs per cluster with the shared
// operand multicast (BN = 256 shapes only; others fall back to pairs).  Measured on B200 (profiles/README.md): the
// multicast variant is NOT faster (1190 vs 1249 TF/s on the qkv shape): the bound is bytes DELIVERED into the SMs,
// which multicast does not reduce, and only 33 four-CTA clusters (132 of 148 SMs) are co-resident.
static int gemm_pair_mode() {
  static int mode = 0;
  if (mode == 0) {
    const char* v = getenv("VITB200_GEMM_PAIR");
    mode = (v && v[0] == '1') ? 1 : (v && v[0] == '4') ? 4 : 2;
  }
  return mode;
}

template <int BN, int kPair, int kPairs>
static int launch_gemm_bn(const CUtensorMap* maps, GemmShape sh, const GemmEpilogue& ep, bool gelu, bool out_f32,
                          cudaStream_t st) {
  const bool resid = ep.resid != nullptr, remap = ep.group_rows > 0, ln_in = ep.row_stats_in != nullptr;
  if (ln_in) {
    if (ep.colsum == nullptr || out_f32 || resid || remap)
      return fail(VITB200_ERR_INVALID, "gemm: folded-LayerNorm epilogue needs colsum and a bf16 output");
    if (sh.K % 128 != 0 || (ep.stats_in_slots != sh.K / 64 && ep.stats_in_slots != sh.K / 128))
      return fail(VITB200_ERR_INVALID, "gemm: folded LayerNorm needs K a multiple of 128 and K / 64 or K / 128 partial sums per row (K=%d, slots=%d)",
                  sh.K, ep.stats_in_slots);
    if (gelu) return launch_gemm_t<BN, kPair, kPairs, true, false, false, false, true>(maps, sh, ep, st);
    return launch_gemm_t<BN, kPair, kPairs, false, false, false, false, true>(maps, sh, ep, st);
  }
  // (one statistics slot per epilogue-warp column group: BN / 2 columns)
  if (ep.xb != nullptr && (!resid || ep.row_stats_out == nullptr || sh.N % 128 != 0 || ep.stats_slots != sh.N / (BN / 2)))
    return fail(VITB200_ERR_INVALID, "gemm: the bf16 copy + row statistics are produced by residual epilogues only, N a multiple "
                                     "of 128, one slot per %d columns (got %d slots for N = %d)", BN / 2, ep.stats_slots, sh.N);
  if (gelu && !out_f32 && !resid && !remap) return launch_gemm_t<BN, kPair, kPairs, true, false, false, false>(maps, sh, ep, st);
  if (!gelu && !out_f32 && !resid && !remap) return launch_gemm_t<BN, kPair, kPairs, false, false, false, false>(maps, sh, ep, st);
  if (!gelu && out_f32 && !resid && !remap) return launch_gemm_t<BN, kPair, kPairs, false, true, false, false>(maps, sh, ep, st);
  if (!gelu && out_f32 && resid && !remap) {
    // short K: the epilogue (fp32 residual read + write) is the bound -> prefetch the addend; long K: keep the smem stage
    if (sh.K <= 1024) return launch_gemm_t<BN, kPair, kPairs, false, true, true, false, false, true>(maps, sh, ep, st);
    return launch_gemm_t<BN, kPair, kPairs, false, true, true, false>(maps, sh, ep, st);
  }
  if (!gelu && out_f32 && resid && remap) return launch_gemm_t<BN, kPair, kPairs, false, true, true, true>(maps, sh, ep, st);
  return fail(VITB200_ERR_INVALID, "gemm: epilogue combination not instantiated (gelu=%d f32=%d resid=%d remap=%d)", gelu,
              out_f32, resid, remap);
}

// Small problem: at most a quarter of the SMs would get a 256 x 256 (or 256 x 128) pair tile.
static bool gemm_small(int M, int N) { return gemm_units(M, N, 256, (N % 256 == 0) ? 256 : 128, 1) * 4 <= device_sms(); }

// Columns per LayerNorm statistics slot for a forward over M token rows of width d: 128 for widths that are multiples of
// 256 -- unless the problem is small, where the producing GEMMs (patch embedding, out_proj, fc2: N = d) run 128-wide
// single-CTA tiles, whose epilogue-warp column groups are 64 wide.
static int stats_width(int M, int d) {
  static const int small_mode = [] {
    const char* v = getenv("VITB200_GEMM_SMALL");
    return (v && v[0] == '0') ? 0 : 1;
  }();
  if (d % 256 != 0) return 64;
  return (small_mode && gemm_pair_mode() == 2 && gemm_small(M, d)) ? 64 : 128;
}

// out = epilogue(A[M,K] * W[N,K]^T): picks the tile width and the epilogue instantiation.
static int launch_gemm(const void* a, int lda, const void* w, int M, int N, int K, const GemmEpilogue& ep, bool gelu,
                       bool out_f32, cudaStream_t st, const void* a_lo = nullptr, const void* w_lo = nullptr) {
  if (M <= 0 || N <= 0 || K <= 0) return fail(VITB200_ERR_INVALID, "gemm: empty shape %d x %d x %d", M, N, K);
  if (N % 8 != 0 || K % 8 != 0) return fail(VITB200_ERR_INVALID, "gemm: N and K must be multiples of 8 (N=%d K=%d)", N, K);
  const int BN = (N % 256 == 0) ? 256 : 128;
  int mode = gemm_pair_mode();
  // the multicast clusters pay off once there is more than one wave of tiles; small problems keep the finer pairs
  // (VITB200_GEMM_CLUSTER_MIN_TILES overrides the threshold: the parity tests force clusters onto small shapes)
  static const int env_min_tiles = [] {   // process-wide knob, not device state
    const char* v = getenv("VITB200_GEMM_CLUSTER_MIN_TILES");
    return v ? atoi(v) : -1;
  }();
  const int min_tiles = env_min_tiles >= 0 ? env_min_tiles : (mode == 4 ? device_sms() / 2 + 1 : 0);
  if (mode == 4 && (BN != 256 || gemm_units(M, N, 256, 256, 1) < min_tiles)) mode = 2;
  int pair = mode == 1 ? 1 : 2;
  int bn = BN;
  // Small problems (the single-image request: M = 197 -> one 256-row tile per 256 columns, 3..12 CTA pairs on 148 SMs)
  // are latency-bound by ONE tile's main loop: single-CTA 128 x 128 tiles put 4x the SMs on the same work (fc2, K = 3072:
  // 48 k-blocks of a quarter of the MMA time each).  Same k order per output element: results are bit-identical.
  static const int small_mode = [] {
    const char* v = getenv("VITB200_GEMM_SMALL");
    return (v && v[0] == '0') ? 0 : 1;
  }();
  // GEMMs that PRODUCE LayerNorm partial sums write one slot per epilogue-warp column group (tile width / 2), so their
  // tile width follows the slot width the engine chose for this batch size (stats_width): 64 -> 128-wide tiles.
  if (ep.row_stats_out != nullptr && ep.stats_slots > 0 && mode != 4) bn = (N / ep.stats_slots == 64) ? 128 : 256;
  if (small_mode && mode == 2 && N % 128 == 0 && (ep.row_stats_out == nullptr || bn == 128) && gemm_small(M, N))
    pair = 1, bn = 128;
  // ... except where ONE 256-row pair tile covers all rows and the output is narrow (patch embedding, out_proj, fc2 of
  // a single image: N = d): pairs of the same 128-wide tiles give the same number of CTAs, and each CTA streams half
  // of the W panel -- at B = 1 a GEMM is bound by what one SM can keep in flight (VITB200_GEMM_SMALL_PAIR=0 disables;
  // same k order, same statistics slots: bit-identical).  Measured at B = 1: 0.659 -> 0.631 ms per forward (fc2 alone:
  // 0.638).
  static const int small_pair = [] {
    const char* v = getenv("VITB200_GEMM_SMALL_PAIR");
    return (v && v[0] == '0') ? 0 : 1;
  }();
  // (the wide GEMMs of one image -- qkv, fc1: 36 / 48 tiles -- gain 0.6 % as pairs: left alone)
  if (small_pair && pair == 1 && bn == 128 && M <= 256 && M > 128 && ep.row_stats_out != nullptr) pair = 2;
  if ((a_lo == nullptr) != (w_lo == nullptr)) return fail(VITB200_ERR_INVALID, "gemm: split-bf16 needs both low operands");
  CUtensorMap maps[4];  // A, W, A_lo, W_lo (the low maps alias the high ones when the operands are plain bf16)
  VT_TRY(make_tmap_bf16(&maps[0], a, M, K, lda, mode == 4 ? 64 : gemm_cfg::BM, gemm_cfg::BK));
  VT_TRY(make_tmap_bf16(&maps[1], w, N, K, K, mode == 4 ? 64 : bn / pair, gemm_cfg::BK));
  maps[2] = maps[0], maps[3] = maps[1];
  GemmShape sh{M, N, K};
  static const int prefetch_w = [] {   // VITB200_GEMM_PREFETCH_W=0: no L2 prefetch of the weight panel in small launches
    const char* v = getenv("VITB200_GEMM_PREFETCH_W");
    return (v && v[0] == '0') ? 0 : 1;
  }();
  sh.prefetch_w = prefetch_w;
  if (a_lo != nullptr) {
    VT_TRY(make_tmap_bf16(&maps[2], a_lo, M, K, lda, mode == 4 ? 64 : gemm_cfg::BM, gemm_cfg::BK));
    VT_TRY(make_tmap_bf16(&maps[3], w_lo, N, K, K, mode == 4 ? 64 : bn / pair, gemm_cfg::BK));
    sh.split = 3;
  }
  if (mode == 4) return launch_gemm_bn<256, 2, 2>(maps, sh, ep, gelu, out_f32, st);
  if (bn == 256) {
    return pair == 2 ? launch_gemm_bn<256, 2, 1>(maps, sh, ep, gelu, out_f32, st)
                     : launch_gemm_bn<256, 1, 1>(maps, sh, ep, gelu, out_f32, st);
  }
  return pair == 2 ? launch_gemm_bn<128, 2, 1>(maps, sh, ep, gelu, out_f32, st)
                   : launch_gemm_bn<128, 1, 1>(maps, sh, ep, gelu, out_f32, st);
}

static int launch_layernorm(const float* x, long in_stride, const float* g, const float* b, __nv_bfloat16* y, int rows,
                            int d, float eps, cudaStream_t st, __nv_bfloat16* y_lo = nullptr) {
  if (rows <= 0) return fail(VITB200_ERR_INVALID, "layernorm: no rows");
  if (d % 128 != 0) return fail(VITB200_ERR_INVALID, "layernorm: width %d must be a multiple of 128", d);
  const int blocks = (rows + 7) / 8;  // 8 warps (rows) per 256-thread block
#define VT_LN_CASE(V)                                                                                \
  case V:                                                                                            \
    CU_TRY(launch_pdl(layernorm_f32_bf16_kernel<V>, dim3(blocks), dim3(256), 0, st, x, in_stride, g, b, y, rows, eps, y_lo)); \
    break;
  switch (d / 128) {
    VT_LN_CASE(1) VT_LN_CASE(2) VT_LN_CASE(3) VT_LN_CASE(4) VT_LN_CASE(5) VT_LN_CASE(6) VT_LN_CASE(8) VT_LN_CASE(10)
    default:
      return fail(VITB200_ERR_INVALID, "layernorm: width %d not instantiated (multiples of 128 up to 1280)", d);
  }
#undef VT_LN_CASE
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

// Long / wide variant (attention_long.cuh): any N, head dims 64..128.  `stats` holds B*H*N float2 (row max, 1/sum).
static int launch_attention_long(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* avg, float* cls, float* heads, int B,
                                 int N, int H, int D, int pitch, float2* stats, cudaStream_t st,
                                 const __nv_bfloat16* qkv_lo = nullptr, __nv_bfloat16* ctx_lo = nullptr) {
  using namespace attn_long_cfg;
  if (D < 64 || D > 128 || D % 16 != 0) return fail(VITB200_ERR_INVALID, "attention: head dim %d not in 64..128 step 16", D);
  if ((avg || heads) && (pitch < N || pitch % 4 != 0))
    return fail(VITB200_ERR_INVALID, "attention: map pitch %d must be >= %d and a multiple of 4", pitch, N);
  if (!stats) return fail(VITB200_ERR_INVALID, "attention: statistics buffer missing");
  const int d = H * D;
  const bool split = qkv_lo != nullptr;
  if (split && ctx_lo == nullptr) return fail(VITB200_ERR_INVALID, "attention: the fp32x3 mode needs the low half of the context");
  if (split && D != 64) {
    // fp32x3 mode at head dims other than 64 (ViT-H: 80): fp32 on the CUDA cores (attention_precise.cuh)
    AttnLongParams p;
    p.B = B, p.N = N, p.H = H, p.D = D, p.d = d, p.q_tiles = 0, p.k_blocks = 0, p.scale_log2 = 0.f;
    p.ctx = ctx, p.ctx_lo = ctx_lo, p.stats = stats, p.avg_map = avg, p.head_map = heads, p.cls_map = cls, p.ldmap = pitch;
    const size_t smem = attn_precise_cfg::smem_bytes(N, D);
    if (smem > 227 * 1024) return fail(VITB200_ERR_INVALID, "attention (fp32x3, head dim %d): %d tokens need %zu bytes of shared memory", D, N, smem);
    VT_TRY(ensure_func_smem((const void*)attention_precise_kernel, (int)smem));
    const int grid = B * ((N + attn_precise_cfg::QT - 1) / attn_precise_cfg::QT);
    attention_precise_kernel<<<grid, attn_precise_cfg::kThreads, smem, st>>>(qkv, qkv_lo, p);
    CU_TRY(cudaGetLastError());
    return VITB200_OK;
  }
  CUtensorMap tqkv, tqkv_lo;
  VT_TRY(make_tmap_bf16_3d(&tqkv, qkv, B, N, 3 * d, 3 * d, 128, 64));
  tqkv_lo = tqkv;
  if (split) VT_TRY(make_tmap_bf16_3d(&tqkv_lo, qkv_lo, B, N, 3 * d, 3 * d, 128, 64));
  {
    const std::pair<const void*, int> kerns[] = {
        {(const void*)attention_long_ctx_kernel<false, false, true>, kSmemCtx},
        {(const void*)attention_long_ctx_kernel<true, false, true>, kSmemCtxSplit},
        {(const void*)attention_long_ctx_kernel<false, true, true>, kSmemCtxCompact},
        {(const void*)attention_long_ctx_kernel<false, false, false>, kSmemCtx},
        {(const void*)attention_long_ctx_kernel<true, false, false>, kSmemCtxSplit},
        {(const void*)attention_long_ctx_kernel<false, true, false>, kSmemCtxCompact},
        {(const void*)attention_long_maps_kernel<false, false, true>, kSmemMapsCompact},
        {(const void*)attention_long_maps_kernel<true, false, true>, kSmemMapsCompact},
        {(const void*)attention_long_maps_kernel<false, false>, kSmemMaps},
        {(const void*)attention_long_maps_kernel<true, false>, kSmemMaps},
        {(const void*)attention_long_maps_kernel<false, true>, kSmemMaps},
        {(const void*)attention_long_maps_kernel<true, true>, kSmemMaps}};
    for (const auto& k : kerns) VT_TRY(ensure_func_smem(k.first, k.second));
  }
  AttnLongParams p;
  p.B = B, p.N = N, p.H = H, p.D = D, p.d = d;
  p.q_tiles = (N + BM - 1) / BM, p.k_blocks = (N + BK - 1) / BK;
  p.scale_log2 = (1.0f / sqrtf((float)D)) * 1.4426950408889634f;
  p.ctx = ctx, p.ctx_lo = ctx_lo, p.stats = stats, p.avg_map = avg, p.head_map = heads, p.cls_map = cls, p.ldmap = pitch;
  // head dim 64 in bf16 mode: the compact instantiation, two CTAs per SM (VITB200_ATTN_LONG_COMPACT=0 keeps one)
  static int compact_mode = -1;
  if (compact_mode < 0) {
    const char* v = getenv("VITB200_ATTN_LONG_COMPACT");
    compact_mode = (v && v[0] == '0') ? 0 : 1;
  }
  // VITB200_ATTN_LONG_ONLINE=0: the two-pass context kernel (separate maximum pass) instead of the lazily rescaled
  // single pass (read at every launch: the parity tests compare both)
  const char* online_env = getenv("VITB200_ATTN_LONG_ONLINE");
  const bool online = !(online_env && online_env[0] == '0');
  const unsigned cgrid = (unsigned)(B * p.q_tiles * H);
  if (online) {
    if (split) attention_long_ctx_kernel<true, false, true><<<cgrid, kThreads, kSmemCtxSplit, st>>>(tqkv, tqkv_lo, p);
    else if (D == 64 && compact_mode) attention_long_ctx_kernel<false, true, true><<<cgrid, kThreads, kSmemCtxCompact, st>>>(tqkv, tqkv_lo, p);
    else attention_long_ctx_kernel<false, false, true><<<cgrid, kThreads, kSmemCtx, st>>>(tqkv, tqkv_lo, p);
  } else {
    if (split) attention_long_ctx_kernel<true, false, false><<<cgrid, kThreads, kSmemCtxSplit, st>>>(tqkv, tqkv_lo, p);
    else if (D == 64 && compact_mode) attention_long_ctx_kernel<false, true, false><<<cgrid, kThreads, kSmemCtxCompact, st>>>(tqkv, tqkv_lo, p);
    else attention_long_ctx_kernel<false, false, false><<<cgrid, kThreads, kSmemCtx, st>>>(tqkv, tqkv_lo, p);
  }
  CU_TRY(cudaGetLastError());
  if (avg || cls || heads) {
    const int grid = B * p.q_tiles * p.k_blocks;
    if (split) {
      if (heads) attention_long_maps_kernel<true, true><<<grid, kThreads, kSmemMaps, st>>>(tqkv, tqkv_lo, p);
      else attention_long_maps_kernel<false, true><<<grid, kThreads, kSmemMaps, st>>>(tqkv, tqkv_lo, p);
    } else if (D == 64 && compact_mode) {
      // 64-key blocks, two CTAs per SM; the second map argument carries the 64-row K box
      CUtensorMap tk64;
      VT_TRY(make_tmap_bf16_3d(&tk64, qkv, B, N, 3 * d, 3 * d, 64, 64));
      AttnLongParams pc = p;
      pc.k_blocks = (N + 63) / 64;
      const int gridc = B * pc.q_tiles * pc.k_blocks;
      if (heads) attention_long_maps_kernel<true, false, true><<<gridc, kThreads, kSmemMapsCompact, st>>>(tqkv, tk64, pc);
      else attention_long_maps_kernel<false, false, true><<<gridc, kThreads, kSmemMapsCompact, st>>>(tqkv, tk64, pc);
    } else {
      if (heads) attention_long_maps_kernel<true, false><<<grid, kThreads, kSmemMaps, st>>>(tqkv, tqkv_lo, p);
      else attention_long_maps_kernel<false, false><<<grid, kThreads, kSmemMaps, st>>>(tqkv, tqkv_lo, p);
    }
    CU_TRY(cudaGetLastError());
  }
  return VITB200_OK;
}

// Row pitch of the attention maps for N tokens: the fused kernel pads keys to 16, the long kernel needs 4.
static int attention_pitch(int N) { return (N + 15) / 16 * 16; }
static bool attention_is_fused(int N, int D) { return D == 64 && (N + 15) / 16 * 16 <= attn_cfg::KP_MAX; }

// Images of head-average scratch a small launch can need (launch_attention: parts x batch <= SMs / query tiles).
static size_t attention_parts_bytes(int N, int pitch) {
  const int q_tiles = (N + attn_cfg::BM - 1) / attn_cfg::BM;
  return (size_t)(device_sms() / q_tiles) * N * pitch * sizeof(float);
}

static int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* ctx, float* avg, float* cls, float* heads, int B,
                            int N, int H, int D, int pitch, float2* stats, cudaStream_t st,
                            const __nv_bfloat16* qkv_lo = nullptr, __nv_bfloat16* ctx_lo = nullptr,
                            float* parts = nullptr /* attention_parts_bytes() of scratch, or none */) {
  using namespace attn_cfg;
  if (qkv_lo != nullptr || !attention_is_fused(N, D))
    return launch_attention_long(qkv, ctx, avg, cls, heads, B, N, H, D, pitch, stats, st, qkv_lo, ctx_lo);
  const int KP = (N + 15) / 16 * 16;
  if ((avg || heads) && (pitch < KP || pitch % 4 != 0))
    return fail(VITB200_ERR_INVALID, "attention: map pitch %d must be >= %d and a multiple of 4", pitch, KP);
  const int d = H * D;
  CUtensorMap tq, tkv, tctx;
  VT_TRY(make_tmap_bf16(&tq, qkv, (uint64_t)B * N, 3 * d, 3 * d, BM, D));
  VT_TRY(make_tmap_bf16(&tkv, qkv, (uint64_t)B * N, 3 * d, 3 * d, KP / 2, D));
  VT_TRY(make_tmap_bf16_3d(&tctx, ctx, B, N, d, d, 32, D));
  CUtensorMap tavg = tctx;  // placeholder when no head-averaged map is requested (never dereferenced)
  if (avg) VT_TRY(make_tmap_f32_3d(&tavg, avg, B, N, pitch, pitch, BM));
  VT_TRY(ensure_func_smem((const void*)attention_kernel<false>, kSmemBytes));
  VT_TRY(ensure_func_smem((const void*)attention_kernel<true>, kSmemBytes));
  VT_TRY(ensure_func_smem((const void*)attention_kernel<false, true>, kSmemBytes));
  AttnParams p;
  p.B = B, p.N = N, p.H = H, p.d = d, p.KP = KP;
  p.scale_log2 = (1.0f / sqrtf((float)D)) * 1.4426950408889634f;
  p.ctx = ctx, p.avg_map = avg, p.head_map = heads, p.cls_map = cls, p.ldmap = pitch;
  p.q_tiles = (N + BM - 1) / BM;
  // Tail balancing (see AttnParams::full_items): the items of a last, at most half-full round are split by heads.
  const int items = B * p.q_tiles, sms = device_sms();
  const int rem = items % sms;
  // VITB200_ATTN_SPLIT: 0 = never split an item, 2 = two parts at most (read at every launch: the parity tests compare the
  // variants in one process)
  const char* split_env = getenv("VITB200_ATTN_SPLIT");
  const int split_mode = (split_env && split_env[0] == '0') ? 0 : (split_env && split_env[0] == '2') ? 2 : 1;
  p.full_items = items;
  // Small launches (the single-image request, batches up to ~24): every item split over up to H CTAs by heads, so that a
  // request's attention is not 6 heads in sequence on 4 of 148 SMs.  The parts' head averages go to scratch slabs and
  // avg_parts_sum_kernel adds them in index order (a reduce-add of more than two parts would not be reproducible).
  const int S = (split_mode == 1 && (parts || !avg)) ? std::min(std::min(H, kAvgPartsMax), sms / items) : 1;
  if (S >= 3) {
    p.full_items = 0, p.split = S, p.part_images = avg ? B : 0;
    if (avg) VT_TRY(make_tmap_f32_3d(&tavg, parts, (uint64_t)S * B, N, pitch, pitch, BM));
  } else if (split_mode && H >= 2 && rem > 0 && 2 * rem <= sms) {
    // (also when the whole launch is such a round: up to sms / 2 items run on twice the SMs.  Two parts: a + b is
    // commutative, so the reduce-add into the zeroed rows stays bit-reproducible.)
    p.full_items = items - rem;
  }
  const int grid = p.full_items + p.split * (items - p.full_items);
  bool after_kernel = true;   // the launch directly follows a kernel in the stream (PDL) -- not when a memset sits between
  if (avg && p.full_items < items && p.part_images == 0) {
    after_kernel = false;
    // the split CTAs reduce-add their halves of the head average: zero the images they touch first (a full CTA of
    // the first such image simply stores over the zeros)
    const int b0 = p.full_items / p.q_tiles;
    CU_TRY(cudaMemsetAsync(avg + (size_t)b0 * N * pitch, 0, (size_t)(B - b0) * N * pitch * sizeof(float), st));
  }
  static int full_mode = -1;   // VITB200_ATTN_FULL=0: generic granule guards also at KP = 208
  if (full_mode < 0) {
    const char* v = getenv("VITB200_ATTN_FULL");
    full_mode = (v && v[0] == '0') ? 0 : 1;
  }
  // 197-token production shape: two heads in flight per SM (attention_pp.cuh); VITB200_ATTN_PP=0 keeps the one-head kernel
  // (read at every launch: the parity tests compare both kernels in one process)
  const char* pp_env = getenv("VITB200_ATTN_PP");
  const bool pp_mode = !(pp_env && pp_env[0] == '0');
  if (!heads && pp_mode && KP == KP_MAX && N >= attn_pp_cfg::kMinTokens && N <= attn_pp_cfg::kMaxTokens) {
    VT_TRY(ensure_func_smem((const void*)attention_pp_kernel, attn_pp_cfg::kSmemBytesPP));
    CU_TRY(launch_pdl_if(after_kernel, attention_pp_kernel, dim3(grid), dim3(kThreads), attn_pp_cfg::kSmemBytesPP, st, tq, tkv, tctx, tavg, p));
  } else if (heads) CU_TRY(launch_pdl_if(after_kernel, attention_kernel<true>, dim3(grid), dim3(kThreads), kSmemBytes, st, tq, tkv, tctx, tavg, p));
  else if (KP == KP_MAX && full_mode) CU_TRY(launch_pdl_if(after_kernel, attention_kernel<false, true>, dim3(grid), dim3(kThreads), kSmemBytes, st, tq, tkv, tctx, tavg, p));
  else CU_TRY(launch_pdl_if(after_kernel, attention_kernel<false>, dim3(grid), dim3(kThreads), kSmemBytes, st, tq, tkv, tctx, tavg, p));
  if (avg && p.part_images > 0) {
    const long n4 = (long)B * N * pitch / 4;
    CU_TRY(launch_pdl(avg_parts_sum_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, (const float4*)parts, (float4*)avg, n4, S));
  }
  return VITB200_OK;
}

// ------------------------------------------------------------------------------------------ engine
struct LayerWeights {
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  __nv_bfloat16 *w_qkv = nullptr, *w_o = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;
  __nv_bfloat16 *w_qkv_lo = nullptr, *w_o_lo = nullptr, *w_fc1_lo = nullptr, *w_fc2_lo = nullptr;  // fp32x3 mode
  float *b_qkv = nullptr, *b_o = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
  // LayerNorm folding: fp32 originals of the two weights that consume a LayerNorm (w_qkv / w_fc1 above hold the
  // gamma-scaled bf16 versions), their column sums and beta-folded biases
  float *w_qkv_f32 = nullptr, *w_fc1_f32 = nullptr;
  float *s_qkv = nullptr, *s_fc1 = nullptr, *bf_qkv = nullptr, *bf_fc1 = nullptr;
};

struct Buffer {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace vitb200

using namespace vitb200;

struct vitb200_engine {
  vitb200_config cfg{};
  int N = 0, n = 0, D = 0, KP = 0, pitch = 0, patch_k = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  uint64_t launches = 0;

  // weights
  bool precise = false;  // fp32x3 mode: every GEMM / attention operand is carried as hi + lo bf16 (cfg.precision == 1)
  __nv_bfloat16* w_patch = nullptr;
  float* w_patch_f32 = nullptr;   // conv_proj.weight in fp32: the B operand of the TMA-im2col patch embedding (kind::tf32)
  bool patch_tma = false;         // VITB200_PATCH_TMA=1 at creation: patch_embed.cuh instead of patchify + GEMM
  __nv_bfloat16 *w_patch_lo = nullptr, *w_head_lo = nullptr;
  float* b_patch = nullptr;
  float *cls_token = nullptr, *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
  __nv_bfloat16* w_head = nullptr;
  float* b_head = nullptr;
  std::vector<LayerWeights> layers;
  std::map<std::string, bool> loaded;
  size_t expected_tensors = 0;
  bool folded = false;  // fold_ln_weight_kernel has run for the current weights
  Buffer stage_f32;  // fp32 staging for weight upload / conversion

  // activations (sized for cap_batch images)
  int cap_batch = 0;
  uint32_t cap_flags = 0;
  Buffer images, patches, x, xb, ln_stats, qkv, ctx, mlp, cls_ln, logits, avg, cls, heads, hidden, rollout, attn_stats, avg_parts;
  Buffer patches_lo, xb_lo, qkv_lo, ctx_lo, mlp_lo, cls_ln_lo;  // fp32x3 mode: low halves of the bf16 operands

  // vitb200_bind_outputs: caller-owned destinations of the small results (typically slices of rank 0's receive
  // buffer mapped over NVLink).  Honoured by vitb200_forward_device only (`use_bound` is set for its duration).
  float* bound_logits = nullptr;
  float* bound_cls = nullptr;
  float* bound_rollout = nullptr;
  long bound_cls_layer_stride = 0;
  bool use_bound = false;

  // vitb200_set_deferred: node-granular calls that have no host INPUT return without synchronising; their host
  // outputs are valid after vitb200_synchronize (one wait per request instead of four per node)
  bool deferred = false;

  // vitb200_submit_host / vitb200_wait: two requests in flight.  H2D of request i+1 (copy_in stream) and D2H of
  // request i-1 (copy_out stream, from per-slot staging copies of the small outputs) overlap the forward of request i.
  struct Slot {
    Buffer images, logits, cls, rollout, avg;
    cudaEvent_t in_done = nullptr, compute_done = nullptr, out_done = nullptr;
    bool busy = false;
    int batch = 0;
    uint32_t flags = 0;
  } slots[2];
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  uint64_t submitted = 0;  // tickets are submission indices; slot = ticket % 2

  // Deferred node outputs (vitb200_set_deferred) leave on their own stream: in-stream, the next node's kernels queue behind
  // the PCIe copies of this node's outputs (three per layer node, ~33 us of a ~75 us layer at batch 1).
  //   side_ev      main stream -> side stream: "everything enqueued so far has run" (recorded lazily, once per node)
  //   map_ev[l]    side stream -> main stream: the copies of layer l's maps have left; waited for by the NEXT launch that
  //                rewrites layer l's maps (the next request)
  //   tok_ring     the token stream is rewritten in place by the next node: snapshot (D2D) into a small ring, copied out
  //                from there; tok_ev[k] says slot k's previous copy has left
  cudaStream_t side = nullptr;
  cudaEvent_t side_ev = nullptr;
  bool side_stale = true;                 // kernels were enqueued on the main stream since side_ev was last recorded
  std::vector<cudaEvent_t> map_ev;        // [L], created on first use
  std::vector<char> map_busy;
  static constexpr int kTokRing = 4;
  Buffer tok_ring[kTokRing];
  cudaEvent_t tok_ev[kTokRing] = {nullptr, nullptr, nullptr, nullptr};
  bool tok_busy[kTokRing] = {false, false, false, false};
  int tok_next = 0;

  // Bumped whenever an activation buffer is re-allocated (vitb200_workspace_generation): device-resident state a caller
  // believes the engine still holds (token stream, maps, preprocessed images) is gone, and so is every captured graph.
  uint64_t generation = 1;

  // CUDA graphs: a forward (or one node-granular stage) is a fixed launch sequence for a given (entry point, layer,
  // batch, flags, input / bound-output addresses); the second time a key is seen its launches are captured from the
  // stream and replayed from then on with ONE cudaGraphLaunch (the single-image request is ~110 launches whose host-side
  // cost -- cudaLaunchKernelEx plus 2-4 cuTensorMapEncodeTiled calls each -- exceeded their device time).  The tensor
  // maps are kernel parameters, so they are frozen into the graph with everything else.
  struct GraphKey {
    int kind = 0, layer = 0, batch = 0;
    uint32_t flags = 0;
    const void* images = nullptr;
    const void *b_logits = nullptr, *b_cls = nullptr, *b_rollout = nullptr;
    long b_stride = 0;
    bool operator<(const GraphKey& o) const {
      return std::tie(kind, layer, batch, flags, images, b_logits, b_cls, b_rollout, b_stride) <
             std::tie(o.kind, o.layer, o.batch, o.flags, o.images, o.b_logits, o.b_cls, o.b_rollout, o.b_stride);
    }
  };
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0, last_use = 0;
    bool failed = false;
  };
  std::map<GraphKey, GraphEntry> graphs;
  uint64_t graphs_generation = 0, graph_clock = 0, graph_replays = 0;
  bool use_graphs = true;

  // vitb200_profile_forward: an event in front of every launch (only while `profiling`)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<std::string> prof_names;

  ~vitb200_engine();
};

namespace vitb200 {

// `generation` (optional): incremented when an EXISTING allocation is replaced (its contents and address are gone).
static int ensure(Buffer& b, size_t bytes, uint64_t* generation = nullptr) {
  if (b.bytes >= bytes) return VITB200_OK;
  if (b.p) {
    CU_TRY(cudaFree(b.p));
    if (generation) ++*generation;
  }
  b.p = nullptr, b.bytes = 0;
  CU_TRY(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return VITB200_OK;
}

static void release(Buffer& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr, b.bytes = 0;
}

static int ensure_workspace(vitb200_engine* e, int B, uint32_t flags) {
  const vitb200_config& c = e->cfg;
  const size_t M = (size_t)B * e->N;
  const size_t L = c.num_layers;
  uint64_t& gen = e->generation;   // counts replaced allocations (also when a later one fails)
  auto ensure = [&](Buffer& b, size_t bytes) { return vitb200::ensure(b, bytes, &gen); };
  VT_TRY(ensure(e->images, (size_t)B * 3 * c.image_size * c.image_size * 4));
  VT_TRY(ensure(e->patches, (size_t)B * e->n * e->patch_k * 2));
  VT_TRY(ensure(e->x, M * c.hidden_dim * 4));
  VT_TRY(ensure(e->xb, M * c.hidden_dim * 2));
  VT_TRY(ensure(e->ln_stats, M * (c.hidden_dim / 64) * sizeof(float2)));   // sized for the narrow slots (stats_width)
  VT_TRY(ensure(e->qkv, M * 3 * c.hidden_dim * 2));
  VT_TRY(ensure(e->ctx, M * c.hidden_dim * 2));
  if (e->precise || !attention_is_fused(e->N, e->D)) VT_TRY(ensure(e->attn_stats, M * c.num_heads * sizeof(float2)));
  else VT_TRY(ensure(e->avg_parts, attention_parts_bytes(e->N, e->pitch)));   // head-split small launches (launch_attention)
  if (e->precise) {
    VT_TRY(ensure(e->patches_lo, (size_t)B * e->n * e->patch_k * 2));
    VT_TRY(ensure(e->xb_lo, M * c.hidden_dim * 2));
    VT_TRY(ensure(e->qkv_lo, M * 3 * c.hidden_dim * 2));
    VT_TRY(ensure(e->ctx_lo, M * c.hidden_dim * 2));
    VT_TRY(ensure(e->mlp_lo, M * c.mlp_dim * 2));
    VT_TRY(ensure(e->cls_ln_lo, (size_t)B * c.hidden_dim * 2));
  }
  VT_TRY(ensure(e->mlp, M * c.mlp_dim * 2));
  VT_TRY(ensure(e->cls_ln, (size_t)B * c.hidden_dim * 2));
  VT_TRY(ensure(e->logits, (size_t)B * c.num_classes * 4));
  if (flags & (VITB200_EMIT_AVG | VITB200_EMIT_ROLLOUT)) VT_TRY(ensure(e->avg, L * M * e->pitch * 4));
  if (flags & VITB200_EMIT_ROLLOUT) VT_TRY(ensure(e->rollout, (size_t)B * (e->N - 1) * 4));
  if (flags & VITB200_EMIT_CLS) VT_TRY(ensure(e->cls, L * B * c.num_heads * e->N * 4));
  if (flags & VITB200_EMIT_HEADS) VT_TRY(ensure(e->heads, L * M * c.num_heads * e->pitch * 4));
  if (flags & VITB200_EMIT_HIDDEN) VT_TRY(ensure(e->hidden, L * M * c.hidden_dim * 4));
  if (B > e->cap_batch) {
    if (e->cap_batch > 0) ++gen;   // the layer stride of the map buffers changes with the capacity
    e->cap_batch = B;
  }
  e->cap_flags |= flags;
  return VITB200_OK;
}

// In profiling mode: a CUDA event on the launch stream in front of the named kernel.
static void prof_mark(vitb200_engine* e, const char* name, cudaStream_t st) {
  if (!e->profiling) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  e->prof_events.push_back(ev);
  e->prof_names.push_back(name);
}

// ---- forward stages (all enqueue on `st`, no synchronisation) --------------------------------------
// Patch embedding with the im2col done by TMA (patch_embed.cuh): bf16 mode, patch 16, width a multiple of 256.  Opt-in
// (VITB200_PATCH_TMA=1, read when the engine is created): correct and more accurate (tf32 operands: 1.7e-3 against the
// fp32 oracle where the bf16 patch matrix gives 5.5e-3) but SLOWER than patchify + bf16 GEMM on the B200 -- 0.33 against
// 0.23 ms per forward at batch 256: with fp32 operands every k-block moves twice the bytes from L2 into shared memory for
// half the math rate, and the weight tile is re-fetched for every 128-row unit (1.8 GB of L2 -> smem traffic per launch).
static bool patch_embed_fused(const vitb200_engine* e) {
  const vitb200_config& c = e->cfg;
  return e->patch_tma && !e->precise && c.patch_size == 16 && c.hidden_dim % 256 == 0 && e->w_patch_f32 != nullptr &&
         c.image_size / c.patch_size <= 128;
}

static int launch_patch_embed(vitb200_engine* e, const float* images_dev, int B, int sw, cudaStream_t st) {
  using namespace patch_cfg;
  const vitb200_config& c = e->cfg;
  const int S = c.image_size, g = S / 16, d = c.hidden_dim;
  PatchEmbedParams p;
  p.B = B, p.S = S, p.g = g, p.N = e->N, p.d = d;
  p.pr = 128 / g, p.groups = (g + p.pr - 1) / p.pr;
  p.bias = e->b_patch, p.pos = e->pos, p.x = (float*)e->x.p, p.xb = (__nv_bfloat16*)e->xb.p;
  p.row_stats = (float2*)e->ln_stats.p, p.slot_width = sw;
  CUtensorMap ti, tw, tx, txb;
  {   // image [B, 3, S, S] viewed as (kx, px, ky, py, b*3 + c)
    const cuuint64_t dims[5] = {16, (cuuint64_t)g, 16, (cuuint64_t)g, (cuuint64_t)B * 3};
    const cuuint64_t str[4] = {64, (cuuint64_t)S * 4, (cuuint64_t)16 * S * 4, (cuuint64_t)S * S * 4};
    const cuuint32_t box[5] = {16, (cuuint32_t)g, 1, (cuuint32_t)p.pr, 1};
    VT_TRY(make_tmap_nd(&ti, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, images_dev, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {   // conv_proj.weight [d, 3 * 16 * 16] fp32
    const cuuint64_t dims[2] = {(cuuint64_t)e->patch_k, (cuuint64_t)d};
    const cuuint64_t str[1] = {(cuuint64_t)e->patch_k * 4};
    const cuuint32_t box[2] = {BK, BN};
    VT_TRY(make_tmap_nd(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, e->w_patch_f32, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {   // token stream [B][N][d] fp32 and its bf16 copy
    const cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)e->N, (cuuint64_t)B};
    const cuuint64_t str32[2] = {(cuuint64_t)d * 4, (cuuint64_t)e->N * d * 4};
    const cuuint64_t str16[2] = {(cuuint64_t)d * 2, (cuuint64_t)e->N * d * 2};
    const cuuint32_t box[3] = {32, 32, 1};
    VT_TRY(make_tmap_nd(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, e->x.p, 3, dims, str32, box, CU_TENSOR_MAP_SWIZZLE_128B));
    VT_TRY(make_tmap_nd(&txb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, e->xb.p, 3, dims, str16, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  VT_TRY(ensure_func_smem((const void*)patch_embed_kernel, kSmemBytes));
  const int units = B * p.groups * (d / BN), sms = device_sms();
  patch_embed_kernel<<<units < sms ? units : sms, kThreads, kSmemBytes, st>>>(ti, tw, tx, txb, p);
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

static int run_embed(vitb200_engine* e, const float* images_dev, int B, cudaStream_t st) {
  const vitb200_config& c = e->cfg;
  if (patch_embed_fused(e)) {
    const int sw = stats_width(B * e->N, c.hidden_dim);
    prof_mark(e, "gemm_patch_embed", st);
    VT_TRY(launch_patch_embed(e, images_dev, B, sw, st));
    const long cthreads = (long)B * ((c.hidden_dim + 127) / 128) * 32;   // one warp per (image, 128 columns)
    prof_mark(e, "cls_rows", st);
    CU_TRY(launch_pdl(cls_rows_kernel, dim3((unsigned)((cthreads + 255) / 256)), dim3(256), 0, st, e->cls_token, e->pos,
                      (float*)e->x.p, (__nv_bfloat16*)e->xb.p, (float2*)e->ln_stats.p, B, e->N, c.hidden_dim, sw,
                      (__nv_bfloat16*)e->xb_lo.p));
    e->launches += 2;
    return VITB200_OK;
  }
  const long items = (long)B * 3 * c.image_size * (c.image_size / c.patch_size) * (c.patch_size / 8);
  prof_mark(e, "patchify", st);
  patchify_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(images_dev, (__nv_bfloat16*)e->patches.p, B,
                                                                   c.image_size, c.patch_size,
                                                                   (__nv_bfloat16*)e->patches_lo.p);
  CU_TRY(cudaGetLastError());
  GemmEpilogue ep;
  ep.bias = e->b_patch;
  ep.out = e->x.p, ep.ldo = c.hidden_dim;
  ep.resid = e->pos, ep.ldr = c.hidden_dim;
  ep.group_rows = e->n, ep.out_group_stride = e->N, ep.out_row_offset = 1;
  ep.resid_broadcast = 1, ep.resid_row_offset = 1;
  ep.xb = (__nv_bfloat16*)e->xb.p, ep.ldxb = c.hidden_dim;
  const int sw = stats_width(B * e->N, c.hidden_dim);
  ep.row_stats_out = (float2*)e->ln_stats.p, ep.stats_slots = c.hidden_dim / sw;
  ep.xb_lo = (__nv_bfloat16*)e->xb_lo.p;
  prof_mark(e, "gemm_patch_embed", st);
  VT_TRY(launch_gemm(e->patches.p, e->patch_k, e->w_patch, B * e->n, c.hidden_dim, e->patch_k, ep, false, true, st,
                     e->patches_lo.p, e->w_patch_lo));
  const long cthreads = (long)B * ((c.hidden_dim + 127) / 128) * 32;   // one warp per (image, 128 columns)
  prof_mark(e, "cls_rows", st);
  CU_TRY(launch_pdl(cls_rows_kernel, dim3((unsigned)((cthreads + 255) / 256)), dim3(256), 0, st, e->cls_token, e->pos,
                    (float*)e->x.p, (__nv_bfloat16*)e->xb.p, (float2*)e->ln_stats.p, B, e->N, c.hidden_dim, sw,
                    (__nv_bfloat16*)e->xb_lo.p));
  e->launches += 3;
  return VITB200_OK;
}

// First half of an EncoderBlock (vision_transformer.py:112-116): x <- x + out_proj(MHA(LN1 x)), maps as requested.
static int run_attn_block(vitb200_engine* e, int l, int B, uint32_t flags, cudaStream_t st) {
  const vitb200_config& c = e->cfg;
  const LayerWeights& w = e->layers[l];
  const int M = B * e->N, d = c.hidden_dim;
  float* x = (float*)e->x.p;
  __nv_bfloat16* xb = (__nv_bfloat16*)e->xb.p;
  float2* stats = (float2*)e->ln_stats.p;
  const int slots = d / stats_width(M, d);
  {
    GemmEpilogue ep;
    prof_mark(e, "gemm_qkv", st);
    ep.bias = w.bf_qkv, ep.out = e->qkv.p, ep.ldo = 3 * d;
    ep.row_stats_in = stats, ep.stats_in_slots = slots, ep.ln_eps = 1e-6f;
    ep.stats_in_pairs = (d % 256 == 0 && slots == d / 64) ? 1 : 0;
    ep.colsum = w.s_qkv, ep.out_lo = (__nv_bfloat16*)e->qkv_lo.p;
    // the fused attention kernels take V in fp16 (their probabilities are fp16; see attention.cuh)
    if (!e->precise && attention_is_fused(e->N, e->D)) ep.f16_from_col = 2 * d;
    VT_TRY(launch_gemm(xb, d, w.w_qkv, M, 3 * d, d, ep, false, false, st, e->xb_lo.p, w.w_qkv_lo));
  }
  const bool want_avg = (flags & (VITB200_EMIT_AVG | VITB200_EMIT_ROLLOUT)) != 0;
  float* avg = want_avg ? (float*)e->avg.p + (size_t)l * e->cap_batch * e->N * e->pitch : nullptr;
  float* cls = (flags & VITB200_EMIT_CLS) ? (float*)e->cls.p + (size_t)l * e->cap_batch * c.num_heads * e->N : nullptr;
  if (cls && e->use_bound && e->bound_cls) cls = e->bound_cls + (size_t)l * e->bound_cls_layer_stride;
  float* hm = (flags & VITB200_EMIT_HEADS)
                  ? (float*)e->heads.p + (size_t)l * e->cap_batch * c.num_heads * e->N * e->pitch
                  : nullptr;
  prof_mark(e, "attention", st);
  VT_TRY(launch_attention((const __nv_bfloat16*)e->qkv.p, (__nv_bfloat16*)e->ctx.p, avg, cls, hm, B, e->N, c.num_heads,
                          e->D, e->pitch, (float2*)e->attn_stats.p, st, (const __nv_bfloat16*)e->qkv_lo.p,
                          (__nv_bfloat16*)e->ctx_lo.p, (float*)e->avg_parts.p));
  {
    GemmEpilogue ep;
    prof_mark(e, "gemm_out_proj", st);
    ep.bias = w.b_o, ep.out = x, ep.ldo = d, ep.resid = x, ep.ldr = d;
    ep.xb = xb, ep.ldxb = d, ep.row_stats_out = stats, ep.stats_slots = slots, ep.xb_lo = (__nv_bfloat16*)e->xb_lo.p;
    VT_TRY(launch_gemm(e->ctx.p, d, w.w_o, M, d, d, ep, false, true, st, e->ctx_lo.p, w.w_o_lo));
  }
  e->launches += 3;
  return VITB200_OK;
}

// Second half (vision_transformer.py:118-119): x <- x + fc2(GELU(fc1(LN2 x))).
static int run_mlp_block(vitb200_engine* e, int l, int B, cudaStream_t st) {
  const vitb200_config& c = e->cfg;
  const LayerWeights& w = e->layers[l];
  const int M = B * e->N, d = c.hidden_dim;
  float* x = (float*)e->x.p;
  __nv_bfloat16* xb = (__nv_bfloat16*)e->xb.p;
  float2* stats = (float2*)e->ln_stats.p;
  const int slots = d / stats_width(M, d);
  {
    GemmEpilogue ep;
    prof_mark(e, "gemm_fc1_gelu", st);
    ep.bias = w.bf_fc1, ep.out = e->mlp.p, ep.ldo = c.mlp_dim;
    ep.row_stats_in = stats, ep.stats_in_slots = slots, ep.ln_eps = 1e-6f;
    ep.stats_in_pairs = (d % 256 == 0 && slots == d / 64) ? 1 : 0;
    ep.colsum = w.s_fc1, ep.out_lo = (__nv_bfloat16*)e->mlp_lo.p;
    VT_TRY(launch_gemm(xb, d, w.w_fc1, M, c.mlp_dim, d, ep, true, false, st, e->xb_lo.p, w.w_fc1_lo));
  }
  {
    GemmEpilogue ep;
    prof_mark(e, "gemm_fc2", st);
    ep.bias = w.b_fc2, ep.out = x, ep.ldo = d, ep.resid = x, ep.ldr = d;
    ep.xb = xb, ep.ldxb = d, ep.row_stats_out = stats, ep.stats_slots = slots, ep.xb_lo = (__nv_bfloat16*)e->xb_lo.p;
    VT_TRY(launch_gemm(e->mlp.p, c.mlp_dim, w.w_fc2, M, d, c.mlp_dim, ep, false, true, st, e->mlp_lo.p, w.w_fc2_lo));
  }
  e->launches += 2;
  return VITB200_OK;
}

static int run_layer(vitb200_engine* e, int l, int B, uint32_t flags, cudaStream_t st) {
  VT_TRY(run_attn_block(e, l, B, flags, st));
  VT_TRY(run_mlp_block(e, l, B, st));
  if (flags & VITB200_EMIT_HIDDEN) {
    const size_t n = (size_t)B * e->N * e->cfg.hidden_dim;
    float* hid = (float*)e->hidden.p + (size_t)l * e->cap_batch * e->N * e->cfg.hidden_dim;
    CU_TRY(cudaMemcpyAsync(hid, e->x.p, n * 4, cudaMemcpyDeviceToDevice, st));
  }
  return VITB200_OK;
}

static int run_head(vitb200_engine* e, int B, cudaStream_t st) {
  const vitb200_config& c = e->cfg;
  const int d = c.hidden_dim;
  prof_mark(e, "layernorm_final", st);
  VT_TRY(launch_layernorm((const float*)e->x.p, (long)e->N * d, e->lnf_g, e->lnf_b, (__nv_bfloat16*)e->cls_ln.p, B, d,
                          1e-6f, st, (__nv_bfloat16*)e->cls_ln_lo.p));
  GemmEpilogue ep;
  ep.bias = e->b_head, ep.out = (e->use_bound && e->bound_logits) ? (void*)e->bound_logits : e->logits.p, ep.ldo = c.num_classes;
  prof_mark(e, "gemm_head", st);
  VT_TRY(launch_gemm(e->cls_ln.p, d, e->w_head, B, c.num_classes, d, ep, false, true, st, e->cls_ln_lo.p, e->w_head_lo));
  e->launches += 2;
  return VITB200_OK;
}

// stages: as many as fit two CTAs per SM (the copies in flight are what saturates HBM)
static int launch_rollout(const float* maps, long layer_stride, int L, int B, int N, int ld, float* out, cudaStream_t st) {
  if (N > kRolloutThreads * kRolloutMaxCols) return fail(VITB200_ERR_INVALID, "rollout: at most %d tokens", kRolloutThreads * kRolloutMaxCols);
  if (ld % 4 != 0 || ld < N) return fail(VITB200_ERR_INVALID, "rollout: pitch %d must be >= %d and a multiple of 4", ld, N);
  int stages = 8;
  // Small batches: a cluster of C CTAs per image (rollout_cluster_kernel), so that a single-image request or a 16-image
  // ViT-H batch does not run its rollout on 1 / 16 of 148 SMs.  VITB200_ROLLOUT_CLUSTER=0 keeps one CTA per image.
  static const int cluster_mode = [] {
    const char* v = getenv("VITB200_ROLLOUT_CLUSTER");
    return (v && v[0] == '0') ? 0 : 1;
  }();
  int C = 1;
  if (cluster_mode) while (C < kRolloutMaxCluster && 2 * C * B <= device_sms()) C *= 2;
  if (C > 1) {
    while (stages > 2 && rollout_cluster_smem_bytes(ld, stages, C) > 110 * 1024) --stages;
    const int smem = rollout_cluster_smem_bytes(ld, stages, C);
    VT_TRY(ensure_func_smem((const void*)rollout_cluster_kernel, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * C), cfg.blockDim = dim3(kRolloutThreads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 2 : 1;
    CU_TRY(cudaLaunchKernelEx(&cfg, rollout_cluster_kernel, maps, layer_stride, L, N, ld, stages, C, out));
    return VITB200_OK;
  }
  static const int warp_mode = [] {   // VITB200_ROLLOUT_WARP=0: the block-synchronous kernel
    const char* v = getenv("VITB200_ROLLOUT_WARP");
    return (v && v[0] == '0') ? 0 : 1;
  }();
  if (warp_mode) {
    while (stages > 2 && rollout_warp_smem_bytes(ld, stages) > 110 * 1024) --stages;
    const int smem = rollout_warp_smem_bytes(ld, stages);
    VT_TRY(ensure_func_smem((const void*)rollout_warp_kernel, smem));
    CU_TRY(launch_pdl(rollout_warp_kernel, dim3(B), dim3(kRolloutWarpThreads), smem, st, maps, layer_stride, L, N, ld, stages, out));
    return VITB200_OK;
  }
  while (stages > 2 && rollout_smem_bytes(ld, stages) > 110 * 1024) --stages;
  const int smem = rollout_smem_bytes(ld, stages);
  VT_TRY(ensure_func_smem((const void*)rollout_cls_kernel, smem));
  CU_TRY(launch_pdl(rollout_cls_kernel, dim3(B), dim3(kRolloutThreads), smem, st, maps, layer_stride, L, N, ld, stages, out));
  return VITB200_OK;
}

static int run_rollout(vitb200_engine* e, int B, cudaStream_t st) {
  prof_mark(e, "rollout", st);
  VT_TRY(launch_rollout((const float*)e->avg.p, (long)e->cap_batch * e->N * e->pitch, e->cfg.num_layers, B, e->N, e->pitch,
                        (e->use_bound && e->bound_rollout) ? e->bound_rollout : (float*)e->rollout.p, st));
  e->launches += 1;
  return VITB200_OK;
}

static int check_ready(vitb200_engine* e) {
  if (e->loaded.size() != e->expected_tensors)
    return fail(VITB200_ERR_STATE, "weights incomplete: %zu of %zu tensors loaded", e->loaded.size(), e->expected_tensors);
  if (!e->folded) {
    // LayerNorm folding, weight side (see the header comment): once per set of weights, on the engine's stream
    const int d = e->cfg.hidden_dim, mlp = e->cfg.mlp_dim;
    for (LayerWeights& w : e->layers) {
      auto alloc = [](auto** p, size_t bytes) { return *p ? cudaSuccess : cudaMalloc((void**)p, bytes); };
      CU_TRY(alloc(&w.w_qkv, (size_t)3 * d * d * 2));
      CU_TRY(alloc(&w.s_qkv, (size_t)3 * d * 4));
      CU_TRY(alloc(&w.bf_qkv, (size_t)3 * d * 4));
      CU_TRY(alloc(&w.w_fc1, (size_t)mlp * d * 2));
      CU_TRY(alloc(&w.s_fc1, (size_t)mlp * 4));
      CU_TRY(alloc(&w.bf_fc1, (size_t)mlp * 4));
      if (e->precise) {
        CU_TRY(alloc(&w.w_qkv_lo, (size_t)3 * d * d * 2));
        CU_TRY(alloc(&w.w_fc1_lo, (size_t)mlp * d * 2));
      }
      fold_ln_weight_kernel<<<(3 * d * 32 + 255) / 256, 256, 0, e->stream>>>(w.w_qkv_f32, w.ln1_g, w.ln1_b, w.b_qkv, w.w_qkv,
                                                                             w.s_qkv, w.bf_qkv, 3 * d, d, w.w_qkv_lo);
      fold_ln_weight_kernel<<<(mlp * 32 + 255) / 256, 256, 0, e->stream>>>(w.w_fc1_f32, w.ln2_g, w.ln2_b, w.b_fc1, w.w_fc1,
                                                                           w.s_fc1, w.bf_fc1, mlp, d, w.w_fc1_lo);
      CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->folded = true;
  }
  return VITB200_OK;
}

static int check_batch(vitb200_engine* e, int B) {
  if (B <= 0) return fail(VITB200_ERR_INVALID, "batch must be positive (got %d)", B);
  (void)e;
  return VITB200_OK;
}

static void clear_graphs(vitb200_engine* e) {
  for (auto& kv : e->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
}

// Run `body` (which only ENQUEUES work on `st`) directly, or -- from the second time `key` is seen -- as a captured
// CUDA graph.  `body` must be a pure function of the key, the engine's configuration and its buffers (whose
// re-allocation bumps `generation` and drops every graph).
template <class Body>
static int run_graphed(vitb200_engine* e, const vitb200_engine::GraphKey& key, cudaStream_t st, Body&& body) {
  e->side_stale = true;   // kernels go onto the main stream: the side stream has to follow it again before its next copy
  // (the legacy default stream cannot be captured; profiling needs an event in front of every launch)
  if (!e->use_graphs || e->profiling || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) return body();
  if (e->graphs_generation != e->generation) {
    clear_graphs(e);
    e->graphs_generation = e->generation;
  }
  auto it = e->graphs.find(key);
  if (it == e->graphs.end()) {
    if (e->graphs.size() >= 128) {   // evict the least recently used entry
      auto lru = e->graphs.begin();
      for (auto j = e->graphs.begin(); j != e->graphs.end(); ++j)
        if (j->second.last_use < lru->second.last_use) lru = j;
      if (lru->second.exec) cudaGraphExecDestroy(lru->second.exec);
      e->graphs.erase(lru);
    }
    e->graphs[key].last_use = ++e->graph_clock;   // first sighting: run eagerly (this also configures the kernels)
    return body();
  }
  vitb200_engine::GraphEntry& g = it->second;
  g.last_use = ++e->graph_clock;
  if (g.failed) return body();
  if (g.exec == nullptr) {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      g.failed = true;
      return body();
    }
    const uint64_t l0 = e->launches;
    const int rc = body();
    cudaGraph_t graph = nullptr;
    const cudaError_t err = cudaStreamEndCapture(st, &graph);
    g.launches = e->launches - l0;
    e->launches = l0;
    bool ok = rc == VITB200_OK && err == cudaSuccess && graph != nullptr;
    if (ok && cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) g.exec = nullptr, ok = false;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      g.failed = true;
      return rc != VITB200_OK ? rc : body();
    }
  }
  CU_TRY(cudaGraphLaunch(g.exec, st));
  e->launches += g.launches;
  e->graph_replays += 1;
  return VITB200_OK;
}

static vitb200_engine::GraphKey graph_key(vitb200_engine* e, int kind, int layer, int B, uint32_t flags, const void* images) {
  vitb200_engine::GraphKey k;
  k.kind = kind, k.layer = layer, k.batch = B, k.flags = flags, k.images = images;
  if (e->use_bound) {
    k.b_logits = e->bound_logits, k.b_cls = e->bound_cls, k.b_rollout = e->bound_rollout;
    k.b_stride = e->bound_cls_layer_stride;
  }
  return k;
}
enum { kGraphForward = 0, kGraphEmbed, kGraphLayer, kGraphAttnBlock, kGraphMlpBlock, kGraphHead, kGraphRollout };

static int forward_stages(vitb200_engine* e, const float* images_dev, int B, uint32_t flags, cudaStream_t st);

static int forward_device_locked(vitb200_engine* e, const float* images_dev, int B, uint32_t flags, cudaStream_t st) {
  VT_TRY(check_ready(e));
  VT_TRY(check_batch(e, B));
  // the layer-strided map buffers are addressed with cap_batch: grow first, then never shrink
  VT_TRY(ensure_workspace(e, B > e->cap_batch ? B : e->cap_batch, flags | e->cap_flags));
  if (e->side)   // a whole forward rewrites every layer's maps: side-stream copies of node outputs must have left
    for (int l = 0; l < e->cfg.num_layers; ++l)
      if (e->map_busy[l]) {
        CU_TRY(cudaStreamWaitEvent(st, e->map_ev[l], 0));
        e->map_busy[l] = 0;
      }
  return run_graphed(e, graph_key(e, kGraphForward, 0, B, flags, images_dev), st,
                     [&] { return forward_stages(e, images_dev, B, flags, st); });
}

static int forward_stages(vitb200_engine* e, const float* images_dev, int B, uint32_t flags, cudaStream_t st) {
  VT_TRY(run_embed(e, images_dev, B, st));
  for (int l = 0; l < e->cfg.num_layers; ++l) VT_TRY(run_layer(e, l, B, flags, st));
  VT_TRY(run_head(e, B, st));
  if (flags & VITB200_EMIT_ROLLOUT) VT_TRY(run_rollout(e, B, st));
  return VITB200_OK;
}

// dense host <- pitched device rows
static int copy_rows_to_host(float* dst, const float* src_dev, size_t rows, int width, int pitch, cudaStream_t st) {
  CU_TRY(cudaMemcpy2DAsync(dst, (size_t)width * 4, src_dev, (size_t)pitch * 4, (size_t)width * 4, rows,
                           cudaMemcpyDeviceToHost, st));
  return VITB200_OK;
}

}  // namespace vitb200

vitb200_engine::~vitb200_engine() {
  clear_graphs(this);
  Buffer* bufs[] = {&stage_f32, &images, &patches, &x, &xb, &ln_stats, &qkv, &ctx, &mlp, &cls_ln, &logits, &avg, &cls, &heads, &hidden, &rollout, &attn_stats, &avg_parts,
                    &patches_lo, &xb_lo, &qkv_lo, &ctx_lo, &mlp_lo, &cls_ln_lo};
  for (Buffer* b : bufs) release(*b);
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(w_patch_lo), fr(w_head_lo), fr(w_patch_f32);
  fr(w_patch), fr(b_patch), fr(cls_token), fr(pos), fr(lnf_g), fr(lnf_b), fr(w_head), fr(b_head);
  for (auto& l : layers) {
    fr(l.ln1_g), fr(l.ln1_b), fr(l.ln2_g), fr(l.ln2_b), fr(l.w_qkv), fr(l.w_o), fr(l.w_fc1), fr(l.w_fc2);
    fr(l.b_qkv), fr(l.b_o), fr(l.b_fc1), fr(l.b_fc2);
    fr(l.w_qkv_lo), fr(l.w_o_lo), fr(l.w_fc1_lo), fr(l.w_fc2_lo);
    fr(l.w_qkv_f32), fr(l.w_fc1_f32), fr(l.s_qkv), fr(l.s_fc1), fr(l.bf_qkv), fr(l.bf_fc1);
  }
  for (Slot& sl : slots) {
    Buffer* sb[] = {&sl.images, &sl.logits, &sl.cls, &sl.rollout, &sl.avg};
    for (Buffer* b : sb) release(*b);
    if (sl.in_done) cudaEventDestroy(sl.in_done);
    if (sl.compute_done) cudaEventDestroy(sl.compute_done);
    if (sl.out_done) cudaEventDestroy(sl.out_done);
  }
  if (copy_in) cudaStreamDestroy(copy_in);
  if (copy_out) cudaStreamDestroy(copy_out);
  if (side) cudaStreamDestroy(side);
  if (side_ev) cudaEventDestroy(side_ev);
  for (cudaEvent_t ev : map_ev) if (ev) cudaEventDestroy(ev);
  for (int k = 0; k < kTokRing; ++k) {
    if (tok_ev[k]) cudaEventDestroy(tok_ev[k]);
    release(tok_ring[k]);
  }
  if (stream) cudaStreamDestroy(stream);
}

// =========================================================================================== C ABI
extern "C" {

static int preprocess_params(int B, int H, int W, int resize, int crop, PreprocessParams* p);

const char* vitb200_last_error(void) { return g_last_error.c_str(); }
int vitb200_version(void) { return 1; }

int vitb200_create(const vitb200_config* cfg, vitb200_engine** out) {
  if (!cfg || !out) return fail(VITB200_ERR_INVALID, "null argument");
  *out = nullptr;
  const vitb200_config& c = *cfg;
  if (c.patch_size <= 0 || c.patch_size % 8 != 0 || c.image_size % c.patch_size != 0)
    return fail(VITB200_ERR_INVALID, "image_size %d must be a multiple of patch_size %d (itself a multiple of 8)",
                c.image_size, c.patch_size);
  if (c.num_heads <= 0 || c.hidden_dim % c.num_heads != 0)
    return fail(VITB200_ERR_INVALID, "hidden_dim %d must be a multiple of the head count %d", c.hidden_dim, c.num_heads);
  const int hd = c.hidden_dim / c.num_heads;
  if (hd < 64 || hd > 128 || hd % 16 != 0)
    return fail(VITB200_ERR_INVALID, "head dim %d unsupported (64..128 in steps of 16)", hd);
  if (c.hidden_dim % 128 != 0 || c.mlp_dim % 64 != 0 || c.num_classes % 8 != 0 || c.num_layers <= 0)
    return fail(VITB200_ERR_INVALID, "unsupported widths: hidden %d mlp %d classes %d layers %d", c.hidden_dim, c.mlp_dim,
                c.num_classes, c.num_layers);
  const int n = (c.image_size / c.patch_size) * (c.image_size / c.patch_size);
  const int N = n + 1;
  const int KP = attention_pitch(N);
  if (N > kRolloutThreads * kRolloutMaxCols)
    return fail(VITB200_ERR_INVALID, "%d tokens per image exceed the engine's limit (%d)", N, kRolloutThreads * kRolloutMaxCols);
  if (c.precision != 0 && c.precision != 1) return fail(VITB200_ERR_INVALID, "precision must be 0 (bf16) or 1 (fp32x3), got %d", c.precision);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(VITB200_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  if (c.device < 0 || c.device >= ndev) return fail(VITB200_ERR_INVALID, "device %d out of range (%d devices)", c.device, ndev);
  CU_TRY(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, c.device));
  if (prop.major != 10) return fail(VITB200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", c.device, prop.major, prop.minor);

  vitb200_engine* e = new vitb200_engine();
  e->cfg = c;
  e->precise = c.precision == 1;
  {
    const char* v = getenv("VITB200_PATCH_TMA");
    e->patch_tma = v && v[0] == '1';
  }
  e->n = n, e->N = N, e->D = hd, e->KP = KP, e->pitch = KP, e->patch_k = 3 * c.patch_size * c.patch_size;
  e->layers.resize(c.num_layers);
  {
    const char* v = getenv("VITB200_GRAPHS");   // VITB200_GRAPHS=0: every call launches its kernels one by one
    e->use_graphs = !(v && v[0] == '0');
  }
  e->expected_tensors = 4 + 12 * (size_t)c.num_layers + 4;
  cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) {
    delete e;
    return fail(VITB200_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(err));
  }
  if (c.max_batch > 0) {
    int s = ensure_workspace(e, c.max_batch, 0);
    if (s != VITB200_OK) {
      delete e;
      return s;
    }
  }
  *out = e;
  return VITB200_OK;
}

void vitb200_destroy(vitb200_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  delete e;
}

int vitb200_load_weight(vitb200_engine* e, const char* name, const float* data_host, size_t count) {
  if (!e || !name || !data_host) return fail(VITB200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU_TRY(cudaSetDevice(e->cfg.device));
  const vitb200_config& c = e->cfg;
  const size_t d = c.hidden_dim;
  const std::string key(name);
  void** slot = nullptr;
  void** slot_lo = nullptr;  // fp32x3 mode: where the low halves of a bf16 weight go
  size_t expect = 0;
  bool to_bf16 = false;
  auto F = [&](float** p, size_t n) { slot = (void**)p, expect = n, to_bf16 = false; };
  auto H = [&](__nv_bfloat16** p, size_t n, __nv_bfloat16** plo) {
    slot = (void**)p, expect = n, to_bf16 = true, slot_lo = (void**)plo;
  };
  if (key == "conv_proj.weight") H(&e->w_patch, d * e->patch_k, &e->w_patch_lo);
  else if (key == "conv_proj.bias") F(&e->b_patch, d);
  else if (key == "class_token") F(&e->cls_token, d);
  else if (key == "encoder.pos_embedding") F(&e->pos, (size_t)e->N * d);
  else if (key == "encoder.ln.weight") F(&e->lnf_g, d);
  else if (key == "encoder.ln.bias") F(&e->lnf_b, d);
  else if (key == "heads.head.weight") H(&e->w_head, (size_t)c.num_classes * d, &e->w_head_lo);
  else if (key == "heads.head.bias") F(&e->b_head, c.num_classes);
  else {
    int li = -1;
    char rest[128] = {0};
    if (sscanf(name, "encoder.layers.encoder_layer_%d.%127s", &li, rest) != 2 || li < 0 || li >= c.num_layers)
      return fail(VITB200_ERR_INVALID, "unknown weight name '%s'", name);
    LayerWeights& w = e->layers[li];
    const std::string r(rest);
    if (r == "ln_1.weight") F(&w.ln1_g, d);
    else if (r == "ln_1.bias") F(&w.ln1_b, d);
    else if (r == "ln_2.weight") F(&w.ln2_g, d);
    else if (r == "ln_2.bias") F(&w.ln2_b, d);
    else if (r == "self_attention.in_proj_weight") F(&w.w_qkv_f32, 3 * d * d);   // folded with ln_1 later
    else if (r == "self_attention.in_proj_bias") F(&w.b_qkv, 3 * d);
    else if (r == "self_attention.out_proj.weight") H(&w.w_o, d * d, &w.w_o_lo);
    else if (r == "self_attention.out_proj.bias") F(&w.b_o, d);
    else if (r == "mlp.0.weight") F(&w.w_fc1_f32, (size_t)c.mlp_dim * d);        // folded with ln_2 later
    else if (r == "mlp.0.bias") F(&w.b_fc1, c.mlp_dim);
    else if (r == "mlp.3.weight") H(&w.w_fc2, d * (size_t)c.mlp_dim, &w.w_fc2_lo);
    else if (r == "mlp.3.bias") F(&w.b_fc2, d);
    else return fail(VITB200_ERR_INVALID, "unknown weight name '%s'", name);
  }
  if (count != expect)
    return fail(VITB200_ERR_INVALID, "weight '%s': expected %zu values, got %zu", name, expect, count);
  if (*slot) {
    CU_TRY(cudaFree(*slot));
    *slot = nullptr;
  }
  if (!to_bf16) {
    CU_TRY(cudaMalloc(slot, count * 4));
    CU_TRY(cudaMemcpyAsync(*slot, data_host, count * 4, cudaMemcpyHostToDevice, e->stream));
  } else {
    VT_TRY(ensure(e->stage_f32, count * 4));
    CU_TRY(cudaMalloc(slot, count * 2));
    CU_TRY(cudaMemcpyAsync(e->stage_f32.p, data_host, count * 4, cudaMemcpyHostToDevice, e->stream));
    if (e->precise) {
      if (*slot_lo) CU_TRY(cudaFree(*slot_lo));
      *slot_lo = nullptr;
      CU_TRY(cudaMalloc(slot_lo, count * 2));
      f32_to_bf16_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, e->stream>>>(
          (const float*)e->stage_f32.p, (__nv_bfloat16*)*slot, (__nv_bfloat16*)*slot_lo, (long)count);
    } else {
      f32_to_bf16_kernel<<<(unsigned)((count / 4 + 256) / 256), 256, 0, e->stream>>>((const float*)e->stage_f32.p,
                                                                                    (__nv_bfloat16*)*slot, (long)count);
    }
    CU_TRY(cudaGetLastError());
    if (key == "conv_proj.weight") {
      if (e->w_patch_f32) CU_TRY(cudaFree(e->w_patch_f32));
      e->w_patch_f32 = nullptr;
      CU_TRY(cudaMalloc(&e->w_patch_f32, count * 4));
      CU_TRY(cudaMemcpyAsync(e->w_patch_f32, e->stage_f32.p, count * 4, cudaMemcpyDeviceToDevice, e->stream));
    }
  }
  CU_TRY(cudaStreamSynchronize(e->stream));
  e->loaded[key] = true;
  e->folded = false;  // any new tensor invalidates the folded LayerNorm weights
  clear_graphs(e);    // ... and every captured launch sequence (weight buffers are re-allocated)
  return VITB200_OK;
}

int vitb200_weights_ready(vitb200_engine* e) {
  if (!e) return fail(VITB200_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);   // check_ready allocates and launches the fold kernels on first use
  CU_TRY(cudaSetDevice(e->cfg.device));
  return check_ready(e);
}

int vitb200_forward_device(vitb200_engine* e, const float* images_dev, int batch, uint32_t flags, void* stream) {
  if (!e || !images_dev) return fail(VITB200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU_TRY(cudaSetDevice(e->cfg.device));
  // `stream` is taken literally: NULL is the legacy default stream, exactly as in the op_* entry points (torch's
  // default stream has the handle 0 too; round 1 silently substituted the engine's private stream for it, which broke
  // every event the caller recorded on "the stream the forward runs on").  vitb200_engine_stream names the private one.
  cudaStream_t st = (cudaStream_t)stream;
  e->use_bound = true;
  const int rc = forward_device_locked(e, images_dev, batch, flags, st);
  e->use_bound = false;
  return rc;
}

int vitb200_bind_outputs(vitb200_engine* e, float* logits_dev, float* cls_dev, long cls_layer_stride, float* rollout_dev) {
  if (!e) return fail(VITB200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  if (cls_dev && cls_layer_stride < (long)e->cfg.num_heads * e->N)
    return fail(VITB200_ERR_INVALID, "bind_outputs: CLS layer stride %ld is smaller than one image (%ld floats)",
                cls_layer_stride, (long)e->cfg.num_heads * e->N);
  // the head GEMM stores float4; the CLS-row writer and the rollout kernel store single floats
  if (((uintptr_t)logits_dev & 15u) || (((uintptr_t)cls_dev | (uintptr_t)rollout_dev) & 3u))
    return fail(VITB200_ERR_INVALID, "bind_outputs: logits must be 16-byte aligned, CLS maps and rollout 4-byte aligned");
  e->bound_logits = logits_dev, e->bound_cls = cls_dev, e->bound_rollout = rollout_dev;
  e->bound_cls_layer_stride = cls_dev ? cls_layer_stride : 0;
  return VITB200_OK;
}

// ---- peer memory + completion flags (csrc/peer.cuh) ---------------------------------------------------
int vitb200_peer_alloc(int device, size_t bytes, void** ptr_dev, void* handle64) {
  if (!ptr_dev || !handle64 || bytes == 0) return fail(VITB200_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU_TRY(cudaSetDevice(device));
  void* p = nullptr;
  CU_TRY(cudaMalloc(&p, bytes));
  cudaError_t err = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (err == cudaSuccess) err = cudaIpcGetMemHandle(&h, p);
  if (err != cudaSuccess) {
    cudaFree(p);
    return fail(VITB200_ERR_CUDA, "peer_alloc: %s", cudaGetErrorString(err));
  }
  memcpy(handle64, &h, 64);
  *ptr_dev = p;
  return VITB200_OK;
}

int vitb200_peer_open(int device, const void* handle64, void** ptr_dev) {
  if (!ptr_dev || !handle64) return fail(VITB200_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_dev = p;
  return VITB200_OK;
}

int vitb200_peer_close(int device, void* ptr_dev) {
  if (!ptr_dev) return VITB200_OK;
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaIpcCloseMemHandle(ptr_dev));
  return VITB200_OK;
}

int vitb200_peer_free(int device, void* ptr_dev) {
  if (!ptr_dev) return VITB200_OK;
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaFree(ptr_dev));
  return VITB200_OK;
}

int vitb200_flag_signal(void* flag_dev, uint32_t value, void* stream) {
  if (!flag_dev || ((uintptr_t)flag_dev & 3u)) return fail(VITB200_ERR_INVALID, "flag_signal: bad flag address");
  flag_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((uint32_t*)flag_dev, value);
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

int vitb200_flag_wait(const void* flag_dev, uint32_t value, void* stream) {
  if (!flag_dev || ((uintptr_t)flag_dev & 3u)) return fail(VITB200_ERR_INVALID, "flag_wait: bad flag address");
  flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const uint32_t*)flag_dev, value, 30ull * 1000 * 1000 * 1000);
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

int vitb200_engine_stream(vitb200_engine* e, void** stream) {
  if (!e || !stream) return fail(VITB200_ERR_INVALID, "null argument");
  *stream = (void*)e->stream;
  return VITB200_OK;
}

uint64_t vitb200_workspace_generation(vitb200_engine* e) {
  if (!e) return 0;
  std::lock_guard<std::mutex> lock(e->mu);
  return e->generation;
}

int vitb200_reserve(vitb200_engine* e, int batch, uint32_t flags) {
  if (!e) return fail(VITB200_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  CU_TRY(cudaSetDevice(e->cfg.device));
  VT_TRY(check_batch(e, batch));
  return ensure_workspace(e, batch > e->cap_batch ? batch : e->cap_batch, flags | e->cap_flags);
}

int vitb200_set_graphs(vitb200_engine* e, int on) {
  if (!e) return fail(VITB200_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  e->use_graphs = on != 0;
  if (!e->use_graphs) {
    CU_TRY(cudaSetDevice(e->cfg.device));
    clear_graphs(e);
  }
  return VITB200_OK;
}

uint64_t vitb200_graph_replays(vitb200_engine* e) {
  if (!e) return 0;
  std::lock_guard<std::mutex> lock(e->mu);
  return e->graph_replays;
}

int vitb200_profile_forward(vitb200_engine* e, const float* images_dev, int batch, uint32_t flags, char* report,
                            size_t report_cap) {
  if (!e || !images_dev || !report || report_cap == 0) return fail(VITB200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t st = e->stream;
  e->profiling = true;
  int rc = forward_device_locked(e, images_dev, batch, flags, st);
  prof_mark(e, "end", st);
  e->profiling = false;
  cudaError_t err = cudaStreamSynchronize(st);
  // one line per kernel kind: name,launches,total_ms  (event-to-event, i.e. including the gap to the next launch)
  std::map<std::string, std::pair<int, double>> agg;
  std::vector<std::string> order;
  double total = 0;
  for (size_t i = 0; rc == VITB200_OK && err == cudaSuccess && i + 1 < e->prof_events.size(); ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e->prof_events[i], e->prof_events[i + 1]);
    if (!agg.count(e->prof_names[i])) order.push_back(e->prof_names[i]);
    agg[e->prof_names[i]].first += 1, agg[e->prof_names[i]].second += ms, total += ms;
  }
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  e->prof_events.clear(), e->prof_names.clear();
  if (rc != VITB200_OK) return rc;
  if (err != cudaSuccess) return fail(VITB200_ERR_CUDA, "profile_forward: %s", cudaGetErrorString(err));
  std::string out;
  char line[160];
  for (const std::string& n : order) {
    snprintf(line, sizeof(line), "%s,%d,%.6f\n", n.c_str(), agg[n].first, agg[n].second);
    out += line;
  }
  snprintf(line, sizeof(line), "total,%zu,%.6f\n", order.size(), total);
  out += line;
  if (out.size() + 1 > report_cap) return fail(VITB200_ERR_INVALID, "profile report needs %zu bytes", out.size() + 1);
  memcpy(report, out.c_str(), out.size() + 1);
  return VITB200_OK;
}

int vitb200_forward_host(vitb200_engine* e, const float* images_host, int batch, uint32_t flags,
                         const vitb200_host_outputs* out) {
  if (!e || !images_host || !out) return fail(VITB200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU_TRY(cudaSetDevice(e->cfg.device));
  const vitb200_config& c = e->cfg;
  if (out->avg_maps) flags |= VITB200_EMIT_AVG;
  if (out->cls_maps) flags |= VITB200_EMIT_CLS;
  if (out->rollout) flags |= VITB200_EMIT_ROLLOUT;
  if (out->heads) flags |= VITB200_EMIT_HEADS;
  if (out->hidden) flags |= VITB200_EMIT_HIDDEN;
  VT_TRY(check_batch(e, batch));
  VT_TRY(ensure_workspace(e, batch > e->cap_batch ? batch : e->cap_batch, flags | e->cap_flags));
  cudaStream_t st = e->stream;
  const size_t img_bytes = (size_t)batch * 3 * c.image_size * c.image_size * 4;
  CU_TRY(cudaMemcpyAsync(e->images.p, images_host, img_bytes, cudaMemcpyHostToDevice, st));
  VT_TRY(forward_device_locked(e, (const float*)e->images.p, batch, flags, st));
  const int B = batch, N = e->N, L = c.num_layers, Hh = c.num_heads, d = c.hidden_dim;
  if (out->logits)
    CU_TRY(cudaMemcpyAsync(out->logits, e->logits.p, (size_t)B * c.num_classes * 4, cudaMemcpyDeviceToHost, st));
  for (int l = 0; l < L; ++l) {
    if (out->avg_maps)
      VT_TRY(copy_rows_to_host(out->avg_maps + (size_t)l * B * N * N, (const float*)e->avg.p + (size_t)l * e->cap_batch * N * e->pitch,
                               (size_t)B * N, N, e->pitch, st));
    if (out->cls_maps)
      CU_TRY(cudaMemcpyAsync(out->cls_maps + (size_t)l * B * Hh * N, (const float*)e->cls.p + (size_t)l * e->cap_batch * Hh * N,
                             (size_t)B * Hh * N * 4, cudaMemcpyDeviceToHost, st));
    if (out->heads)
      VT_TRY(copy_rows_to_host(out->heads + (size_t)l * B * Hh * N * N,
                               (const float*)e->heads.p + (size_t)l * e->cap_batch * Hh * N * e->pitch, (size_t)B * Hh * N, N,
                               e->pitch, st));
    if (out->hidden)
      CU_TRY(cudaMemcpyAsync(out->hidden + (size_t)l * B * N * d, (const float*)e->hidden.p + (size_t)l * e->cap_batch * N * d,
                             (size_t)B * N * d * 4, cudaMemcpyDeviceToHost, st));
  }
  if (out->rollout)
    CU_TRY(cudaMemcpyAsync(out->rollout, e->rollout.p, (size_t)B * (N - 1) * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

static int wait_slot_locked(vitb200_engine* e, vitb200_engine::Slot& sl) {
  if (!sl.busy) return VITB200_OK;
  CU_TRY(cudaEventSynchronize(sl.out_done));
  sl.busy = false;
  return VITB200_OK;
}

int vitb200_submit_host(vitb200_engine* e, const float* images_host, int batch, uint32_t flags,
                        const vitb200_host_outputs* out, uint64_t* ticket) {
  if (!e || !images_host || !out || !ticket) return fail(VITB200_ERR_INVALID, "null argument");
  if (out->heads || out->hidden) return fail(VITB200_ERR_INVALID, "submit_host: per-head maps / hidden states are only available synchronously");
  std::lock_guard<std::mutex> lock(e->mu);
  CU_TRY(cudaSetDevice(e->cfg.device));
  const vitb200_config& c = e->cfg;
  if (out->avg_maps) flags |= VITB200_EMIT_AVG;
  if (out->cls_maps) flags |= VITB200_EMIT_CLS;
  if (out->rollout) flags |= VITB200_EMIT_ROLLOUT;
  VT_TRY(check_ready(e));
  VT_TRY(check_batch(e, batch));
  VT_TRY(ensure_workspace(e, batch > e->cap_batch ? batch : e->cap_batch, flags | e->cap_flags));
  // Bound outputs (vitb200_bind_outputs): the producing kernels store them into the caller's memory (rank 0's receive
  // set in the multi-GPU case), so they are neither staged nor copied to THIS rank's host
  const bool b_logits = e->bound_logits != nullptr, b_cls = e->bound_cls != nullptr && (flags & VITB200_EMIT_CLS),
             b_roll = e->bound_rollout != nullptr && (flags & VITB200_EMIT_ROLLOUT);
  if ((b_logits && out->logits) || (b_cls && out->cls_maps) || (b_roll && out->rollout))
    return fail(VITB200_ERR_INVALID, "submit_host: an output that is bound to caller memory cannot also be copied to the host");
  if (!e->copy_in) {
    CU_TRY(cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking));
    for (auto& sl : e->slots) {
      CU_TRY(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
    }
  }
  vitb200_engine::Slot& sl = e->slots[e->submitted & 1];
  VT_TRY(wait_slot_locked(e, sl));  // third request in flight: the oldest one has to drain first
  const int B = batch, N = e->N, L = c.num_layers, Hh = c.num_heads;
  const size_t img_bytes = (size_t)B * 3 * c.image_size * c.image_size * 4;
  const size_t logit_bytes = (size_t)B * c.num_classes * 4, cls_bytes = (size_t)B * Hh * N * 4;
  VT_TRY(ensure(sl.images, img_bytes));
  VT_TRY(ensure(sl.logits, logit_bytes));
  if (out->cls_maps) VT_TRY(ensure(sl.cls, cls_bytes * L));
  if (out->rollout) VT_TRY(ensure(sl.rollout, (size_t)B * (N - 1) * 4));
  if (out->avg_maps) VT_TRY(ensure(sl.avg, (size_t)L * B * N * N * 4));
  // H2D on the copy stream; the slot's previous forward has finished (wait_slot above)
  CU_TRY(cudaMemcpyAsync(sl.images.p, images_host, img_bytes, cudaMemcpyHostToDevice, e->copy_in));
  CU_TRY(cudaEventRecord(sl.in_done, e->copy_in));
  cudaStream_t st = e->stream;
  CU_TRY(cudaStreamWaitEvent(st, sl.in_done, 0));
  e->use_bound = true;
  const int frc = forward_device_locked(e, (const float*)sl.images.p, B, flags, st);
  e->use_bound = false;
  VT_TRY(frc);
  // staging copies of the outputs (device to device, tens of MB): the next forward may overwrite the engine's buffers
  if (!b_logits) CU_TRY(cudaMemcpyAsync(sl.logits.p, e->logits.p, logit_bytes, cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < L; ++l) {
    if (out->cls_maps)
      CU_TRY(cudaMemcpyAsync((char*)sl.cls.p + l * cls_bytes, (const float*)e->cls.p + (size_t)l * e->cap_batch * Hh * N, cls_bytes,
                             cudaMemcpyDeviceToDevice, st));
    if (out->avg_maps)
      CU_TRY(cudaMemcpy2DAsync((float*)sl.avg.p + (size_t)l * B * N * N, (size_t)N * 4,
                               (const float*)e->avg.p + (size_t)l * e->cap_batch * N * e->pitch, (size_t)e->pitch * 4, (size_t)N * 4,
                               (size_t)B * N, cudaMemcpyDeviceToDevice, st));
  }
  if (out->rollout) CU_TRY(cudaMemcpyAsync(sl.rollout.p, e->rollout.p, (size_t)B * (N - 1) * 4, cudaMemcpyDeviceToDevice, st));
  CU_TRY(cudaEventRecord(sl.compute_done, st));
  // D2H on the second copy stream
  CU_TRY(cudaStreamWaitEvent(e->copy_out, sl.compute_done, 0));
  if (out->logits) CU_TRY(cudaMemcpyAsync(out->logits, sl.logits.p, logit_bytes, cudaMemcpyDeviceToHost, e->copy_out));
  if (out->cls_maps) CU_TRY(cudaMemcpyAsync(out->cls_maps, sl.cls.p, cls_bytes * L, cudaMemcpyDeviceToHost, e->copy_out));
  if (out->rollout) CU_TRY(cudaMemcpyAsync(out->rollout, sl.rollout.p, (size_t)B * (N - 1) * 4, cudaMemcpyDeviceToHost, e->copy_out));
  if (out->avg_maps) CU_TRY(cudaMemcpyAsync(out->avg_maps, sl.avg.p, (size_t)L * B * N * N * 4, cudaMemcpyDeviceToHost, e->copy_out));
  CU_TRY(cudaEventRecord(sl.out_done, e->copy_out));
  sl.busy = true, sl.batch = B, sl.flags = flags;
  *ticket = e->submitted++;
  return VITB200_OK;
}

int vitb200_wait(vitb200_engine* e, uint64_t ticket) {
  if (!e) return fail(VITB200_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  CU_TRY(cudaSetDevice(e->cfg.device));
  if (ticket >= e->submitted) return fail(VITB200_ERR_STATE, "wait: ticket %llu was never issued", (unsigned long long)ticket);
  // a ticket whose slot has been taken again was drained when the newer request claimed the slot
  if (ticket + 2 < e->submitted) return VITB200_OK;
  return wait_slot_locked(e, e->slots[ticket & 1]);
}

int vitb200_staged_output(vitb200_engine* e, uint64_t ticket, uint32_t which, float** ptr_dev) {
  if (!e || !ptr_dev) return fail(VITB200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  if (ticket >= e->submitted || ticket + 2 < e->submitted) return fail(VITB200_ERR_STATE, "staged_output: ticket not in flight");
  vitb200_engine::Slot& sl = e->slots[ticket & 1];
  void* p = nullptr;
  switch (which) {
    case 0: p = sl.logits.p; break;
    case VITB200_EMIT_AVG: p = sl.avg.p; break;
    case VITB200_EMIT_CLS: p = sl.cls.p; break;
    case VITB200_EMIT_ROLLOUT: p = sl.rollout.p; break;
    default: return fail(VITB200_ERR_INVALID, "unknown staged output selector %u", which);
  }
  if (!p) return fail(VITB200_ERR_STATE, "output %u was not requested for this ticket", which);
  *ptr_dev = (float*)p;
  return VITB200_OK;
}

int vitb200_device_output(vitb200_engine* e, uint32_t which, float** ptr_dev, int* pitch) {
  if (!e || !ptr_dev) return fail(VITB200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  int pt = 0;
  void* p = nullptr;
  switch (which) {
    case 0: p = e->logits.p, pt = e->cfg.num_classes; break;
    case VITB200_EMIT_AVG: p = e->avg.p, pt = e->pitch; break;
    case VITB200_EMIT_CLS: p = e->cls.p, pt = e->N; break;
    case VITB200_EMIT_ROLLOUT: p = e->rollout.p, pt = e->N - 1; break;
    case VITB200_EMIT_HEADS: p = e->heads.p, pt = e->pitch; break;
    case VITB200_EMIT_HIDDEN: p = e->hidden.p, pt = e->cfg.hidden_dim; break;
    default: return fail(VITB200_ERR_INVALID, "unknown output selector %u", which);
  }
  if (!p) return fail(VITB200_ERR_STATE, "output %u has not been produced", which);
  *ptr_dev = (float*)p;
  if (pitch) *pitch = pt;
  return VITB200_OK;
}

int vitb200_synchronize(vitb200_engine* e) {
  if (!e) return fail(VITB200_ERR_INVALID, "null engine");
  CU_TRY(cudaSetDevice(e->cfg.device));
  CU_TRY(cudaStreamSynchronize(e->stream));
  if (e->side) CU_TRY(cudaStreamSynchronize(e->side));
  return VITB200_OK;
}

uint64_t vitb200_launch_count(vitb200_engine* e) { return e ? e->launches : 0; }

// ---- deferred node outputs on a side stream --------------------------------------------------------
static constexpr size_t kSideMaxBytes = 32u << 20;   // larger outputs (batched requests) keep the in-stream copy

static int side_ready(vitb200_engine* e) {
  if (e->side) return VITB200_OK;
  CU_TRY(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&e->side_ev, cudaEventDisableTiming));
  e->map_ev.assign(e->cfg.num_layers, nullptr);
  e->map_busy.assign(e->cfg.num_layers, 0);
  for (auto& ev : e->map_ev) CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (int k = 0; k < vitb200_engine::kTokRing; ++k) CU_TRY(cudaEventCreateWithFlags(&e->tok_ev[k], cudaEventDisableTiming));
  return VITB200_OK;
}
// The side stream may start once everything enqueued on the main stream so far has run (one record + wait per node).
static int side_follow(vitb200_engine* e, cudaStream_t st) {
  if (!e->side_stale) return VITB200_OK;
  CU_TRY(cudaEventRecord(e->side_ev, st));
  CU_TRY(cudaStreamWaitEvent(e->side, e->side_ev, 0));
  e->side_stale = false;
  return VITB200_OK;
}
// A launch that rewrites layer l's maps waits for their previous copies (normally the previous request's: long gone).
static int side_guard_maps(vitb200_engine* e, int layer, cudaStream_t st) {
  if (e->side && e->map_busy[layer]) {
    CU_TRY(cudaStreamWaitEvent(st, e->map_ev[layer], 0));
    e->map_busy[layer] = 0;
  }
  return VITB200_OK;
}
static int side_copy_map(vitb200_engine* e, int layer, float* dst, const float* src, size_t rows, int width, int pitch,
                         cudaStream_t st) {
  VT_TRY(side_ready(e));
  VT_TRY(side_follow(e, st));
  VT_TRY(copy_rows_to_host(dst, src, rows, width, pitch, e->side));
  CU_TRY(cudaEventRecord(e->map_ev[layer], e->side));
  e->map_busy[layer] = 1;
  return VITB200_OK;
}

// ---- node-granular stages --------------------------------------------------------------------------
#define STAGE_PROLOGUE(B, FLAGS)                                                            \
  if (!e) return fail(VITB200_ERR_INVALID, "null engine");                                  \
  std::lock_guard<std::mutex> lock(e->mu);                                                  \
  CU_TRY(cudaSetDevice(e->cfg.device));                                                     \
  VT_TRY(check_ready(e));                                                                   \
  VT_TRY(check_batch(e, (B)));                                                              \
  VT_TRY(ensure_workspace(e, (B) > e->cap_batch ? (B) : e->cap_batch, (FLAGS) | e->cap_flags)); \
  cudaStream_t st = e->stream;


int vitb200_set_deferred(vitb200_engine* e, int on) {
  if (!e) return fail(VITB200_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  e->deferred = on != 0;
  return VITB200_OK;
}

int vitb200_stage_embed(vitb200_engine* e, const float* images_host, int batch) {
  if (!images_host) return fail(VITB200_ERR_INVALID, "null images");
  STAGE_PROLOGUE(batch, 0)
  const size_t img_bytes = (size_t)batch * 3 * e->cfg.image_size * e->cfg.image_size * 4;
  CU_TRY(cudaMemcpyAsync(e->images.p, images_host, img_bytes, cudaMemcpyHostToDevice, st));
  VT_TRY(run_embed(e, (const float*)e->images.p, batch, st));
  CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_transform(vitb200_engine* e, const float* images_host, int batch, int H, int W, int resize,
                            float* out_host) {
  if (!images_host) return fail(VITB200_ERR_INVALID, "null images");
  STAGE_PROLOGUE(batch, 0)
  const int S = e->cfg.image_size;
  PreprocessParams p;
  VT_TRY(preprocess_params(batch, H, W, resize, S, &p));
  const size_t in_bytes = (size_t)batch * 3 * H * W * 4, out_bytes = (size_t)batch * 3 * S * S * 4;
  VT_TRY(ensure(e->stage_f32, in_bytes));
  CU_TRY(cudaMemcpyAsync(e->stage_f32.p, images_host, in_bytes, cudaMemcpyHostToDevice, st));
  preprocess_kernel<<<(unsigned)((out_bytes / 4 + 255) / 256), 256, 0, st>>>((const float*)e->stage_f32.p, (float*)e->images.p, p);
  CU_TRY(cudaGetLastError());
  e->launches += 1;
  if (out_host) CU_TRY(cudaMemcpyAsync(out_host, e->images.p, out_bytes, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_embed_resident(vitb200_engine* e, int batch) {
  STAGE_PROLOGUE(batch, 0)
  VT_TRY(run_graphed(e, graph_key(e, kGraphEmbed, 0, batch, 0, e->images.p), st,
                     [&] { return run_embed(e, (const float*)e->images.p, batch, st); }));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_layer(vitb200_engine* e, int layer, int batch, uint32_t flags) {
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, flags)
  VT_TRY(side_guard_maps(e, layer, st));
  VT_TRY(run_graphed(e, graph_key(e, kGraphLayer, layer, batch, flags, nullptr), st,
                     [&] { return run_layer(e, layer, batch, flags, st); }));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_layer_fetch(vitb200_engine* e, int layer, int batch, uint32_t flags, int half, float* tokens_host,
                              float* avg_host, float* cls_grid_host) {
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  if ((avg_host && !(flags & VITB200_EMIT_AVG)) || (cls_grid_host && !(flags & VITB200_EMIT_CLS)))
    return fail(VITB200_ERR_INVALID, "stage_layer_fetch: a map is requested that the flags do not emit");
  STAGE_PROLOGUE(batch, flags)
  VT_TRY(side_guard_maps(e, layer, st));
  if (half == 0)
    VT_TRY(run_graphed(e, graph_key(e, kGraphLayer, layer, batch, flags, nullptr), st,
                       [&] { return run_layer(e, layer, batch, flags, st); }));
  else
    VT_TRY(run_graphed(e, graph_key(e, kGraphAttnBlock, layer, batch, flags, nullptr), st,
                       [&] { return run_attn_block(e, layer, batch, flags, st); }));
  const int N = e->N, H = e->cfg.num_heads;
  const size_t tok_bytes = (size_t)batch * N * e->cfg.hidden_dim * 4;
  const float* avg_src = (const float*)e->avg.p + (size_t)layer * e->cap_batch * N * e->pitch;
  const float* cls_src = (const float*)e->cls.p + (size_t)layer * e->cap_batch * H * N + 1;
  if (e->deferred && tok_bytes <= kSideMaxBytes && (size_t)batch * N * N * 4 <= kSideMaxBytes) {
    // every copy of the node on the side stream behind ONE event (see the getters for the individual steps)
    VT_TRY(side_ready(e));
    int k = -1;
    if (tokens_host) {
      k = e->tok_next;
      e->tok_next = (k + 1) % vitb200_engine::kTokRing;
      if (e->tok_busy[k]) {
        CU_TRY(cudaStreamWaitEvent(st, e->tok_ev[k], 0));
        e->tok_busy[k] = false;
      }
      VT_TRY(ensure(e->tok_ring[k], tok_bytes));
      CU_TRY(cudaMemcpyAsync(e->tok_ring[k].p, e->x.p, tok_bytes, cudaMemcpyDeviceToDevice, st));
    }
    e->side_stale = true;
    VT_TRY(side_follow(e, st));
    if (tokens_host) {
      CU_TRY(cudaMemcpyAsync(tokens_host, e->tok_ring[k].p, tok_bytes, cudaMemcpyDeviceToHost, e->side));
      CU_TRY(cudaEventRecord(e->tok_ev[k], e->side));
      e->tok_busy[k] = true;
    }
    if (avg_host) VT_TRY(copy_rows_to_host(avg_host, avg_src, (size_t)batch * N, N, e->pitch, e->side));
    if (cls_grid_host) VT_TRY(copy_rows_to_host(cls_grid_host, cls_src, (size_t)batch * H, N - 1, N, e->side));
    if (avg_host || cls_grid_host) {
      CU_TRY(cudaEventRecord(e->map_ev[layer], e->side));
      e->map_busy[layer] = 1;
    }
    return VITB200_OK;
  }
  if (tokens_host) CU_TRY(cudaMemcpyAsync(tokens_host, e->x.p, tok_bytes, cudaMemcpyDeviceToHost, st));
  if (avg_host) VT_TRY(copy_rows_to_host(avg_host, avg_src, (size_t)batch * N, N, e->pitch, st));
  if (cls_grid_host) VT_TRY(copy_rows_to_host(cls_grid_host, cls_src, (size_t)batch * H, N - 1, N, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_attn_block(vitb200_engine* e, int layer, int batch, uint32_t flags) {
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, flags)
  VT_TRY(side_guard_maps(e, layer, st));
  VT_TRY(run_graphed(e, graph_key(e, kGraphAttnBlock, layer, batch, flags, nullptr), st,
                     [&] { return run_attn_block(e, layer, batch, flags, st); }));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_mlp_block(vitb200_engine* e, int layer, int batch) {
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, 0)
  VT_TRY(run_graphed(e, graph_key(e, kGraphMlpBlock, layer, batch, 0, nullptr), st,
                     [&] { return run_mlp_block(e, layer, batch, st); }));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_head(vitb200_engine* e, int batch, float* logits_host) {
  STAGE_PROLOGUE(batch, 0)
  VT_TRY(run_graphed(e, graph_key(e, kGraphHead, 0, batch, 0, nullptr), st, [&] { return run_head(e, batch, st); }));
  if (logits_host)
    CU_TRY(cudaMemcpyAsync(logits_host, e->logits.p, (size_t)batch * e->cfg.num_classes * 4, cudaMemcpyDeviceToHost, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_stage_rollout(vitb200_engine* e, int batch, float* rollout_host) {
  STAGE_PROLOGUE(batch, VITB200_EMIT_ROLLOUT)
  VT_TRY(run_graphed(e, graph_key(e, kGraphRollout, 0, batch, 0, nullptr), st, [&] { return run_rollout(e, batch, st); }));
  if (rollout_host)
    CU_TRY(cudaMemcpyAsync(rollout_host, e->rollout.p, (size_t)batch * (e->N - 1) * 4, cudaMemcpyDeviceToHost, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_set_tokens(vitb200_engine* e, const float* tokens_host, int batch) {
  if (!tokens_host) return fail(VITB200_ERR_INVALID, "null tokens");
  STAGE_PROLOGUE(batch, 0)
  CU_TRY(cudaMemcpyAsync(e->x.p, tokens_host, (size_t)batch * e->N * e->cfg.hidden_dim * 4, cudaMemcpyHostToDevice, st));
  // what the folded LayerNorm of the next GEMM reads: bf16 copy + per-chunk partial sums of every row
  const long rows = (long)batch * e->N;
  rows_bf16_stats_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>((const float*)e->x.p, (__nv_bfloat16*)e->xb.p,
                                                                              (float2*)e->ln_stats.p, rows, e->cfg.hidden_dim,
                                                                              stats_width((int)rows, e->cfg.hidden_dim),
                                                                              (__nv_bfloat16*)e->xb_lo.p);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_get_tokens(vitb200_engine* e, float* tokens_host, int batch) {
  if (!tokens_host) return fail(VITB200_ERR_INVALID, "null tokens");
  STAGE_PROLOGUE(batch, 0)
  const size_t bytes = (size_t)batch * e->N * e->cfg.hidden_dim * 4;
  if (e->deferred && bytes <= kSideMaxBytes) {
    // the next node rewrites x in place: snapshot it (a few us on the main stream) and copy the snapshot out on the side
    VT_TRY(side_ready(e));
    const int k = e->tok_next;
    e->tok_next = (k + 1) % vitb200_engine::kTokRing;
    if (e->tok_busy[k]) {   // the slot's previous copy (four token outputs ago) has left
      CU_TRY(cudaStreamWaitEvent(st, e->tok_ev[k], 0));
      e->tok_busy[k] = false;
    }
    VT_TRY(ensure(e->tok_ring[k], bytes));
    CU_TRY(cudaMemcpyAsync(e->tok_ring[k].p, e->x.p, bytes, cudaMemcpyDeviceToDevice, st));
    e->side_stale = true;
    VT_TRY(side_follow(e, st));
    CU_TRY(cudaMemcpyAsync(tokens_host, e->tok_ring[k].p, bytes, cudaMemcpyDeviceToHost, e->side));
    CU_TRY(cudaEventRecord(e->tok_ev[k], e->side));
    e->tok_busy[k] = true;
    return VITB200_OK;
  }
  CU_TRY(cudaMemcpyAsync(tokens_host, e->x.p, bytes, cudaMemcpyDeviceToHost, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_set_avg_map(vitb200_engine* e, int layer, const float* map_host, int batch) {
  if (!map_host) return fail(VITB200_ERR_INVALID, "null map");
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, VITB200_EMIT_AVG)
  VT_TRY(side_guard_maps(e, layer, st));   // a side-stream copy of this layer's map may still be reading it
  e->side_stale = true;
  float* dst = (float*)e->avg.p + (size_t)layer * e->cap_batch * e->N * e->pitch;
  CU_TRY(cudaMemcpy2DAsync(dst, (size_t)e->pitch * 4, map_host, (size_t)e->N * 4, (size_t)e->N * 4, (size_t)batch * e->N,
                           cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_get_avg_map(vitb200_engine* e, int layer, float* map_host, int batch) {
  if (!map_host) return fail(VITB200_ERR_INVALID, "null map");
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, VITB200_EMIT_AVG)
  const float* src = (const float*)e->avg.p + (size_t)layer * e->cap_batch * e->N * e->pitch;
  if (e->deferred && (size_t)batch * e->N * e->N * 4 <= kSideMaxBytes)
    return side_copy_map(e, layer, map_host, src, (size_t)batch * e->N, e->N, e->pitch, st);
  VT_TRY(copy_rows_to_host(map_host, src, (size_t)batch * e->N, e->N, e->pitch, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_get_cls_map(vitb200_engine* e, int layer, float* map_host, int batch) {
  if (!map_host) return fail(VITB200_ERR_INVALID, "null map");
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, VITB200_EMIT_CLS)
  const float* src = (const float*)e->cls.p + (size_t)layer * e->cap_batch * e->cfg.num_heads * e->N;
  if (e->deferred && (size_t)batch * e->cfg.num_heads * e->N * 4 <= kSideMaxBytes)
    return side_copy_map(e, layer, map_host, src, (size_t)batch * e->cfg.num_heads, e->N, e->N, st);
  CU_TRY(cudaMemcpyAsync(map_host, src, (size_t)batch * e->cfg.num_heads * e->N * 4, cudaMemcpyDeviceToHost, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_get_cls_grid(vitb200_engine* e, int layer, float* map_host, int batch) {
  if (!map_host) return fail(VITB200_ERR_INVALID, "null map");
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, VITB200_EMIT_CLS)
  // the class token's attention to the PATCH tokens only: column 0 (class -> class) is dropped by the strided copy
  const float* src = (const float*)e->cls.p + (size_t)layer * e->cap_batch * e->cfg.num_heads * e->N + 1;
  if (e->deferred && (size_t)batch * e->cfg.num_heads * e->N * 4 <= kSideMaxBytes)
    return side_copy_map(e, layer, map_host, src, (size_t)batch * e->cfg.num_heads, e->N - 1, e->N, st);
  VT_TRY(copy_rows_to_host(map_host, src, (size_t)batch * e->cfg.num_heads, e->N - 1, e->N, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

int vitb200_get_head_map(vitb200_engine* e, int layer, float* map_host, int batch) {
  if (!map_host) return fail(VITB200_ERR_INVALID, "null map");
  if (e && (layer < 0 || layer >= e->cfg.num_layers)) return fail(VITB200_ERR_INVALID, "layer %d out of range", layer);
  STAGE_PROLOGUE(batch, VITB200_EMIT_HEADS)
  const float* src = (const float*)e->heads.p + (size_t)layer * e->cap_batch * e->cfg.num_heads * e->N * e->pitch;
  if (e->deferred && (size_t)batch * e->cfg.num_heads * e->N * e->N * 4 <= kSideMaxBytes)
    return side_copy_map(e, layer, map_host, src, (size_t)batch * e->cfg.num_heads * e->N, e->N, e->pitch, st);
  VT_TRY(copy_rows_to_host(map_host, src, (size_t)batch * e->cfg.num_heads * e->N, e->N, e->pitch, st));
  if (!e->deferred) CU_TRY(cudaStreamSynchronize(st));
  return VITB200_OK;
}

#ifdef VITB200_ATTN_TRACE
// tracing build only (tools/attn_trace.py): copies the clock64() stamps of CTA 0 to the host
int vitb200_debug_attn_trace(long long* host, int count) {
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(long long) * count));
  return VITB200_OK;
}
#endif

// ---- single-kernel entry points (parity tests) -----------------------------------------------------
int vitb200_op_gemm(const void* a, const void* w, const float* bias, const float* resid, void* out, int M, int N, int K,
                    int gelu, int out_f32, void* stream) {
  if (!a || !w || !out) return fail(VITB200_ERR_INVALID, "null argument");
  GemmEpilogue ep;
  ep.bias = bias, ep.out = out, ep.ldo = N, ep.resid = resid, ep.ldr = N;
  return launch_gemm(a, K, w, M, N, K, ep, gelu != 0, out_f32 != 0, (cudaStream_t)stream);
}

int vitb200_op_gemm_ex(const void* a, const void* w, const float* bias, const float* resid, void* out, int M, int N, int K,
                       int gelu, int out_f32, void* xb_out, float* stats_out, const float* stats_in, const float* colsum,
                       float ln_eps, void* stream) {
  if (!a || !w || !out) return fail(VITB200_ERR_INVALID, "null argument");
  GemmEpilogue ep;
  ep.bias = bias, ep.out = out, ep.ldo = N, ep.resid = resid, ep.ldr = N;
  if (xb_out) {
    if (N % 128 != 0) return fail(VITB200_ERR_INVALID, "gemm: row statistics need N to be a multiple of 128");
    ep.xb = (__nv_bfloat16*)xb_out, ep.ldxb = N, ep.row_stats_out = (float2*)stats_out, ep.stats_slots = N / ln_slot_width(N);
  }
  if (stats_in) {
    if (K % 128 != 0) return fail(VITB200_ERR_INVALID, "gemm: folded LayerNorm needs K to be a multiple of 128");
    ep.row_stats_in = (const float2*)stats_in, ep.stats_in_slots = K / ln_slot_width(K), ep.ln_eps = ln_eps, ep.colsum = colsum;
  }
  return launch_gemm(a, K, w, M, N, K, ep, gelu != 0, out_f32 != 0, (cudaStream_t)stream);
}

int vitb200_op_split_bf16(const float* in_dev, void* hi_dev, void* lo_dev, size_t count, void* stream) {
  if (!in_dev || !hi_dev || !lo_dev) return fail(VITB200_ERR_INVALID, "null argument");
  f32_to_bf16_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in_dev, (__nv_bfloat16*)hi_dev,
                                                                                            (__nv_bfloat16*)lo_dev, (long)count);
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

int vitb200_op_gemm_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                          const float* resid, void* out, void* out_lo, int M, int N, int K, int gelu, int out_f32, void* stream) {
  if (!a_hi || !a_lo || !w_hi || !w_lo || !out) return fail(VITB200_ERR_INVALID, "null argument");
  GemmEpilogue ep;
  ep.bias = bias, ep.out = out, ep.ldo = N, ep.resid = resid, ep.ldr = N;
  ep.out_lo = (__nv_bfloat16*)out_lo;
  return launch_gemm(a_hi, K, w_hi, M, N, K, ep, gelu != 0, out_f32 != 0, (cudaStream_t)stream, a_lo, w_lo);
}

int vitb200_op_fold_ln(const float* w, const float* gamma, const float* beta, const float* bias, void* wq, float* colsum,
                       float* bias_out, int N, int K, void* stream) {
  if (!w || !gamma || !beta || !bias || !wq || !colsum || !bias_out) return fail(VITB200_ERR_INVALID, "null argument");
  fold_ln_weight_kernel<<<(N * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, gamma, beta, bias, (__nv_bfloat16*)wq, colsum,
                                                                              bias_out, N, K);
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

int vitb200_op_layernorm(const float* x, const float* gamma, const float* beta, void* y, int rows, int d, float eps,
                         void* stream) {
  if (!x || !gamma || !beta || !y) return fail(VITB200_ERR_INVALID, "null argument");
  return launch_layernorm(x, d, gamma, beta, (__nv_bfloat16*)y, rows, d, eps, (cudaStream_t)stream);
}

// statistics scratch of the long-sequence attention path for the single-kernel entry point (grown on demand)
static float2* op_attention_stats(size_t count) {
  std::lock_guard<std::mutex> lock(g_dev_mu);
  DeviceCtx& dc = dev_ctx_locked();
  if (count > dc.op_stats_cap) {
    if (dc.op_stats) cudaFree(dc.op_stats);
    dc.op_stats = nullptr, dc.op_stats_cap = 0;
    if (cudaMalloc(&dc.op_stats, count * sizeof(float2)) == cudaSuccess) dc.op_stats_cap = count;
  }
  return dc.op_stats;
}

int vitb200_op_attention_ex(const void* qkv, void* ctx, float* avg, float* cls, float* heads, int batch, int tokens,
                            int nheads, int head_dim, int pitch, int v_is_f16, void* stream) {
  if (!qkv || !ctx) return fail(VITB200_ERR_INVALID, "null argument");
  float2* stats = nullptr;
  const bool fused = attention_is_fused(tokens, head_dim);
  if (!fused) {
    if (v_is_f16) return fail(VITB200_ERR_INVALID, "attention: the key-blocked path takes V in bf16");
    stats = op_attention_stats((size_t)batch * nheads * tokens);
    if (!stats) return fail(VITB200_ERR_CUDA, "attention: cannot allocate the statistics scratch");
  } else if (!v_is_f16) {
    // the fused kernels want the V third in fp16 (the forward's qkv GEMM writes it that way): convert a scratch copy
    const long rows = (long)batch * tokens;
    const int d = nheads * head_dim;
    const size_t bytes = (size_t)rows * 3 * d * 2;
    uint16_t* tmp = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_dev_mu);
      DeviceCtx& dc = dev_ctx_locked();
      if (bytes > dc.op_qkv_cap) {
        if (dc.op_qkv) cudaFree(dc.op_qkv);
        dc.op_qkv = nullptr, dc.op_qkv_cap = 0;
        CU_TRY(cudaMalloc(&dc.op_qkv, bytes));
        dc.op_qkv_cap = bytes;
      }
      tmp = (uint16_t*)dc.op_qkv;
    }
    const long vecs = rows * 3 * d / 8;
    qkv_v_to_f16_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, tmp, rows, d);
    CU_TRY(cudaGetLastError());
    qkv = tmp;
  }
  float* parts = nullptr;
  if (fused && avg) {   // scratch of the head-split small launches (grown on demand; without it they split two ways at most)
    const size_t bytes = attention_parts_bytes(tokens, pitch);
    std::lock_guard<std::mutex> lock(g_dev_mu);
    DeviceCtx& dc = dev_ctx_locked();
    if (bytes > dc.op_parts_cap) {
      if (dc.op_parts) cudaFree(dc.op_parts);
      dc.op_parts = nullptr, dc.op_parts_cap = 0;
      if (cudaMalloc(&dc.op_parts, bytes) == cudaSuccess) dc.op_parts_cap = bytes;
    }
    parts = (float*)dc.op_parts;
  }
  return launch_attention((const __nv_bfloat16*)qkv, (__nv_bfloat16*)ctx, avg, cls, heads, batch, tokens, nheads, head_dim,
                          pitch, stats, (cudaStream_t)stream, nullptr, nullptr, parts);
}

int vitb200_op_attention(const void* qkv, void* ctx, float* avg, float* cls, float* heads, int batch, int tokens,
                         int nheads, int pitch, void* stream) {
  return vitb200_op_attention_ex(qkv, ctx, avg, cls, heads, batch, tokens, nheads, 64, pitch, 0, stream);
}

// torchvision F.resize with a single int (shorter side -> resize, longer = int(resize * long / short)) followed by
// F.center_crop (offsets = round-half-even((size - crop) / 2)), ImageNet mean / std.
static int preprocess_params(int B, int H, int W, int resize, int crop, PreprocessParams* p) {
  if (B <= 0 || H <= 0 || W <= 0 || resize <= 0 || crop <= 0) return fail(VITB200_ERR_INVALID, "preprocess: bad geometry");
  p->B = B, p->H = H, p->W = W;
  if (H <= W) p->RH = resize, p->RW = (int)((long)resize * W / H);
  else p->RW = resize, p->RH = (int)((long)resize * H / W);
  if (p->RH < crop || p->RW < crop) return fail(VITB200_ERR_INVALID, "preprocess: crop %d larger than the resized image %dx%d", crop, p->RH, p->RW);
  p->crop = crop;
  p->top = (int)nearbyint((p->RH - crop) / 2.0), p->left = (int)nearbyint((p->RW - crop) / 2.0);  // ties to even, like Python's round
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c) p->mean[c] = mean[c], p->inv_std[c] = 1.0f / stdv[c];
  return VITB200_OK;
}

int vitb200_op_preprocess(const float* images_dev, float* out_dev, int batch, int H, int W, int resize, int crop, void* stream) {
  if (!images_dev || !out_dev) return fail(VITB200_ERR_INVALID, "null argument");
  PreprocessParams p;
  VT_TRY(preprocess_params(batch, H, W, resize, crop, &p));
  const long total = (long)batch * 3 * crop * crop;
  preprocess_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(images_dev, out_dev, p);
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

int vitb200_op_patchify(const float* images, void* patches, int batch, int image_size, int patch, void* stream) {
  if (!images || !patches) return fail(VITB200_ERR_INVALID, "null argument");
  if (patch <= 0 || patch % 8 != 0 || image_size % patch != 0) return fail(VITB200_ERR_INVALID, "bad patch geometry");
  const long items = (long)batch * 3 * image_size * (image_size / patch) * (patch / 8);
  patchify_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(images, (__nv_bfloat16*)patches,
                                                                                     batch, image_size, patch);
  CU_TRY(cudaGetLastError());
  return VITB200_OK;
}

int vitb200_op_rollout(const float* maps, long layer_stride, int layers, int batch, int tokens, int pitch, float* out,
                       void* stream) {
  if (!maps || !out) return fail(VITB200_ERR_INVALID, "null argument");
  return launch_rollout(maps, layer_stride, layers, batch, tokens, pitch, out, (cudaStream_t)stream);
}

}  // extern "C"
