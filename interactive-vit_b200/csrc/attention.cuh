// Fused multi-head self-attention with attention-map emission (tcgen05 + TMEM), for sequences whose padded
// key count KP = round_up(N, 16) fits one UMMA N and this kernel's TMEM plan (KP <= 208): N = 197
// (224 px / patch 16) -> KP = 208.
//
// Arithmetic follows torch.nn.functional.multi_head_attention_forward's weights branch
// (torch/nn/functional.py:6630-6659: q * 1/sqrt(D), bmm(q, k^T), softmax, bmm(P, v), optional head mean)
// which is what the oracle obtains from torchvision's EncoderBlock (vision_transformer.py:110-119) with
// need_weights=True.  The 1/sqrt(D) scale is applied to the fp32 scores instead of to q (identical in
// exact arithmetic; exact in floating point too when D = 64 since the factor is a power of two).
//
// One CTA per (image b, 128-row query tile qt); the CTA loops over all H heads so that the head-averaged
// probabilities can be accumulated on chip (TMEM) and written to HBM exactly once.  640 threads; roles (the four
// control roles sit on the highest hardware warp ids):
//   role 0 (one elected lane) : TMA producer  Q_h [128 x D], K_h [KP x D], V_h [KP x D] tiles of the packed qkv
//                   activation, 2-stage ring over heads
//   role 1 (one elected lane) : UMMA issuer   S = Q K^T (M=128, N=KP, K=D)   -> TMEM cols [64, 64+KP)
//                                 O = P V   (M=128, N=D,  K=KP)  -> TMEM cols [0, D), V is the MN-major B
//                                 Pbar += P I16 (M=128, N=16 per 16 keys)      -> TMEM cols [288, 288+KP)
//                   Issue order QK(0), QK(1), PV(0)+AVG(0), QK(2), ...: the S columns are released as soon as the
//                   softmax warps have copied them to registers, so QK^T of head h+1 and P V of head h run on the
//                   tensor pipe underneath the softmax of head h.
//   role 2        : TMEM allocator (all 512 columns)
//   role 3        : writes the 16 x 16 identity operand, then streams the class-token rows (row 0 of P, fp32) that
//                   the softmax warps stage in smem to HBM
//   roles 4..19   : softmax, FOUR threads per query row (one warp per TMEM lane quarter and column group; a warp may
//                   only touch the lane quarter warp % 4).  Each thread reads its <= 56 scores from TMEM ONCE and keeps
//                   them in registers: thread-local max -> e = exp2((s - m_t) c) on the MUFU -> ONE exchange of
//                   (m_t, sum_t) through smem and a named barrier -> p = e f_t normalised -> bf16 P tile in shared
//                   memory (128-B swizzled K-major A operand of P V and of the head-average MMAs).  The context tile
//                   of the previous head goes TMEM -> bf16 -> smem -> TMA store in the same pass; one
//                   fence.proxy.async per head covers both tiles.  After the last head the head-average tile leaves
//                   through smem slabs and TMA stores (reduce-add for CTAs that share an item, see AttnParams).
//
// History of the design (measured, profiles/README.md): two threads per row with the head average added in TMEM by the
// softmax threads took 7,285 cycles per head; four threads per row, the head average on the tensor pipe, TMA-stored
// tiles, thread-local maxima and the class rows on a spare warp brought it to ~5,400.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

namespace vitb200 {

__device__ __forceinline__ uint32_t pack_bf16x2_u(uint32_t a_f32_bits, uint32_t b_f32_bits) {
  __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(a_f32_bits), __uint_as_float(b_f32_bits));
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp32 pair -> packed fp16x2 / packed fp16x2 multiply.  The probabilities are carried in FP16 (not bf16): they lie in
// [0, 1], where fp16 has 10 mantissa bits against bf16's 7, and fp16 makes the normalisation a packed multiply.
__device__ __forceinline__ uint32_t pack_f16x2_f(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t hmul2_u(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

struct AttnParams {
  int B, N, H;          // images, tokens per image, heads
  int d;                // model width = H * D
  int KP;               // keys padded to a multiple of 16 (<= 208)
  float scale_log2;     // (1/sqrt(D)) * log2(e)
  __nv_bfloat16* ctx;   // [B*N, d] attention context (input of out_proj)
  float* avg_map;       // [B, N, ldmap] head-averaged probabilities, or nullptr
  float* head_map;      // [B, H, N, ldmap] per-head probabilities, or nullptr (opt-in, large)
  float* cls_map;       // [B, H, N] per-head probabilities of query token 0, or nullptr
  int ldmap;            // row stride of avg_map/head_map in floats (>= KP, multiple of 4)
  int q_tiles;          // 128-row query tiles per image: ceil(N / 128) (1 or 2)
  // Tail balancing: work items (image, query tile) [0, full_items) get one CTA each; every later item is split over
  // TWO CTAs by heads ([0, H/2) and [H/2, H)) whose head-average tiles are combined with a TMA reduce-add into rows the
  // host zeroed.  512 equal items on 148 SMs otherwise leave 80 SMs idle for the whole last round.
  int full_items;
  // Small launches (every item split: full_items = 0, up to sms / items parts): `split` parts per item; with more than two
  // parts a reduce-add would depend on arrival order, so part s stores its share of the head average (already scaled by
  // 1 / H) into image s * part_images + b of a scratch tensor (tmap_avg then describes the scratch) and
  // avg_parts_sum_kernel adds the parts in index order.
  int split = 2;
  int part_images = 0;   // 0: two parts, reduce-add into avg_map
  int one = 1;          // always 1 (attention_pp.cuh: trip count of its scheduling-fence loops)
};

// Optional phase tracing (built only with -DVITB200_ATTN_TRACE into a separate library, tools/attn_trace.py):
// CTA 0's softmax warp 4 / lane 0 and the MMA warp write clock64() stamps per head to a global buffer.
#ifdef VITB200_ATTN_TRACE
__device__ long long g_attn_trace[64 * 32];
#define ATTN_TS(slot) do { if (blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 1)) g_attn_trace[(h) * 32 + (slot)] = clock64(); \
    else if (blockIdx.x == 0 && lane == 0 && warp == 9 && (slot) >= 4 && (slot) < 16) g_attn_trace[(h) * 32 + 16 + (slot)] = clock64(); } while (0)
#else
#define ATTN_TS(slot) do { } while (0)
#endif

namespace attn_cfg {
constexpr int kCtrlWarps = 4;                      // TMA, MMA, TMEM allocator, spare
constexpr int kColGroups = 4;                      // softmax threads per query row
constexpr int kSoftmaxWarps = 4 * kColGroups;      // one warp per (TMEM lane quarter, column group)
constexpr int kSoftmaxThreads = 32 * kSoftmaxWarps;
constexpr int kThreads = 32 * kCtrlWarps + kSoftmaxThreads;  // 640
constexpr int BM = 128;
constexpr int D = 64;
constexpr int KP_MAX = 208;  // TMEM plan: O [0,64) | S [64,64+KP) | Pbar [288,288+KP)  -> KP <= 208
constexpr int kMaxGran = 7;                        // 8-key granules per softmax thread: ceil(26 / 4)
constexpr int kMaxKSteps = KP_MAX / 16;            // 13
constexpr int kTmemO = 0;
constexpr int kTmemS = 64;
constexpr int kTmemAvg = 288;
constexpr int kQBytes = BM * D * 2;                // 16 KB
constexpr int kKVBytes = KP_MAX * D * 2;           // 26 KB (two TMA boxes of KP/2 rows)
constexpr int kStageBytes = kQBytes + 2 * kKVBytes;
constexpr int kPBlockBytes = BM * 128;             // one 64-key K-block of P: 128 rows x 128 B
constexpr int kPBlocks = (KP_MAX + 63) / 64;       // 4
constexpr int kPBytes = kPBlocks * kPBlockBytes;   // 64 KB
constexpr int kIdentBytes = 16 * 128;              // 16 x 16 bf16 identity in 128-B rows (B operand of the head-average MMAs)
constexpr int kCtxStageBytes = BM * D * 2;         // 16 KB: bf16 context tile of one head, staged for the TMA store
constexpr int kRedBytes = 2 * 2 * kColGroups * BM * 4;  // row max / row sum exchange: [head parity][2 kinds][4 groups][128 rows]
constexpr int kClsStageBytes = KP_MAX * 4 + 16;    // exponentials of query row 0 (one head) + the 4 column groups' factors
constexpr int kSmemBytes = 2 * kStageBytes + kPBytes + kCtxStageBytes + kIdentBytes + kRedBytes + kClsStageBytes + 144;   // 16 mbarriers + the TMEM slot
static_assert(kSmemBytes <= 227 * 1024, "attention: shared memory budget");
}  // namespace attn_cfg

// kHeads: also write the full per-head probabilities (opt-in; a separate instantiation keeps that code out of
// the instruction stream of the common variant).
// kFull: KP == 208 (197 tokens, the production shape): every thread owns granules 0..5 and column groups 0 and 1 a
// seventh, so the per-granule guards fold at compile time instead of costing a branch pair per granule and phase.
template <bool kHeads, bool kFull = false>
__global__ void __launch_bounds__(attn_cfg::kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box 64 x 128 over qkv [B*N, 3d]
                 const __grid_constant__ CUtensorMap tmap_kv,  // box 64 x (KP/2) over the same tensor
                 const __grid_constant__ CUtensorMap tmap_ctx, // box 64 x 32 x 1 over ctx viewed as [B][N][d]: rows >= N clip
                 const __grid_constant__ CUtensorMap tmap_avg, // fp32, box 32 x 128 x 1 over avg_map viewed as [B][N][ldmap]
                 AttnParams p) {
  using namespace attn_cfg;
  // Dynamic smem starts 1024-B aligned (it follows the 1 KB the driver reserves); keeping the array typed lets
  // the compiler emit LDS/STS instead of generic loads for the exchange buffers.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_p = smem + 2 * kStageBytes;
  uint8_t* smem_ctx = smem_p + kPBytes;      // 4 quarter tiles of 32 rows x 128 B, 128-B swizzle
  uint8_t* smem_id = smem_ctx + kCtxStageBytes;
  float* red = reinterpret_cast<float*>(smem_id + kIdentBytes);  // [kind][group][row]
  float* cls_stage = red + 2 * 2 * kColGroups * BM;               // [KP_MAX] exp2 values of query row 0 (not yet normalised)
  float* cls_factor = cls_stage + KP_MAX;                         // [4] per column group: exp2(m_t - M) / sum
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_id + kIdentBytes + kRedBytes + kClsStageBytes);
  // Q / K and V of a head travel separately: the Q and K tiles of a stage are dead as soon as S = Q K^T has been
  // computed (early in the head), V only after P V -- with one barrier pair per stage (round 1) the loads of head h + 2
  // could only start behind P V of head h and S of head h + 2 arrived ~430 cycles late at every loop top (traced).
  uint64_t* qk_full = bars;         // [2] Q and K of a head landed
  uint64_t* v_full = bars + 2;      // [2] V of a head landed
  uint64_t* qk_empty = bars + 4;    // [2] Q / K consumed by QK^T
  uint64_t* v_empty = bars + 6;     // [2] V consumed by P V
  uint64_t* s_full = bars + 8;      // S = QK^T complete
  uint64_t* s_free = bars + 9;      // S copied to registers by all softmax warps
  uint64_t* p_full = bars + 10;     // fp16 P tile written to smem by all softmax warps
  uint64_t* o_full = bars + 11;     // O = PV complete
  uint64_t* o_free = bars + 12;     // O columns read by the four control warps (one per TMEM lane quarter)
  uint64_t* p_free = bars + 13;     // Pbar += P complete as well: the P tile may be overwritten
  uint64_t* cls_full = bars + 14;   // row 0 of the exponentials + factors staged in smem (4 column-group warps)
  uint64_t* cls_free = bars + 15;   // ... and copied out by warp 3
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  // role index: control roles 0..3 live on the highest hardware warp ids (scheduler priority), softmax roles 4..19 on
  // the lowest; role % 4 == hardware warp % 4 (TMEM lane quarters)
  const int warp = static_cast<int>((threadIdx.x >> 5) + kCtrlWarps) % (kThreads / 32);
  const int lane = threadIdx.x & 31;
  int item = blockIdx.x, h0 = 0, nh = p.H;
  bool split_cta = false;
  int avg_img = 0;   // p.part_images > 0: this part's slab of the head-average scratch starts at image avg_img
  if (item >= p.full_items) {
    const int r = item - p.full_items;
    const int it = r / p.split, part = r - it * p.split;
    item = p.full_items + it;
    h0 = part * p.H / p.split;                 // two parts: [0, H / 2) and [H / 2, H)
    nh = (part + 1) * p.H / p.split - h0;
    split_cta = true;
    avg_img = part * p.part_images;
  }
  const int b = item / p.q_tiles;
  const int qt = item - b * p.q_tiles;
  const int KP = p.KP;
  const int half_rows = KP >> 1;
  const uint32_t kv_tx = static_cast<uint32_t>(KP) * D * 2;
  const bool want_avg = p.avg_map != nullptr;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_kv);
    ptx::prefetch_tmap(&tmap_ctx);
    ptx::prefetch_tmap(&tmap_avg);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&qk_full[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&qk_empty[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
    }
    ptx::mbar_init(s_full, 1);
    // one arrival per softmax WARP (lane 0 after __syncwarp): 512 lanes arriving on one smem word serialise
    ptx::mbar_init(s_free, kSoftmaxWarps);
    ptx::mbar_init(p_full, kSoftmaxWarps);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_free, kCtrlWarps);
    ptx::mbar_init(p_free, 1);
    ptx::mbar_init(cls_full, kColGroups);
    ptx::mbar_init(cls_free, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  if (warp == 3) {
    // 16 x 16 fp16 identity (the operand format of P), K-major rows of 128 B in the 128-B swizzle: element (n, k) lives in 16-B chunk
    // (k / 8) ^ (n % 8) of row n.  Pbar[:, 16 ks + n] += sum_k P[:, 16 ks + k] * I[n, k] accumulates the normalised
    // probabilities over the heads on the tensor pipe (fp32 in TMEM) instead of a TMEM load/add/store pass per head
    // in the softmax threads.
    uint4* id = reinterpret_cast<uint4*>(smem_id);
    for (int i = lane; i < kIdentBytes / 16; i += 32) {
      const int n = i >> 3, c = (i & 7) ^ (n & 7);  // this physical chunk holds k in [8 c, 8 c + 8)
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((n >> 3) == c) {
        const uint32_t one = 0x3C00u << (16 * (n & 1));  // fp16 1.0 at position n % 8 of the chunk
        const int w = (n & 7) >> 1;
        v.x = w == 0 ? one : 0u, v.y = w == 1 ? one : 0u, v.z = w == 2 ? one : 0u, v.w = w == 3 ? one : 0u;
      }
      id[i] = v;
    }
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::grid_dep_launch();   // PDL (ptx.cuh): prologue above overlaps the previous kernel; nothing below runs before it has completed
  ptx::grid_dep_wait();

  const int row0 = b * p.N;  // first token row of this image in the [B*N, 3d] activation

  // Producer and MMA loops run on whole warps with one elected lane issuing, so that addresses and descriptors
  // stay in uniform registers (a single-lane loop pays an R2UR chain in front of every UTMALDG / UTCHMMA).
  // Context epilogue, one control warp per TMEM lane quarter (role q reads lanes 32 q .. 32 q + 31): O of head hh is final
  // (P was normalised before the MMA) -> bf16 -> this quarter's smem tile (32 rows x 128 B, 128-B swizzle) -> one TMA
  // store.  The tensor map views ctx as [B][N][d], so rows of the last query tile that lie beyond the image are clipped.
  // Round 1 did this inside the softmax warps (~670 cycles of their ~4,800-cycle head period, traced); the control
  // warps are otherwise waiting.
  auto o_epilogue = [&](int hh) {
    const int quarter = warp;   // control roles 0..3 sit on hardware warps = role (mod 4)
    uint8_t* ctx_tile = smem_ctx + quarter * (32 * 128);
    const uint32_t ctx_dst = ptx::smem_u32(ctx_tile) + lane * 128;
    ptx::mbar_wait(o_full, hh & 1);
    ptx::tc_fence_after();
    uint32_t o[2][32];
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kTmemO;
    ptx::tmem_ld_x32(t_o, o[0]);
    ptx::tmem_ld_x32(t_o + 32, o[1]);
    ptx::tmem_ld_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(o_free);
    // the tile of the previous head has been read out by its TMA store (only the issuing lane has a bulk group)
    ptx::tma_store_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {   // 16-byte chunk k = context columns 8 k .. 8 k + 7 of this row
      const uint32_t* v = &o[k >> 2][(k & 3) * 8];
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ctx_dst + ((k ^ (lane & 7)) << 4)),
                   "r"(pack_bf16x2_u(v[0], v[1])), "r"(pack_bf16x2_u(v[2], v[3])), "r"(pack_bf16x2_u(v[4], v[5])),
                   "r"(pack_bf16x2_u(v[6], v[7]))
                   : "memory");
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (ptx::elect_one()) {
      ptx::tma_store_3d(&tmap_ctx, ctx_tile, (h0 + hh) * D, qt * BM + quarter * 32, b);
      ptx::tma_store_commit();
    }
    __syncwarp();
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (+ context epilogue of lane quarter 0)
    for (int h = 0; h < nh + 2; ++h) {   // h: head index local to this CTA (absolute head h0 + h)
      if (h < nh) {
        const int st = h & 1;
        const uint32_t ph = (h >> 1) & 1;
        uint8_t* sq = smem + st * kStageBytes;
        uint8_t* sk = sq + kQBytes;
        uint8_t* sv = sk + kKVBytes;
        ptx::mbar_wait(&qk_empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&qk_full[st], kQBytes + kv_tx);
          ptx::tma_load_2d(sq, &tmap_q, &qk_full[st], (h0 + h) * D, row0 + qt * BM);
          ptx::tma_load_2d(sk, &tmap_kv, &qk_full[st], p.d + (h0 + h) * D, row0);
          ptx::tma_load_2d(sk + half_rows * 128, &tmap_kv, &qk_full[st], p.d + (h0 + h) * D, row0 + half_rows);
        }
        __syncwarp();
        ptx::mbar_wait(&v_empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&v_full[st], kv_tx);
          ptx::tma_load_2d(sv, &tmap_kv, &v_full[st], 2 * p.d + (h0 + h) * D, row0);
          ptx::tma_load_2d(sv + half_rows * 128, &tmap_kv, &v_full[st], 2 * p.d + (h0 + h) * D, row0 + half_rows);
        }
        __syncwarp();
      }
      // P V of head h - 2 has just released this V slot (or the loop is draining): its context is complete
      if (h >= 2) o_epilogue(h - 2);
    }
    ptx::tma_store_wait_read<0>();
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer (+ context epilogue of lane quarter 1)
    const uint32_t idesc_qk = ptx::make_idesc_bf16(BM, static_cast<uint32_t>(KP), 0, 0);
    // P V and the head average run with FP16 operands: P is fp16, and the qkv GEMM writes the V third of its output in
    // fp16 for this kernel (kind::f16 traps on mixed A / B formats: measured); QK^T stays bf16
    const uint32_t idesc_pv = ptx::make_idesc_f16kind(BM, D, 0, 0, 0, 1);  // B (= V) is MN-major
    const uint32_t sp = ptx::smem_u32(smem_p);
    const int ksteps = KP >> 4;
    const uint32_t idesc_avg = ptx::make_idesc_f16kind(BM, 16, 0, 0, 0, 0);
    // A: P K-block (ks / 4), +32 B per 16 keys inside the swizzle span (descriptor addresses are in 16-B units:
    // +2 per 16 keys, +1024 = 16 KB to the next K-block).
    const uint64_t dp0 = ptx::make_smem_desc_sw128(sp, 16, 1024);
    const uint64_t did = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_id), 16, 1024);
    auto issue_qk = [&](int h) {
      const int st = h & 1;
      ptx::mbar_wait(&qk_full[st], (h >> 1) & 1);
      if (h > 0) ptx::mbar_wait(s_free, (h - 1) & 1);
      ptx::tc_fence_after();
      const uint32_t sq = ptx::smem_u32(smem + st * kStageBytes);
      const uint64_t dq = ptx::make_smem_desc_sw128(sq, 16, 1024);
      const uint64_t dk = ptx::make_smem_desc_sw128(sq + kQBytes, 16, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16_ss(tmem_base + kTmemS, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        ptx::umma_commit(&qk_empty[st]);   // Q / K of this stage may be reloaded (head h + 2)
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int h = 0; h < nh; ++h) {
      if (h + 1 < nh) issue_qk(h + 1);
      const int st = h & 1;
      ATTN_TS(16);
      ptx::mbar_wait(&v_full[st], (h >> 1) & 1);
      ptx::mbar_wait(p_full, h & 1);
      ATTN_TS(17);
      if (h > 0) ptx::mbar_wait(o_free, (h - 1) & 1);
      ATTN_TS(18);
      ptx::tc_fence_after();
      const uint32_t sv = ptx::smem_u32(smem + st * kStageBytes) + kQBytes + kKVBytes;
      // B = V rows [16 ks, 16 ks + 16), MN-major: 8-key groups 1024 B apart (SBO); the single 64-wide MN group makes
      // LBO irrelevant; +128 (= 2 KB) per 16 keys.
      const uint64_t dv0 = ptx::make_smem_desc_sw128(sv, 1024, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kMaxKSteps; ++ks) {
          if (ks < ksteps) {
            const uint64_t dp = dp0 + static_cast<uint64_t>((ks >> 2) * (kPBlockBytes >> 4) + 2 * (ks & 3));
            ptx::umma_bf16_ss(tmem_base + kTmemO, dp, dv0 + static_cast<uint64_t>(ks * (2048 >> 4)), idesc_pv,
                              ks != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(&v_empty[st]);
        ptx::umma_commit(o_full);
        if (want_avg) {
          // Pbar[:, 16 ks + n] += sum_k P[:, 16 ks + k] * I[n, k]: one M = 128, N = 16, K = 16 instruction per 16 keys
#pragma unroll
          for (int ks = 0; ks < kMaxKSteps; ++ks) {
            if (ks < ksteps) {
              const uint64_t dp = dp0 + static_cast<uint64_t>((ks >> 2) * (kPBlockBytes >> 4) + 2 * (ks & 3));
              ptx::umma_bf16_ss(tmem_base + kTmemAvg + 16 * ks, dp, did, idesc_avg, h != 0 ? 1u : 0u);
            }
          }
        }
        ptx::umma_commit(p_free);
      }
      __syncwarp();
      ATTN_TS(19);
      o_epilogue(h);   // P V was issued ahead of the head-average MMAs: O completes while those are being issued
    }
    ptx::tma_store_wait_read<0>();
  } else if (warp == 2) {
    // ------------------------------------------------------------ context epilogue of lane quarter 2
    for (int h = 0; h < nh; ++h) o_epilogue(h);
    ptx::tma_store_wait_read<0>();
  } else if (warp == 3) {
    // ------------------------------------------------------------ CLS-row writer (+ context epilogue of lane quarter 3)
    // Row 0 of the probabilities (fp32) is staged in smem by the four warps that own it; this otherwise idle warp
    // streams it to HBM with coalesced stores, off the softmax warps' critical path.
    const bool do_cls = p.cls_map != nullptr && qt == 0;
    for (int h = 0; h < nh; ++h) {
      if (do_cls) {
        ptx::mbar_wait(cls_full, h & 1);
        float* cp = p.cls_map + (static_cast<size_t>(b) * p.H + h0 + h) * p.N;
        // staged: e = exp2((s - m_t) c) of row 0 in fp32 (written by the owning lanes during their exponential pass,
        // off their critical path) and the four column groups' factors f_t; p = e f_t is the same fp32 product the
        // softmax threads used to form
        for (int j = lane; j < p.N; j += 32) cp[j] = cls_stage[j] * cls_factor[(j >> 3) & (kColGroups - 1)];
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(cls_free);
      }
      o_epilogue(h);
    }
    ptx::tma_store_wait_read<0>();
  } else if (warp >= kCtrlWarps) {
    // ------------------------------------------------------------ softmax / epilogue
    const int cg = (warp - kCtrlWarps) >> 2;    // column group: which quarter of the key granules
    const int quarter = warp & 3;               // TMEM lane quarter
    const int r = quarter * 32 + lane;          // row inside the tile
    const int qrow = qt * BM + r;               // token index inside the image
    const bool row_ok = qrow < p.N;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int ngran = KP >> 3;                  // 8-key granules in a row
    // Granules are dealt round-robin: this thread owns g = 4 c + cg, c = 0 .. nmy-1.  Then g % 8 = cg | 4 (c & 1),
    // g / 8 = c / 2, g / 4 = c, g % 4 = cg: every TMEM column and every swizzled smem address below is one of two
    // per-thread bases plus a compile-time offset (contiguous ranges needed ~13 address instructions per granule).
    const int nmy = ngran > cg ? (ngran - cg + kColGroups - 1) / kColGroups : 0;   // granules of this thread
#define OWNS(c) (kFull ? ((c) < kMaxGran - 1 || cg < 2) : ((c) < nmy))
    const float inv_h = 1.0f / static_cast<float>(p.H);
    const bool has_pad = nmy > 0 && (kColGroups * (nmy - 1) + cg) * 8 + 8 > p.N;   // warp-uniform: owns padded keys
    const int valid_last = p.N - (kColGroups * (nmy - 1) + cg) * 8;   // valid keys in this thread's last granule (< 8 iff has_pad)
    const bool want_cls = p.cls_map != nullptr && qt == 0 && quarter == 0;
    const uint32_t p_row = ptx::smem_u32(smem_p) + r * 128;
    const int sw = r & 7;
    const uint32_t sw4 = static_cast<uint32_t>(sw) << 4;
    const uint32_t p_even = p_row + ((static_cast<uint32_t>(cg) << 4) ^ sw4);        // chunk cg     of K-block c / 2
    const uint32_t p_odd = p_row + ((static_cast<uint32_t>(cg | 4) << 4) ^ sw4);     // chunk cg + 4 of K-block c / 2
    const uint32_t t_s = lane_base + kTmemS + cg * 8;      // granule c: + 32 c columns
    const uint32_t t_avg = lane_base + kTmemAvg + cg * 8;

    for (int h = 0; h < nh; ++h) {
      // ---- this thread's part of the S row -> registers (single TMEM read), then release the S columns.  The read is
      //      issued in two halves so that the maximum of the first granules is taken while the others are in flight.
      uint32_t s[kMaxGran][8];
      ATTN_TS(0);
      // (row-0 warps: the class-token stage of the previous head must have been streamed out before the exponential
      //  pass below overwrites it; warp 3 finished that long ago, this is the barrier's fast path)
      if (want_cls && h > 0) ptx::mbar_wait(cls_free, (h - 1) & 1);
      ptx::mbar_wait(s_full, h & 1);
      ATTN_TS(1);
      ptx::tc_fence_after();
      constexpr int kFirst = 4;   // granules of the first half
#pragma unroll
      for (int c = 0; c < kFirst; ++c) {
        if OWNS(c) {
          ptx::tmem_ld_x8(t_s + c * 32, s[c]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) s[c][j] = 0xff800000u;   // -inf: no score, exp2 gives 0
        }
      }
      ptx::tmem_ld_wait();
#pragma unroll
      for (int c = kFirst; c < kMaxGran; ++c) {
        if OWNS(c) {
          ptx::tmem_ld_x8(t_s + c * 32, s[c]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) s[c][j] = 0xff800000u;
        }
      }
      // ---- padded keys (>= N) count as -inf.  KP - N < 16, so they sit in the last two granules of the row, each of
      //      which is the LAST granule of its owner: one granule per thread to patch, and only in warps that own one
      auto patch = [&](uint32_t (&row)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j >= valid_last) row[j] = 0xff800000u;  // -inf: exp2 gives 0, max ignores it
      };
      float mx = -INFINITY;
      if (has_pad && nmy <= kFirst) {   // warp-uniform; compile-time row indices keep s[][] in registers
        switch (nmy) {
          case 4: patch(s[3]); break;
          case 3: patch(s[2]); break;
          case 2: patch(s[1]); break;
          default: patch(s[0]); break;
        }
      }
#pragma unroll
      for (int c = 0; c < kFirst; ++c) {
        if OWNS(c) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) mx = ptx::fmax3(mx, __uint_as_float(s[c][j]), __uint_as_float(s[c][j + 1]));
        }
      }
      ptx::tmem_ld_wait();
      ATTN_TS(2);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_free);
      ATTN_TS(3);
      static_assert(kMaxGran == 7, "one case per possible granule count");
      if (has_pad && nmy > kFirst) {
        switch (nmy) {
          case 7: patch(s[6]); break;
          case 6: patch(s[5]); break;
          default: patch(s[4]); break;
        }
      }
#pragma unroll
      for (int c = kFirst; c < kMaxGran; ++c) {
        if OWNS(c) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) mx = ptx::fmax3(mx, __uint_as_float(s[c][j]), __uint_as_float(s[c][j + 1]));
        }
      }

      // ---- e = exp2((s - m_t) * c) with the maximum m_t of this thread's OWN columns: no exchange is needed before
      //      the exponentials.  Softmax is invariant to the shift, the four threads of a row reconcile afterwards:
      //      p = e * f_t,  f_t = exp2((m_t - M) c) / sum_u(sum_u exp2((m_u - M) c)),  M = max_u m_u.
      //      The exponentials are packed to fp16 pairs as they are produced (the MUFU bounds this pass, the conversions
      //      ride along on the ALU); the fp32 values of query row 0 go to the class-token stage from the owning lanes.
      // a thread whose columns are all padding (tiny test shapes) must not produce -inf - -inf
      const float mxs = (mx == -INFINITY) ? 0.f : mx * p.scale_log2;
      ATTN_TS(4);
      float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
      uint32_t e16[kMaxGran][4];
#pragma unroll
      for (int c = 0; c < kMaxGran; ++c) {
        if (!OWNS(c)) {   // (compile-time for all but the last granule when kFull)
          e16[c][0] = e16[c][1] = e16[c][2] = e16[c][3] = 0u;
          continue;
        }
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          const float v0 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j]), p.scale_log2, -mxs));
          const float v1 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j + 1]), p.scale_log2, -mxs));
          const float v2 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j + 2]), p.scale_log2, -mxs));
          const float v3 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j + 3]), p.scale_log2, -mxs));
          ps0 += v0, ps1 += v1, ps2 += v2, ps3 += v3;
          e16[c][j >> 1] = pack_f16x2_f(v0, v1), e16[c][(j >> 1) + 1] = pack_f16x2_f(v2, v3);
          if (kHeads) {
            s[c][j] = __float_as_uint(v0), s[c][j + 1] = __float_as_uint(v1);
            s[c][j + 2] = __float_as_uint(v2), s[c][j + 3] = __float_as_uint(v3);
          }
          if (want_cls) {   // warp-uniform; only lane 0 (query row 0) stores
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"r"(
                             ptx::smem_u32(cls_stage + (kColGroups * c + cg) * 8 + j)),
                         "f"(v0), "f"(v1), "f"(v2), "f"(v3), "r"(static_cast<uint32_t>(lane))
                         : "memory");
          }
        }
      }
      ATTN_TS(5);
      // exchange buffers alternate with the head parity: with one barrier per head a fast thread may already write
      // head h+1's values while a slow one still reads head h's
      float* red_max = red + (h & 1) * (2 * kColGroups * BM);  // [group][row]
      float* red_sum = red_max + kColGroups * BM;
      red_max[cg * BM + r] = (mx == -INFINITY) ? -INFINITY : mxs;   // already in the exp2 domain
      red_sum[cg * BM + r] = (ps0 + ps1) + (ps2 + ps3);
      // The exchange concerns the four warps of ONE lane quarter only -- they sit on one SM sub-partition -- so the
      // barrier is per quarter: the quarters do not wait for each other here.
      asm volatile("bar.sync %0, 128;" ::"r"(8 + quarter) : "memory");
      ATTN_TS(6);
      float inv;
      {
        const float m0 = red_max[r], m1 = red_max[BM + r], m2 = red_max[2 * BM + r], m3 = red_max[3 * BM + r];
        const float M = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        // groups without a valid column have m = -inf and sum = 0: exp2(-inf) = 0 keeps them out
        const float tot = (red_sum[r] * ptx::ex2_approx(m0 - M) + red_sum[BM + r] * ptx::ex2_approx(m1 - M)) +
                          (red_sum[2 * BM + r] * ptx::ex2_approx(m2 - M) + red_sum[3 * BM + r] * ptx::ex2_approx(m3 - M));
        inv = ptx::ex2_approx(mxs - M) * ptx::rcp_approx(tot);
        if (mx == -INFINITY) inv = 0.f;
      }
      if (want_cls && lane == 0) {
        cls_factor[cg] = inv;
        ptx::mbar_arrive(cls_full);  // release: stage + factor are visible to warp 3 once the four owners have arrived
      }
      const uint32_t inv16 = pack_f16x2_f(inv, inv);
      ATTN_TS(9);

      // ---- the P tile is about to be overwritten: the previous head's P V and Pbar MMAs (issued a whole softmax
      //      pass ago) must have retired
      if (h > 0) ptx::mbar_wait(p_free, (h - 1) & 1);
      ATTN_TS(10);

      // ---- p = e * f_t as packed fp16 multiplies -> fp16 P tile (swizzled K-major A operand of P V and of the
      //      head-average MMAs).  All kMaxGran granules are multiplied (a thread with fewer granules scales zeros) and
      //      only the store is predicated: per-granule branches cost more issue slots than the arithmetic they skip.
#pragma unroll
      for (int c = 0; c < kMaxGran; ++c) {
        // keys [8 g, 8 g + 8), g = 4 c + cg: K-block c / 2, 16-byte chunk (cg | 4 (c & 1)) ^ (row % 8) of this row
        const uint32_t dst = ((c & 1) ? p_odd : p_even) + (c >> 1) * kPBlockBytes;
        const uint32_t w0 = hmul2_u(e16[c][0], inv16), w1 = hmul2_u(e16[c][1], inv16);
        const uint32_t w2 = hmul2_u(e16[c][2], inv16), w3 = hmul2_u(e16[c][3], inv16);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t}" ::"r"(dst), "r"(w0),
                     "r"(w1), "r"(w2), "r"(w3), "r"(static_cast<uint32_t>OWNS(c))
                     : "memory");
      }
      ATTN_TS(12);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
      ATTN_TS(7);


      if (kHeads) {
        if (p.head_map != nullptr && row_ok) {
          float* hp = p.head_map + ((static_cast<size_t>(b) * p.H + h0 + h) * p.N + qrow) * p.ldmap;
#pragma unroll
          for (int c = 0; c < kMaxGran; ++c) {
            if OWNS(c) {
              float4* dst = reinterpret_cast<float4*>(hp + (kColGroups * c + cg) * 8);
              dst[0] = make_float4(__uint_as_float(s[c][0]) * inv, __uint_as_float(s[c][1]) * inv,
                                   __uint_as_float(s[c][2]) * inv, __uint_as_float(s[c][3]) * inv);
              dst[1] = make_float4(__uint_as_float(s[c][4]) * inv, __uint_as_float(s[c][5]) * inv,
                                   __uint_as_float(s[c][6]) * inv, __uint_as_float(s[c][7]) * inv);
            }
          }
        }
      }
      ATTN_TS(8);
    }
    // head-averaged map rows -> HBM, once per (image, query tile): Pbar holds the SUM over heads.  The tile goes
    // through shared memory (the Q/K/V stages are dead by now) as 32-column slabs of 128 rows x 128 B in the 128-B
    // swizzle and leaves with one TMA store per slab; rows beyond the image and columns beyond ldmap are clipped by
    // the [B][N][ldmap] tensor map.  (32-B-per-thread global stores from here cost ~14% of the kernel: every
    // warp-wide STG touched 32 different lines.)
    if (want_avg) {
      // The head-average MMAs of the LAST head are issued behind the commit of o_full: their own completion is p_free's
      // last phase (which also says that P V of the last head no longer reads the stage area reused below).  Without this wait the read below raced the tensor pipe (rarely lost:
      // the one run-to-run difference of round 1's batch-256 reproducibility check).
      ptx::mbar_wait(p_free, (nh - 1) & 1);
      ptx::tc_fence_after();
      uint32_t a[kMaxGran][8];
#pragma unroll
      for (int c = 0; c < kMaxGran; ++c)
        if OWNS(c) ptx::tmem_ld_x8(t_avg + c * 32, a[c]);
      ptx::tmem_ld_wait();
      const uint32_t slab0 = ptx::smem_u32(smem) + r * 128;
#pragma unroll
      for (int c = 0; c < kMaxGran; ++c) {
        if OWNS(c) {
          // granule g = 4 c + cg: 32-column slab c, 16-byte chunks 2 cg and 2 cg + 1 of this row
          const uint32_t dst = slab0 + c * (BM * 128);
          const int ch = 2 * cg;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((ch ^ sw) << 4)),
                       "f"(__uint_as_float(a[c][0]) * inv_h), "f"(__uint_as_float(a[c][1]) * inv_h),
                       "f"(__uint_as_float(a[c][2]) * inv_h), "f"(__uint_as_float(a[c][3]) * inv_h)
                       : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((ch + 1) ^ sw) << 4)),
                       "f"(__uint_as_float(a[c][4]) * inv_h), "f"(__uint_as_float(a[c][5]) * inv_h),
                       "f"(__uint_as_float(a[c][6]) * inv_h), "f"(__uint_as_float(a[c][7]) * inv_h)
                       : "memory");
        }
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (warp == kCtrlWarps && ptx::elect_one()) {
        const int nslabs = (KP + 31) >> 5;
        for (int j = 0; j < nslabs; ++j) {
          if (split_cta && p.part_images == 0) ptx::tma_reduce_add_3d(&tmap_avg, smem + j * (BM * 128), 32 * j, qt * BM, b);
          else ptx::tma_store_3d(&tmap_avg, smem + j * (BM * 128), 32 * j, qt * BM, avg_img + b);
        }
        ptx::tma_store_commit();
      }
      __syncwarp();
    }
    // shared memory must outlive the bulk stores' READS only (the writes land on their own; kernel completion makes
    // them visible): threads without an outstanding bulk group fall through
    ptx::tma_store_wait_read<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
#undef OWNS
}

}  // namespace vitb200
