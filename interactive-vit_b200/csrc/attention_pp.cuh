// Fused multi-head self-attention with attention-map emission, TWO HEADS IN FLIGHT per SM ("ping-pong"): the
// production kernel for 197-token sequences (193 <= N <= 200, head dim 64, KP = 208).  Same arithmetic, same TMEM /
// shared-memory plan and the same outputs as attention_kernel (attention.cuh: torch/nn/functional.py:6630-6659 via
// torchvision's EncoderBlock, vision_transformer.py:110-119, need_weights=True); what changes is how the softmax is
// scheduled.
//
// Why: in attention_kernel all 16 softmax warps work on ONE head in lock-step (four threads per query row).  Its
// traced head period is ~4,250 cycles of which only the exponential pass (~1,700 cycles) is bound by a pipe (the
// MUFU, 4 results per cycle and SM sub-partition); the rest is a dependent chain of barrier waits, TMEM round
// trips and a shared-memory exchange during which each sub-partition has nothing else to issue (round-1 VERDICT:
// tensor pipe 15 %, XU 31 %, issue slots 47 %).  TMEM has no room for a second S buffer (O 64 + S 208 + Pbar 208
// columns), and the register file none for a second S row per thread.
//
// How: the 16 softmax warps form two GROUPS of 8 (two threads per query row each); group g owns the heads
// h = g, g + 2, ...  The single S buffer is time-shared: a group copies its share of the S row out of TMEM
// IMMEDIATELY -- pass 1a: thread-local row maximum (TMEM read, FMNMX3 only), pass 1b: second TMEM read,
// d = (s - m_t) c <= 0 packed to FP16 pairs (52 registers per thread) -- and releases S after ~600 cycles instead
// of holding it through the exponentials, so S = Q K^T of the next head is produced for the OTHER group while this
// group runs its MUFU pass (ex2.approx.f16x2 straight on the packed differences, row sums accumulated in fp32 with
// mixed-precision adds), exchanges (max, sum) with the row's partner thread, normalises with packed fp16
// multiplies and stores the fp16 P tile.  On every sub-partition one group's MUFU pass now runs under the other
// group's latency-bound phases.  d has an absolute error <= 2^-12 for the entries within a factor 2 of the row
// maximum (relative error of p 1.7e-4, below the fp16 rounding of p itself: 4.9e-4) and falls off with p.
//
// Two warps issue MMAs (role 1: P V and the head average, role 2: S = Q K^T) because with two heads in flight the
// order of "S can be produced" and "P is ready" is no longer fixed and a single issuer serialised them.
//
// Roles: role 0 TMA producer, role 1 UMMA issuer (P V, Pbar), role 2 TMEM allocator and UMMA issuer (Q K^T), role 3
// class-token row writer -- each of them also owns the context epilogue of one TMEM lane quarter; roles 4..19 softmax:
// quarter = role % 4 (TMEM lane quarter = SM sub-partition), group = ((role - 4) / 4) % 2, column half
// c2 = (role - 4) / 8: the thread owns the 8-key granules 2 c + c2, c = 0..12 (granule 25 = keys 200..207 is
// always padding for N <= 200 and is zeroed once).
#pragma once
#include "attention.cuh"

namespace vitb200 {

namespace attn_pp_cfg {
using namespace attn_cfg;
constexpr int kGran = 13;                                   // granules per thread (c2 == 1: 12)
constexpr int kGranA = 8;                                   // ... of which block A (keys < 128); block B: 5 (4)
constexpr int kGranB = kGran - kGranA;
constexpr int kKeysA = 16 * kGranA;                         // block A of both column halves = keys [0, 128)
constexpr int kMaxTokens = 200;                             // granule 25 must be pure padding
constexpr int kMinTokens = 193;                             // KP = 208 and granule 24 holds a valid key
constexpr int kClsStagePP = 2 * kMaxTokens * 2;             // [group][200] fp16 exponentials of query row 0
constexpr int kClsFactorPP = 2 * 2 * 2 * 4;                 // [group][half][block] normalising factors
constexpr int kNumBars = 24;
constexpr int kSmemBytesPP = 2 * kStageBytes + kPBytes + kCtxStageBytes + kIdentBytes + kClsStagePP + kClsFactorPP +
                             kNumBars * 8 + 8;
static_assert(kSmemBytesPP <= 227 * 1024, "attention (ping-pong): shared memory budget");
static_assert((kClsStagePP + kClsFactorPP) % 8 == 0, "mbarriers are 8-byte aligned");
}  // namespace attn_pp_cfg

// Phase tracing (tracing build only, tools/attn_trace_pp.py): CTA 0, lane 0 of one softmax warp per group (quarter 0,
// column half 0) stamps row h of g_attn_trace for its own heads (slots 0..15), the MMA warp slots 16..23.
#ifdef VITB200_ATTN_TRACE
#ifndef PP_TRACE_Q
#define PP_TRACE_Q 0
#endif
#ifndef PP_TRACE_C2
#define PP_TRACE_C2 0
#endif
#define PP_TS(slot) do { if (blockIdx.x == 0 && lane == 0 && quarter == PP_TRACE_Q && c2 == PP_TRACE_C2) g_attn_trace[(h) * 32 + (slot)] = clock64(); } while (0)
#define PP_TS_MMA(hh, slot) do { if (blockIdx.x == 0 && lane == 0) g_attn_trace[(hh) * 32 + (slot)] = clock64(); } while (0)
#else
#define PP_TS(slot) do { } while (0)
#define PP_TS_MMA(hh, slot) do { } while (0)
#endif

__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
// Four packed exponentials in place, predicated on `gate != 0xffffffff` (always true: `gate` is an fp16 pair of finite
// exponentials produced two granules earlier).  The predicate is a REAL dependency: without it ptxas hoists all 104
// MUFU.EX2.F16 of a pass in front of the PRMTs that merge their half results and spills ~50 registers.
__device__ __forceinline__ void ex2_f16x2_x4_gated(uint32_t (&v)[4], uint32_t gate) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0xffffffff;\n\t@q ex2.approx.f16x2 %0, %0;\n\t@q ex2.approx.f16x2 %1, %1;\n\t"
      "@q ex2.approx.f16x2 %2, %2;\n\t@q ex2.approx.f16x2 %3, %3;\n\t}"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3])
      : "r"(gate));
}
// acc0 += float(lo half), acc1 += float(hi half): one mixed-precision add each (FHADD), no conversion instruction
__device__ __forceinline__ void add_f16x2_to_f32(float& acc0, float& acc1, uint32_t v) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.f16 %0, lo, %0;\n\tadd.rn.f32.f16 %1, hi, %1;\n\t}"
      : "+f"(acc0), "+f"(acc1)
      : "r"(v));
}

__global__ void __launch_bounds__(attn_cfg::kThreads, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box 64 x 128 over qkv [B*N, 3d]
                    const __grid_constant__ CUtensorMap tmap_kv,  // box 64 x (KP/2) over the same tensor
                    const __grid_constant__ CUtensorMap tmap_ctx, // box 64 x 32 x 1 over ctx viewed as [B][N][d]
                    const __grid_constant__ CUtensorMap tmap_avg, // fp32, box 32 x 128 x 1 over avg_map [B][N][ldmap]
                    AttnParams p) {
  using namespace attn_pp_cfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_p = smem + 2 * kStageBytes;
  uint8_t* smem_ctx = smem_p + kPBytes;      // 4 quarter tiles of 32 rows x 128 B, 128-B swizzle
  uint8_t* smem_id = smem_ctx + kCtxStageBytes;
  __half* cls_stage = reinterpret_cast<__half*>(smem_id + kIdentBytes);
  float* cls_factor = reinterpret_cast<float*>(smem_id + kIdentBytes + kClsStagePP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_id + kIdentBytes + kClsStagePP + kClsFactorPP);
  uint64_t* qk_full = bars;          // [stage] Q and K of a head landed
  uint64_t* v_full = bars + 2;       // [stage] V of a head landed
  uint64_t* qk_empty = bars + 4;     // [stage] Q / K consumed by QK^T
  uint64_t* v_empty = bars + 6;      // [stage] V consumed by P V
  // Everything a group waits on or signals has one barrier PER GROUP: a barrier shared by both groups would let a
  // group that runs ahead mistake the other group's phase (same parity two phases later) for its own.
  uint64_t* s_full = bars + 8;       // [half][group] S = QK^T of one of the group's heads complete: keys [0, 128) / [128, 208)
  uint64_t* s_free = bars + 22;      // [group] S copied out of TMEM by the group's 8 warps
  uint64_t* p_full = bars + 12;      // [group] fp16 P tile of the group's head written
  uint64_t* p_free = bars + 14;      // [group] P V and Pbar += P of the group's head complete: P tile reusable
  uint64_t* o_full = bars + 16;      // O = P V complete
  uint64_t* o_free = bars + 17;      // O read by the four control warps
  uint64_t* cls_full = bars + 18;    // [group] exponentials of query row 0 + factors staged
  uint64_t* cls_free = bars + 20;    // [group] ... and written out by role 3
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int warp = static_cast<int>((threadIdx.x >> 5) + kCtrlWarps) % (kThreads / 32);   // role (see attention.cuh)
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
  constexpr int KP = KP_MAX;
  constexpr int half_rows = KP >> 1;
  constexpr uint32_t kv_tx = static_cast<uint32_t>(KP) * D * 2;
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
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_full[2 + i], 1);
      ptx::mbar_init(&s_free[i], kSoftmaxWarps / 2);   // one arrival per warp of the group
      ptx::mbar_init(&p_full[i], kSoftmaxWarps / 2);
      ptx::mbar_init(&p_free[i], 1);
      ptx::mbar_init(&cls_full[i], 2);                 // the two column halves of query row 0
      ptx::mbar_init(&cls_free[i], 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_free, kCtrlWarps);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  if (warp == 3) {
    // 16 x 16 fp16 identity, K-major 128-B rows in the 128-B swizzle (B operand of Pbar += P I16; see attention.cuh)
    uint4* id = reinterpret_cast<uint4*>(smem_id);
    for (int i = lane; i < kIdentBytes / 16; i += 32) {
      const int n = i >> 3, c = (i & 7) ^ (n & 7);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((n >> 3) == c) {
        const uint32_t one = 0x3C00u << (16 * (n & 1));
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
  const int row0 = b * p.N;

  // Context epilogue of one TMEM lane quarter (control role q reads lanes 32 q .. 32 q + 31): O of head hh is final
  // (P was normalised before the MMA) -> bf16 -> this quarter's smem tile -> one TMA store (rows beyond the image clip).
  auto o_epilogue = [&](int hh) {
    const int quarter = warp;
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
    ptx::tma_store_wait_read<0>();   // the previous head's tile has been read out by its TMA store
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
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
    // Order of this warp's events in the steady state: the Q / K slot of head i frees when QK^T of head i - 2 completes
    // (early in that head), the V slot of head i - 1 and O of head i - 3 when P V of head i - 3 completes (~3,000
    // cycles later), then the Q / K slot of head i + 1, ...  The loop takes them in that order.  (Traced: with the
    // one-head kernel's order -- Q / K and V of the SAME head back to back -- the Q / K loads of head i + 1 queued
    // behind the V wait and landed ~1,500 cycles after the S buffer had been handed back: qkv comes from HBM.)
    for (int i = 0; i < nh + 3; ++i) {
      if (i < nh) {
        const int st = i & 1;
        uint8_t* sq = smem + st * kStageBytes;
        uint8_t* sk = sq + kQBytes;
        ptx::mbar_wait(&qk_empty[st], ((i >> 1) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&qk_full[st], kQBytes + kv_tx);
          ptx::tma_load_2d(sq, &tmap_q, &qk_full[st], (h0 + i) * D, row0 + qt * BM);
          ptx::tma_load_2d(sk, &tmap_kv, &qk_full[st], p.d + (h0 + i) * D, row0);
          ptx::tma_load_2d(sk + half_rows * 128, &tmap_kv, &qk_full[st], p.d + (h0 + i) * D, row0 + half_rows);
        }
        __syncwarp();
      }
      if (i >= 1 && i <= nh) {
        const int j = i - 1, st = j & 1;
        uint8_t* sv = smem + st * kStageBytes + kQBytes + kKVBytes;
        ptx::mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&v_full[st], kv_tx);
          ptx::tma_load_2d(sv, &tmap_kv, &v_full[st], 2 * p.d + (h0 + j) * D, row0);
          ptx::tma_load_2d(sv + half_rows * 128, &tmap_kv, &v_full[st], 2 * p.d + (h0 + j) * D, row0 + half_rows);
        }
        __syncwarp();
      }
      if (i >= 3) o_epilogue(i - 3);   // P V of head i - 3 has just released the V slot (or the loop is draining)
    }
    ptx::tma_store_wait_read<0>();
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer of P V and of the head average (+ context
    // epilogue of lane quarter 1).  Two warps issue MMAs in this kernel: with two heads in flight the order of
    // "S of the next head can be produced" and "P of the oldest head is ready" is not fixed, and one issuer that blocks
    // on either (or is busy pushing the 26 P V / average instructions of a head through the MMA queue, ~800 cycles, or
    // in its epilogue, ~550) delays the other -- traced: QK^T was issued 1,400-2,300 cycles after S had been released.
    // The two issuers touch disjoint TMEM columns (S | O, Pbar) and each commits its own instructions.
    constexpr uint32_t idesc_pv = ptx::make_idesc_f16kind(BM, D, 0, 0, 0, 1);   // fp16 P x fp16 V (MN-major)
    constexpr uint32_t idesc_avg = ptx::make_idesc_f16kind(BM, 16, 0, 0, 0, 0);
    constexpr int ksteps = KP >> 4;
    const uint32_t sp = ptx::smem_u32(smem_p);
    const uint64_t dp0 = ptx::make_smem_desc_sw128(sp, 16, 1024);
    const uint64_t did = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_id), 16, 1024);
    for (int h = 0; h < nh; ++h) {
      const int st = h & 1;
      ptx::mbar_wait(&v_full[st], (h >> 1) & 1);
      ptx::mbar_wait(&p_full[h & 1], (h >> 1) & 1);
      if (h > 0) ptx::mbar_wait(o_free, (h - 1) & 1);
      ptx::tc_fence_after();
      PP_TS_MMA(h, 17);
      const uint32_t sv = ptx::smem_u32(smem + st * kStageBytes) + kQBytes + kKVBytes;
      const uint64_t dv0 = ptx::make_smem_desc_sw128(sv, 1024, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t dp = dp0 + static_cast<uint64_t>((ks >> 2) * (kPBlockBytes >> 4) + 2 * (ks & 3));
          ptx::umma_bf16_ss(tmem_base + kTmemO, dp, dv0 + static_cast<uint64_t>(ks * (2048 >> 4)), idesc_pv,
                            ks != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&v_empty[st]);
        ptx::umma_commit(o_full);
        if (want_avg) {
#pragma unroll
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t dp = dp0 + static_cast<uint64_t>((ks >> 2) * (kPBlockBytes >> 4) + 2 * (ks & 3));
            ptx::umma_bf16_ss(tmem_base + kTmemAvg + 16 * ks, dp, did, idesc_avg, h != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(&p_free[h & 1]);
      }
      __syncwarp();
      PP_TS_MMA(h, 18);
      PP_TS_MMA(h, 19);
      o_epilogue(h);   // P V was issued ahead of the head-average MMAs: O completes while those are being issued
      PP_TS_MMA(h, 20);
    }
    ptx::tma_store_wait_read<0>();
  } else if (warp == 2) {
    // ------------------------------------------------------------ UMMA issuer of S = Q K^T (+ context epilogue of lane
    // quarter 2).  QK(h) needs the operands and the S buffer (copied out by the group of head h - 1); the epilogue of
    // head h needs O(h).  Which comes first depends on how the groups are running, so this warp polls both (test_wait
    // does not suspend; a short sleep between polls leaves the issue slots to this sub-partition's softmax warps).
    // S is PRODUCED in two halves with a barrier each -- keys [0, 128) = block A of every softmax thread first, then
    // keys [128, 208) -- so that the group starts copying block A out while the second half is still being computed
    // (the S buffer is handed back whole).  Order of this warp's events in the steady state (traced): QK(h + 1) ~1,300
    // cycles after S(h) became ready, the epilogue of head h - 1 ~1,100 later, QK(h + 2) a head period after QK(h + 1):
    // strictly alternating, so plain blocking waits in that order (no polling: the first version polled both events
    // and took issue slots from this sub-partition's softmax warps).  Neither wait can deadlock the other: everything
    // S(h)'s consumers and O(h - 1)'s producers need was issued earlier.
    constexpr uint32_t idesc_qa = ptx::make_idesc_bf16(BM, static_cast<uint32_t>(kKeysA), 0, 0);
    constexpr uint32_t idesc_qb = ptx::make_idesc_bf16(BM, static_cast<uint32_t>(KP - kKeysA), 0, 0);
    auto issue_qk = [&](int h) {
      const int st = h & 1;
      ptx::mbar_wait(&qk_full[st], (h >> 1) & 1);
      if (h > 0) ptx::mbar_wait(&s_free[(h - 1) & 1], ((h - 1) >> 1) & 1);
      ptx::tc_fence_after();
      PP_TS_MMA(h, 16);
      const uint32_t sq = ptx::smem_u32(smem + st * kStageBytes);
      const uint64_t dq = ptx::make_smem_desc_sw128(sq, 16, 1024);
      const uint64_t dk = ptx::make_smem_desc_sw128(sq + kQBytes, 16, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16_ss(tmem_base + kTmemS, dq + 2 * k, dk + 2 * k, idesc_qa, k != 0 ? 1u : 0u);
        ptx::umma_commit(&s_full[h & 1]);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)   // K rows 128..207: + 128 rows x 128 B = 1024 descriptor units
          ptx::umma_bf16_ss(tmem_base + kTmemS + kKeysA, dq + 2 * k, dk + (kKeysA * 128 >> 4) + 2 * k, idesc_qb, k != 0 ? 1u : 0u);
        ptx::umma_commit(&qk_empty[st]);   // Q / K of this stage may be reloaded (head h + 2)
        ptx::umma_commit(&s_full[2 + (h & 1)]);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int h = 0; h < nh; ++h) {
      if (h + 1 < nh) issue_qk(h + 1);
      if (h >= 1) o_epilogue(h - 1);
    }
    o_epilogue(nh - 1);
    ptx::tma_store_wait_read<0>();
  } else if (warp == 3) {
    // ------------------------------------------------------------ class-token row writer (+ context epilogue of quarter 3)
    const bool do_cls = p.cls_map != nullptr && qt == 0;
    for (int h = 0; h < nh; ++h) {
      if (do_cls) {
        const int g = h & 1;
        ptx::mbar_wait(&cls_full[g], (h >> 1) & 1);
        float* cp = p.cls_map + (static_cast<size_t>(b) * p.H + h0 + h) * p.N;
        const __half* e = cls_stage + g * kMaxTokens;
        // p = e f: the fp16 exponential the P tile is built from times the fp32 factor of the column half that owns key j
        for (int j = lane; j < p.N; j += 32) cp[j] = __half2float(e[j]) * cls_factor[(g * 2 + ((j >> 3) & 1)) * 2 + (j >= 16 * kGranA ? 1 : 0)];
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&cls_free[g]);
      }
      o_epilogue(h);
    }
    ptx::tma_store_wait_read<0>();
  } else {
    // ------------------------------------------------------------ softmax
    const int idx = warp - kCtrlWarps;
    const int quarter = idx & 3;                // TMEM lane quarter (= hardware warp % 4)
    const int sub = idx >> 2;                   // 0..3
    const int g = sub & 1;                      // group: heads g, g + 2, ...
    const int c2 = sub >> 1;                    // column half: granules 2 c + c2
    const bool own13 = c2 == 0;                 // owns granule 24 (keys 192..199); granule 25 is padding
    const int r = quarter * 32 + lane;          // row inside the tile
    const bool warp_has_rows = qt * BM + quarter * 32 < p.N;   // warp-uniform
    const int vlast = p.N - 192;                // valid keys in granule 24 (1..8)
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = lane_base + kTmemS + c2 * 8;   // granule c of this thread: + 16 c columns
    const bool want_cls = p.cls_map != nullptr && qt == 0 && quarter == 0;
    const uint32_t p_row = ptx::smem_u32(smem_p) + r * 128;
    const int sw = r & 7;
    // granule 2 c + c2 = keys of K-block c / 4, 16-byte chunk (2 (c % 4) + c2) ^ sw = (2 (c % 4)) ^ (c2 ^ sw)
    auto pb = [&](int x) { return p_row + (static_cast<uint32_t>((2 * x) ^ (c2 ^ sw)) << 4); };
    // (max, sum) exchange of the row's two threads: one 16-byte chunk per (group, half) in the UNUSED part of the P
    // tile's last K-block -- keys 192..207 occupy the logical chunks 0 and 1 of its 128-byte rows (the MMAs read 32
    // bytes per row and K step), chunks 2..5 carry this exchange (swizzled like the tile: conflict-free)
    auto red_addr = [&](int half) {
      return p_row + 3 * kPBlockBytes + (static_cast<uint32_t>((2 + 2 * g + half) ^ sw) << 4);
    };
    const uint32_t cls_dst = ptx::smem_u32(cls_stage + g * kMaxTokens) + c2 * 16;   // granule c: + 32 c bytes
    const uint32_t bar_id = 1 + quarter * 2 + g;

    if (g == 0 && c2 == 1) {
      // keys 200..207 of the P tile are never written by the head loop: zero them once (the fence in front of this
      // thread's first p_full arrival publishes them to the tensor pipe with the rest of the tile)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(pb(0) + 3 * kPBlockBytes), "r"(0u) : "memory");
    }

    int j = 0;
    for (int h = g; h < nh; h += 2, ++j) {
      const uint32_t ph = j & 1;
      PP_TS(0);
      if (want_cls && j > 0) ptx::mbar_wait(&cls_free[g], ph ^ 1);
      ptx::mbar_wait(&s_full[g], ph);
      ptx::tc_fence_after();
      PP_TS(1);
      if (!warp_has_rows) {
        // every row of this warp lies beyond the image (second query tile): keep the protocol going, no arithmetic
        ptx::tc_fence_before();
        __syncwarp();
        ptx::mbar_wait(&s_full[2 + g], ph);
        if (lane == 0) ptx::mbar_arrive(&s_free[g]);
        if (h > 0) ptx::mbar_wait(&p_free[g ^ 1], ((h - 1) >> 1) & 1);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[g]);
        continue;
      }

      // ---- one pass over S, TWO TMEM round trips per thread: block A = this thread's granules c = 0..7 (keys < 128),
      //      block B = c = 8..12 (keys >= 128; fewer, because block A's 32 packed registers stay live under it).  Per block: every read of the block in flight at once -> block maximum
      //      m -> d = (s - m) c <= 0 packed to fp16 pairs.  A thread-wide maximum would need the whole share (104
      //      fp32 registers) or a second pass over TMEM (a TMEM round trip costs 200-300 cycles under load: the first
      //      version of this kernel made 20 per head and was slower than the one-head kernel), so each block keeps
      //      its own maximum and row sum and is reconciled like a thread of its own.
      //      Each phase sits in a loop that runs ONCE (p.one == 1, unknown to the compiler): ptxas schedules within a
      //      loop body, so the reads of block B cannot be hoisted above the conversion of block A (it did exactly that,
      //      and delayed the fp16 packing into the exponential pass: 104 live fp32 values, 600 bytes of spills).
      uint32_t d16[kGran][4];
      float mxA, mxB, mxsA = 0.f, mxsB = 0.f;
#pragma unroll 1
      for (int once = 0; once < p.one; ++once) {
        uint32_t sA[kGranA][8], sB[kGranB][8];
#pragma unroll
        for (int c = 0; c < kGranA; ++c) ptx::tmem_ld_x8(t_s + c * 16, sA[c]);
        ptx::tmem_ld_wait();
        // (the chain starts with the LAST granule read: ptxas tracks every tcgen05.ld on its own scoreboard and would
        //  otherwise start consuming granule 0 after two reads and issue the rest one round trip at a time)
        float mxA2 = -INFINITY;   // two chains: FMNMX3 has a 2-cycle issue and ~5-cycle latency
        mxA = -INFINITY;
#pragma unroll
        for (int c = kGranA - 1; c >= 0; --c)
#pragma unroll
          for (int k = 0; k < 8; k += 4) {
            mxA = ptx::fmax3(mxA, __uint_as_float(sA[c][k]), __uint_as_float(sA[c][k + 1]));
            mxA2 = ptx::fmax3(mxA2, __uint_as_float(sA[c][k + 2]), __uint_as_float(sA[c][k + 3]));
          }
        mxA = fmaxf(mxA, mxA2);
        mxsA = mxA * p.scale_log2;
        constexpr int kEarly = kGranA - 2;   // block B's reads are issued once this many granules of A are packed
#pragma unroll
        for (int c = 0; c < kGranA; ++c) {
#pragma unroll
          for (int k = 0; k < 8; k += 2)
            d16[c][k >> 1] = pack_f16x2_f(fmaf(__uint_as_float(sA[c][k]), p.scale_log2, -mxsA),
                                          fmaf(__uint_as_float(sA[c][k + 1]), p.scale_log2, -mxsA));
          if (c == kEarly - 1) {
            // Block B's TMEM reads go out HERE, under the rest of block A's conversion: their address carries a real
            // (never taken) dependency on the last packed register, so ptxas can neither hoist them above the
            // conversion (104 live registers) nor sink them to their first use.
            ptx::mbar_wait(&s_full[2 + g], ph);   // second half of S (long complete by now)
            ptx::tc_fence_after();
            uint32_t bump;
            asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %1, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(bump) : "r"(d16[c][3]));
#pragma unroll
            for (int cb = 0; cb < kGranB; ++cb)
              if (cb < kGranB - 1 || own13) ptx::tmem_ld_x8(t_s + bump + (kGranA + cb) * 16, sB[cb]);
          }
        }
        PP_TS(2);
        ptx::tmem_ld_wait();
        // every read of S has landed in registers: the S columns go back to the QK^T issuer (next head)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s_free[g]);
        if (own13) {   // granule 24: keys >= N count as -inf
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k >= vlast) sB[kGranB - 1][k] = 0xff800000u;
        }
        float mxB2 = -INFINITY;
        mxB = -INFINITY;
#pragma unroll
        for (int c = kGranB - 1; c >= 0; --c) {
          if (c == kGranB - 1 && !own13) continue;
#pragma unroll
          for (int k = 0; k < 8; k += 4) {
            mxB = ptx::fmax3(mxB, __uint_as_float(sB[c][k]), __uint_as_float(sB[c][k + 1]));
            mxB2 = ptx::fmax3(mxB2, __uint_as_float(sB[c][k + 2]), __uint_as_float(sB[c][k + 3]));
          }
        }
        mxB = fmaxf(mxB, mxB2);
        mxsB = mxB * p.scale_log2;
#pragma unroll
        for (int c = 0; c < kGranB; ++c) {
          if (c == kGranB - 1 && !own13) break;
#pragma unroll
          for (int k = 0; k < 8; k += 2)
            d16[kGranA + c][k >> 1] = pack_f16x2_f(fmaf(__uint_as_float(sB[c][k]), p.scale_log2, -mxsB),
                                              fmaf(__uint_as_float(sB[c][k + 1]), p.scale_log2, -mxsB));
        }
      }
      PP_TS(3);

      // ---- e = exp2(d) on the packed halves (two MUFU results per register), block sums in fp32
      float psA0 = 0.f, psA1 = 0.f, psB0 = 0.f, psB1 = 0.f;
#pragma unroll 1
      for (int once = 0; once < p.one; ++once)
#pragma unroll
      for (int c = 0; c < kGran; ++c) {
        if (c == kGran - 1 && !own13) break;
        ex2_f16x2_x4_gated(d16[c], c >= 2 ? d16[c - 2][3] : 0u);   // at most two granules of half results in flight
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c < kGranA) add_f16x2_to_f32(psA0, psA1, d16[c][k]);
          else add_f16x2_to_f32(psB0, psB1, d16[c][k]);
        }
        if (want_cls) {   // warp-uniform; only lane 0 (query row 0) stores
          asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %5, 0;\n\t@p st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t}" ::"r"(
                           cls_dst + c * 32),
                       "r"(d16[c][0]), "r"(d16[c][1]), "r"(d16[c][2]), "r"(d16[c][3]), "r"(static_cast<uint32_t>(lane))
                       : "memory");
        }
      }
      PP_TS(4);

      // ---- the four blocks of a row (two per thread) reconcile: p = e f_b,
      //      f_b = exp2(m_b - M) / sum_u(sum_u exp2(m_u - M)), M = max_u m_u.  (Every block holds a valid key for
      //      193 <= N <= 200, so every m_b is finite.)
      //      Each thread first folds its own two blocks into (M_t, tot_t) -- under the wait for its partner -- so that
      //      one MUFU level (not two) follows the exchange.
      const float Mt = fmaxf(mxsA, mxsB);
      const float eA = ptx::ex2_approx(mxsA - Mt), eB = ptx::ex2_approx(mxsB - Mt);
      const float tot_t = (psA0 + psA1) * eA + (psB0 + psB1) * eB;
      //      (first barrier: the partner has read the previous head's values -- the exchange buffer is single)
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red_addr(c2)), "f"(Mt), "f"(tot_t) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      PP_TS(5);
      float invA, invB;
      {
        float2 a0, a1;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a0.x), "=f"(a0.y) : "r"(red_addr(0)) : "memory");
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a1.x), "=f"(a1.y) : "r"(red_addr(1)) : "memory");
        const float M = fmaxf(a0.x, a1.x);
        // the same expression in both threads of the row: the factors of a row's blocks share one denominator
        const float tot = a0.y * ptx::ex2_approx(a0.x - M) + a1.y * ptx::ex2_approx(a1.x - M);
        const float gr = ptx::ex2_approx(Mt - M) * ptx::rcp_approx(tot);
        invA = eA * gr;
        invB = eB * gr;
      }
      if (want_cls && lane == 0) {
        cls_factor[(g * 2 + c2) * 2 + 0] = invA;
        cls_factor[(g * 2 + c2) * 2 + 1] = invB;
        ptx::mbar_arrive(&cls_full[g]);   // release: stage + factors are visible to role 3 once both halves have arrived
      }
      const uint32_t invA16 = pack_f16x2_f(invA, invA), invB16 = pack_f16x2_f(invB, invB);
      PP_TS(6);

      // ---- the P tile is shared by both groups: the previous head's P V and Pbar MMAs must have retired
      if (h > 0) ptx::mbar_wait(&p_free[g ^ 1], ((h - 1) >> 1) & 1);
      PP_TS(7);
#pragma unroll
      for (int c = 0; c < kGran; ++c) {
        if (c == kGran - 1 && !own13) break;
        const uint32_t inv16 = c < kGranA ? invA16 : invB16;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pb(c & 3) + (c >> 2) * kPBlockBytes),
                     "r"(hmul2_u(d16[c][0], inv16)), "r"(hmul2_u(d16[c][1], inv16)), "r"(hmul2_u(d16[c][2], inv16)),
                     "r"(hmul2_u(d16[c][3], inv16))
                     : "memory");
      }
      PP_TS(8);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[g]);
      PP_TS(9);
    }

    // ---- head-averaged map rows -> HBM once per (image, query tile): Pbar holds the SUM over heads; the tile leaves
    //      through smem slabs (the Q / K / V stages are dead) and TMA stores, exactly as in attention_kernel.  Here the
    //      16 warps split the columns four ways again: cg = sub owns granules 4 c + cg.
    if (want_avg) {
      // Pbar += P of the LAST head (and its P V, which reads the stage area reused below) must have completed
      ptx::mbar_wait(&p_free[(nh - 1) & 1], ((nh - 1) >> 1) & 1);
      ptx::tc_fence_after();
      if (warp_has_rows) {
        const int cg = sub;
        const float inv_h = 1.0f / static_cast<float>(p.H);
        const uint32_t t_avg = lane_base + kTmemAvg + cg * 8;
        uint32_t a[kMaxGran][8];
#pragma unroll
        for (int c = 0; c < kMaxGran; ++c)
          if (c < kMaxGran - 1 || cg < 2) ptx::tmem_ld_x8(t_avg + c * 32, a[c]);
        ptx::tmem_ld_wait();
        const uint32_t slab0 = ptx::smem_u32(smem) + r * 128;
#pragma unroll
        for (int c = 0; c < kMaxGran; ++c) {
          if (c < kMaxGran - 1 || cg < 2) {
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
      }
      asm volatile("bar.sync 9, 512;" ::: "memory");
      if (warp == kCtrlWarps && ptx::elect_one()) {
        constexpr int nslabs = (KP + 31) >> 5;
        for (int s = 0; s < nslabs; ++s) {
          if (split_cta && p.part_images == 0) ptx::tma_reduce_add_3d(&tmap_avg, smem + s * (BM * 128), 32 * s, qt * BM, b);
          else ptx::tma_store_3d(&tmap_avg, smem + s * (BM * 128), 32 * s, qt * BM, avg_img + b);
        }
        ptx::tma_store_commit();
      }
      __syncwarp();
    }
    ptx::tma_store_wait_read<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vitb200
