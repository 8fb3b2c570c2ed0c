// Flash-style attention on tcgen05 for sm_100a (spatial self-attention with reference-feature
// K/V, CLIP cross-attention).  One CTA = one (image, head, 128-query tile); 192 threads:
//   warp 0   : TMA producer — Q once, then K / V^T tiles through a KST-stage ring
//   warp 1   : MMA issuer   — S_j = Q K_j^T (TMEM), O += P_j V_j (TMEM); S_{j+1} is issued as soon
//                             as the softmax warps have drained S_j into registers
//   warps 2-5: softmax      — one query row per thread: tcgen05.ld S_j, online softmax in the
//                             exp2 domain with a lazy rescale of O (only when the running max grows
//                             by more than 2^8, so P stays <= 256 in fp16), P_j -> shared memory in
//                             the 128B-swizzled K-major layout, final O / l epilogue
// All MMA operands are K-major 128B-swizzled tiles (same descriptors as the GEMM): Q [128 x d],
// K [BKV x d] loaded through a 4-D tensor map [d, heads, L, image] whose out-of-bounds fill zero-pads
// d to a multiple of 16 and L to the tile; V is consumed as V^T [d x BKV] (emitted transposed by the
// K/V projection GEMM), so P V needs no MN-major operand.
//
// Algorithmic FLOPs per launch: 4 * nimg * heads * lq * lkv * d.
#include <stdlib.h>

#include <type_traits>

#include "host_common.h"
#include "ptx.cuh"
#include "attn_common.cuh"
#include "../../include/mdk.h"

namespace mdk {

template <int NCH, int BKV, int KST>
struct AttnCfg {
  static constexpr int Q_BYTES = NCH * ATT_BQ * 128;
  static constexpr int K_STAGE = NCH * BKV * 128;
  static constexpr int V_CHUNK = NCH * 64 * 128;  // up to 64*NCH rows of 128 B per 64-wide kv chunk
  static constexpr int V_STAGE = (BKV / 64) * V_CHUNK;
  static constexpr int P_BYTES = (BKV / 64) * ATT_BQ * 128;
  // dynamic shared memory is declared __align__(1024) (checked at run time), so no alignment slack:
  // at head_dim <= 64 two CTAs must fit in one SM's 228 KB
  static constexpr int SMEM_BYTES = Q_BYTES + KST * (K_STAGE + V_STAGE) + P_BYTES + 256;
  // S (128 x BKV fp32, single buffer: it is drained into registers at the start of each softmax
  // step) at column 0, O (128 x 64*NCH fp32) at column 128
  static constexpr uint32_t S_COL = 0, O_COL = BKV;
  static constexpr uint32_t TMEM_COLS = (BKV + 64 * NCH <= 128) ? 128 : (BKV + 64 * NCH <= 256) ? 256 : 512;
  // head_dim <= 64: 3 CTAs/SM with 64-key tiles (64 KB smem, 128 TMEM columns, <= 112 registers) or
  // 2 CTAs/SM with 128-key tiles.  The softmax warps are latency-bound (ncu: MUFU 62 %, issue 45 % at
  // 2 CTAs/SM), so more independent CTAs per SM is what raises the MUFU utilisation.
  static constexpr int CTAS_PER_SM = (NCH == 1) ? (BKV == 64 ? 3 : 2) : ((NCH == 2 && BKV == 64) ? 2 : 1);
};

// ONES: row d of every head of V^T holds ones (d % 16 == 8), so column d of O = P V accumulates the
// softmax row sums of the fp16-rounded P on the tensor core; the softmax warps then neither add up the
// exponentials nor rescale a running sum (one FADD per score less on the latency-bound softmax path).
// ONES == 2 additionally takes the exponentials of tile j against the reference maximum known BEFORE
// tile j (the running maximum of tiles < j): the row maximum of tile j is then computed inside the
// exponential loop (ALU pipe, interleaved with the MUFU work) instead of in front of it, which removes
// the 128-score max pass from the stretch in which the MUFU pipe idles.  The reference is updated (and O
// rescaled) at the start of tile j+1 when tile j's maximum exceeded it by more than 2^8; if a row's
// maximum exceeds the stale reference by more than 2^14 (fp16 P would overflow) the tile is redone
// against the new maximum (rare: the running maximum settles after the first tiles).
// SPLIT: separate rings (full / empty barriers) for the K tiles and the V^T tiles of the same KST stages.
// With one ring a stage is released by the completion of P_{j-1} V_{j-1} and the load of K_{j+1} | V_{j+1}
// into it must land before the MMA warp (which waits for K_{j+1} ahead of P_j V_j) can move on: one TMA
// round trip per tile sits on the critical path whenever it is longer than the exponential phase.  A K
// stage is really free as soon as S = Q K^T has been computed from it — a whole tile period earlier — so
// with split rings K_{j+1} is requested right after Q K_{j-1}^T and V_j right after P_{j-2} V_{j-2}: both
// loads get about two tile periods to arrive, with the same shared-memory footprint.
// TRACE (diagnostic instantiation, MDK_ATTN_TRACE=1 + mdk_attn_debug_trace): the CTA in the middle of the
// grid's x range (head 0, image 0) records clock64() at the hand-off points of every tile:
//   softmax warp 2, lane 0:  0 S_j full   1 S_j in registers   2 P_{j-1}V_{j-1} done   3 exponentials done   4 P_j published
//   MMA warp:                5 K_{j+1} landed   6 S_j drained   7 QK_{j+1}^T issued   8 P_j full (+ V_j landed)   9 P_j V_j issued
//   TMA warp:               10 stage for tile j free   11 loads of tile j issued     (split rings: 10/11 = V^T, 12/13 = K)
template <int NCH, int BKV, int KST, int ONES, bool SPLIT = false, bool TRACE = false>
__global__ void __launch_bounds__(ATT_THREADS, (NCH == 1) ? (BKV == 64 ? 3 : 2) : ((NCH == 2 && BKV == 64) ? 2 : 1))
attn_tc_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<NCH, BKV, KST>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mdk attn: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + KST * Cfg::K_STAGE;
  uint8_t* sP = sV + KST * Cfg::V_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_bar = bars;                    // [1]
  uint64_t* kv_full = bars + 1;              // [KST]
  uint64_t* kv_empty = bars + 1 + KST;       // [KST]
  uint64_t* s_full = bars + 1 + 2 * KST;     // [1] S_j complete in TMEM
  uint64_t* s_free = bars + 2 + 2 * KST;     // [1] S_j drained into registers (4 warp arrivals)
  uint64_t* p_full = bars + 3 + 2 * KST;     // [1]
  uint64_t* pv_done = bars + 4 + 2 * KST;    // [1]
  // SPLIT: kv_full / kv_empty serve the K tiles, v_full / v_empty the V^T tiles
  uint64_t* v_full = bars + 5 + 2 * KST;     // [KST]
  uint64_t* v_empty = bars + 5 + 3 * KST;    // [KST]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 5 + 4 * KST);
  static_assert((5 + 4 * KST) * 8 + 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int kvimg = img / p.kv_div;
  const int n_tiles = p.n_kv_tiles;
  const bool tr_on = TRACE && p.trace != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
  auto tr = [&](int tile, int slot) {
    if constexpr (TRACE) {
      if (tr_on && tile < p.trace_cap) p.trace[tile * 16 + slot] = clock64();
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_bar, 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      if constexpr (SPLIT) {
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);  // one arrive per softmax warp
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {   // single-thread region known to the compiler: no ELECT/R2UR loop per TMA
      mbar_expect_tx(q_bar, Cfg::Q_BYTES);
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        tma_load_4d(sQ + c * ATT_BQ * 128, &p.tmQ, q_bar, c * 64, head, q0, img);
      const uint32_t stage_bytes =
          static_cast<uint32_t>(Cfg::K_STAGE + (BKV / 64) * p.dn * 128);
      if constexpr (SPLIT) {
        const uint32_t v_bytes = static_cast<uint32_t>((BKV / 64) * p.dn * 128);
        auto load_k = [&](int t) {
          const int st = t % KST;
          mbar_wait(&kv_empty[st], static_cast<uint32_t>((t / KST) & 1) ^ 1u);   // Q K_{t-KST}^T has retired
          tr(t, 12);
          mbar_expect_tx(&kv_full[st], Cfg::K_STAGE);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
            tma_load_4d(sK + st * Cfg::K_STAGE + c * BKV * 128, &p.tmK, &kv_full[st], c * 64, head,
                        t * BKV, kvimg);
          tr(t, 13);
        };
        load_k(0);
        for (int j = 0; j < n_tiles; ++j) {
          if (j + 1 < n_tiles) load_k(j + 1);     // K runs one tile ahead of V^T
          const int st = j % KST;
          mbar_wait(&v_empty[st], static_cast<uint32_t>((j / KST) & 1) ^ 1u);    // P_{j-KST} V_{j-KST} has retired
          tr(j, 10);
          mbar_expect_tx(&v_full[st], v_bytes);
#pragma unroll
          for (int c = 0; c < BKV / 64; ++c)
            tma_load_3d(sV + st * Cfg::V_STAGE + c * Cfg::V_CHUNK, &p.tmV, &v_full[st], j * BKV + c * 64,
                        head * p.vt_head_rows, kvimg);
          tr(j, 11);
        }
      } else {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        tr(j, 10);
        mbar_expect_tx(&kv_full[stage], stage_bytes);
        const int kv0 = j * BKV;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          tma_load_4d(sK + stage * Cfg::K_STAGE + c * BKV * 128, &p.tmK, &kv_full[stage], c * 64,
                      head, kv0, kvimg);
#pragma unroll
        for (int c = 0; c < BKV / 64; ++c)
          tma_load_3d(sV + stage * Cfg::V_STAGE + c * Cfg::V_CHUNK, &p.tmV, &kv_full[stage],
                      kv0 + c * 64, head * p.vt_head_rows, kvimg);
        tr(j, 11);
        if (++stage == KST) {
          stage = 0;
          phase ^= 1u;
        }
      }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc_s = make_idesc_f16(ATT_BQ, BKV);
    const uint32_t idesc_o = make_idesc_f16(ATT_BQ, static_cast<uint32_t>(p.dn));
    const uint32_t tS = tmem_base + Cfg::S_COL;
    const uint32_t tO = tmem_base + Cfg::O_COL;
    auto issue_s = [&](int stage) {
      if (elect_one()) {
        for (int ks = 0; ks < p.dk16; ++ks) {
          const int c = ks >> 2, w = ks & 3;
          const uint64_t adesc = make_sdesc_sw128(smem_u32(sQ + c * ATT_BQ * 128)) + 2u * w;
          const uint64_t bdesc =
              make_sdesc_sw128(smem_u32(sK + stage * Cfg::K_STAGE + c * BKV * 128)) + 2u * w;
          tc_mma_f16_ss(tS, adesc, bdesc, idesc_s, ks > 0 ? 1u : 0u);
        }
        tc_commit(s_full);
        if constexpr (SPLIT) tc_commit(&kv_empty[stage]);   // the K stage is free once S has been computed
      }
      __syncwarp();
    };
    mbar_wait(q_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      // look ahead: S_{j+1} as soon as the softmax warps hold S_j in registers
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == KST) {
        nstage = 0;
        nphase ^= 1u;
      }
      if (j + 1 < n_tiles) {
        mbar_wait(&kv_full[nstage], nphase);
        if (lane == 0) tr(j, 5);
        mbar_wait(s_free, static_cast<uint32_t>(j & 1));
        if (lane == 0) tr(j, 6);
        tc_fence_after();
        issue_s(nstage);
        if (lane == 0) tr(j, 7);
      }
      mbar_wait(p_full, static_cast<uint32_t>(j & 1));
      if constexpr (SPLIT) mbar_wait(&v_full[stage], phase);
      if (lane == 0) tr(j, 8);
      tc_fence_after();
      if (elect_one()) {
        const int kv = p.lkv - j * BKV;                      // keys in this tile
        const int ksteps = (kv >= BKV) ? (BKV / 16) : ((kv + 31) >> 5) * 2;   // whole 32-column chunks of P
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          if (ks >= ksteps) break;
          const int c = ks >> 2, w = ks & 3;
          const uint64_t adesc = make_sdesc_sw128(smem_u32(sP + c * ATT_BQ * 128)) + 2u * w;
          const uint64_t bdesc =
              make_sdesc_sw128(smem_u32(sV + stage * Cfg::V_STAGE + c * Cfg::V_CHUNK)) + 2u * w;
          tc_mma_f16_ss(tO, adesc, bdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        if constexpr (SPLIT)
          tc_commit(&v_empty[stage]);
        else
          tc_commit(&kv_empty[stage]);
        tc_commit(pv_done);
      }
      __syncwarp();
      if (lane == 0) tr(j, 9);
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ======================= softmax / correction / epilogue warps =======================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + Cfg::S_COL;
    const uint32_t tO = tmem_base + lane_off + Cfg::O_COL;
    float m_used = -INFINITY;  // reference max (scaled, log2 domain) the exponentials are taken against
    float l_sum = 0.f;
    float pend_alpha = 1.0f;    // ONES == 2: reference update decided at the end of the previous tile
    bool pend_rescale = false;
    const uint32_t prow = smem_u32(sP) + static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const bool poly = p.poly != 0;

    const bool tr_sm = (warp == 2 && lane == 0);
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, static_cast<uint32_t>(j & 1));
      if (tr_sm) tr(j, 0);
      tc_fence_after();
      // 32-column chunks of this tile that hold real keys (the last tile of a ragged sequence, e.g. the
      // 257 CLIP tokens, may need only one): the others are neither loaded, exponentiated nor fed to P V
      const int nvalid = p.lkv - j * BKV;  // columns >= nvalid are padding (last tile only)
      const int nch = (nvalid >= BKV) ? (BKV / 32) : ((nvalid + 31) >> 5);
      uint32_t v[BKV / 32][32];
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c)
        if (c < nch) tmem_ld_x32(tS + c * 32, v[c]);
      tmem_wait_ld();
      if (tr_sm) tr(j, 1);
      // S_j now lives in registers: the tensor core may overwrite it with S_{j+1}
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      constexpr bool STALE = (ONES == 2);
      const bool fresh = !STALE || j == 0;   // the tile's own maximum is the reference (always for tile 0)
      float mx = -INFINITY;
      if (nvalid < BKV) {
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (c < nch && c * 32 + e >= nvalid) v[c][e] = 0xff800000u;  // -inf
          }
        }
      }
      float alpha = 1.0f;
      bool rescale = false;
      if (fresh) {
        // 4 independent running maxima: a single fmaxf chain over 128 scores is 128 x 4 cycles of
        // pure dependency latency per tile
        float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          if (c < nch) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mp[e & 3] = fmaxf(mp[e & 3], __uint_as_float(v[c][e]));
          }
        }
        mx = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
        mx *= p.scale_log2;
        if (j == 0) {
          m_used = mx;
        } else if (mx > m_used + ATT_RESCALE_THRESHOLD) {
          alpha = ex2_approx(m_used - mx);
          m_used = mx;
          if constexpr (!ONES) l_sum *= alpha;
          rescale = true;
        }
      } else {
        // stale reference: the update decided at the end of the previous tile takes effect now
        alpha = pend_alpha;
        rescale = pend_rescale;
      }
      // P_{j-1} V_{j-1} must have retired before P (single buffer) or O may be touched
      if (j > 0) {
        mbar_wait(pv_done, static_cast<uint32_t>((j - 1) & 1));
        tc_fence_after();
      }
      if (tr_sm) tr(j, 2);
      if (__any_sync(0xffffffffu, rescale)) {
        for (int c = 0; c < p.dn; c += 16) {
          uint32_t o[16];
          tmem_ld_x16(tO + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tmem_st_x16(tO + c, o);
        }
        tmem_wait_st();
      }
      // exp2, row sum, fp16 pack and the store of P, 8 columns (one 16-byte piece) at a time.
      // P -> smem, K-major, 128B swizzle: 16-byte piece q of row r lands at piece (q ^ (r & 7))
      float rsp[2] = {0.f, 0.f};  // independent partial row sums (short dependency chains)
      float tmx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // ONES == 2: this tile's maximum, for tile j+1
      auto exp_tile = [&](auto poly_tag) {
        constexpr bool POLY = decltype(poly_tag)::value;
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          if (c >= nch) break;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (STALE) {
                tmx[e] = fmaxf(tmx[e], fmaxf(__uint_as_float(v[c][q4 * 8 + 2 * e]),
                                             __uint_as_float(v[c][q4 * 8 + 2 * e + 1])));
              }
              const float x0 = fmaf(__uint_as_float(v[c][q4 * 8 + 2 * e]), p.scale_log2, -m_used);
              const float x1 = fmaf(__uint_as_float(v[c][q4 * 8 + 2 * e + 1]), p.scale_log2, -m_used);
              if constexpr (POLY) {
                if (e & 1) {
                  pk[e] = ex2_poly_h2(x0, x1);   // FMA pipe, half2 (row sums come from the ones row of V^T)
                  continue;
                }
              }
              const float p0 = ex2_approx(x0);
              const float p1 = ex2_approx(x1);
              if constexpr (!ONES) rsp[e & 1] += p0 + p1;
              pk[e] = pack_half2(p0, p1);
            }
            const uint32_t col8 = c * 4 + q4;   // 8-column piece index inside the BKV tile
            const uint32_t cc = col8 >> 3, q = col8 & 7u;
            st_shared_v4(prow + cc * (ATT_BQ * 128) + ((q ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
        }
      };
      if (ONES && poly)
        exp_tile(std::true_type{});
      else
        exp_tile(std::false_type{});
      if constexpr (STALE) {
        pend_alpha = 1.0f;
        pend_rescale = false;
        if (!fresh) {
          const float mxs = fmaxf(fmaxf(tmx[0], tmx[1]), fmaxf(tmx[2], tmx[3])) * p.scale_log2;
          const bool over = mxs > m_used + 14.0f;   // exp2 of more than 14 would leave fp16's range in P
          if (__any_sync(0xffffffffu, over)) {
            // redo this tile against its own maximum (P has not been published yet; P_{j-1} V_{j-1} has
            // retired: pv_done was awaited above), rows below the limit keep their reference
            const float a2 = over ? ex2_approx(m_used - mxs) : 1.0f;
            if (over) m_used = mxs;
            for (int c = 0; c < p.dn; c += 16) {
              uint32_t o[16];
              tmem_ld_x16(tO + c, o);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * a2);
              tmem_st_x16(tO + c, o);
            }
            tmem_wait_st();
            if (poly)
              exp_tile(std::true_type{});
            else
              exp_tile(std::false_type{});
          } else if (mxs > m_used + ATT_RESCALE_THRESHOLD) {
            pend_alpha = ex2_approx(m_used - mxs);   // applied to O at the start of tile j+1
            m_used = mxs;
            pend_rescale = true;
          }
        }
      }
      if constexpr (!ONES) l_sum += rsp[0] + rsp[1];
      if (tr_sm) tr(j, 3);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tr_sm) tr(j, 4);
    }
    // ---- epilogue: O / l ----
    mbar_wait(pv_done, static_cast<uint32_t>((n_tiles - 1) & 1));
    tc_fence_after();
    float inv;
    if constexpr (ONES) {
      uint32_t o[16];
      tmem_ld_x16(tO + static_cast<uint32_t>(p.d & ~15), o);   // d % 16 == 8: the sums sit in column 8
      tmem_wait_ld();
      inv = 1.0f / __uint_as_float(o[8]);
    } else {
      inv = 1.0f / l_sum;
    }
    const int qrow = q0 + row;
    __half* dst = p.out + (static_cast<long long>(blockIdx.z) * p.lq + qrow) * p.ldo + head * p.d;
    for (int c = 0; c < p.dn; c += 16) {
      uint32_t o[16];
      tmem_ld_x16(tO + c, o);
      tmem_wait_ld();
      if (qrow < p.lq) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (c + q * 8 < p.d) {
            uint4 val;
            val.x = pack_half2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            val.y = pack_half2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            val.z = pack_half2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            val.w = pack_half2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c + q * 8) = val;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

static long long* g_attn_trace = nullptr;   // set by mdk_attn_debug_trace (diagnostics only)
static int g_attn_trace_cap = 0;

template <int NCH, int BKV, int KST, int ONES, bool SPLIT = false, bool TRACE = false>
static int launch_attn(const mdk_ctx* ctx, AttnParams& p, const mdk_attn_args* a,
                       cudaStream_t stream) {
  using Cfg = AttnCfg<NCH, BKV, KST>;
  static unsigned long long attr_set_mask = 0;   // bit d: attribute set on device d (it is a per-device property)
  const bool attr_set = ((attr_set_mask >> (ctx->device & 63)) & 1ull) != 0;
  p.trace = TRACE ? g_attn_trace : nullptr;
  p.trace_cap = TRACE ? g_attn_trace_cap : 0;
  if (!attr_set) {
    MDK_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<NCH, BKV, KST, ONES, SPLIT, TRACE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set_mask |= 1ull << (ctx->device & 63);
  }
  if (encode_attn_maps(p, a, BKV)) return -1;
  p.n_kv_tiles = (a->lkv + BKV - 1) / BKV;
  dim3 grid((a->lq + ATT_BQ - 1) / ATT_BQ, a->heads, a->nimg);
  MDK_CHECK_CUDA(launch_pdl(attn_tc_kernel<NCH, BKV, KST, ONES, SPLIT, TRACE>, grid, dim3(ATT_THREADS), Cfg::SMEM_BYTES,
                            stream, p));
  count_launch();
  (void)ctx;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Ping-pong variant: one CTA = one (image, head, 256-query block) = two 128-row query tiles, each with
// its own softmax warp group, S/O accumulators in TMEM and P buffer, sharing every K / V^T tile.
// Why: with two independent 128-row CTAs per SM the softmax warps of both CTAs phase-lock (ncu source
// counters: both sit on the MUFU.EX2 instructions at half rate, then both wait on the tensor core /
// TMEM with the MUFU pipe idle: 62 % MUFU utilisation at d = 40 where the exponentials are the
// bound).  Here a baton (two mbarriers) lets only ONE group at a time run its exp2 loop; the other
// group meanwhile drains its next S tile, takes the row maxima, waits for its P V MMA and rescales —
// so the MUFU pipe always has one warp per SM sub-partition feeding it.
//   warp 0: TMA (Q block once, K / V^T ring of KST stages)   warp 1: MMA issuer
//   warps 2-5: softmax group 0 (rows 0..127)                 warps 6-9: softmax group 1 (rows 128..255)
// MMA order per K/V tile j: P_0 V_j, P_1 V_j, release stage j, then S_0 / S_1 of tile j+2 as soon as
// the groups have drained tile j+1 (S is single-buffered per group, look-ahead of two tiles needs
// KST >= 3).
constexpr int ATT_PP_THREADS = 320;

template <int NCH, int BKV, int KST>
struct AttnPPCfg {
  static constexpr int Q_TILE = NCH * ATT_BQ * 128;
  static constexpr int K_STAGE = NCH * BKV * 128;
  static constexpr int V_CHUNK = NCH * 64 * 128;
  static constexpr int V_STAGE = (BKV / 64) * V_CHUNK;
  static constexpr int P_TILE = (BKV / 64) * ATT_BQ * 128;
  static constexpr int SMEM_BYTES = 2 * Q_TILE + KST * (K_STAGE + V_STAGE) + 2 * P_TILE + 256;
  static constexpr uint32_t O_COL0 = 2 * BKV;        // S_g at column g * BKV, O_g at O_COL0 + g * 64 * NCH
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(2 * BKV + 2 * 64 * NCH <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int NCH, int BKV, int KST, bool ONES>
__global__ void __launch_bounds__(ATT_PP_THREADS, 1)
attn_pp_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = AttnPPCfg<NCH, BKV, KST>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mdk attn: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * Cfg::Q_TILE;
  uint8_t* sV = sK + KST * Cfg::K_STAGE;
  uint8_t* sP = sV + KST * Cfg::V_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::P_TILE);
  uint64_t* q_bar = bars;                     // [1]
  uint64_t* kv_full = bars + 1;               // [KST]
  uint64_t* kv_empty = bars + 1 + KST;        // [KST]
  uint64_t* s_full = bars + 1 + 2 * KST;      // [2] S_g(j) complete in TMEM
  uint64_t* s_free = s_full + 2;              // [2] S_g(j) drained into registers (4 warp arrivals)
  uint64_t* p_full = s_full + 4;              // [2] P_g(j) in shared memory (4 warp arrivals)
  uint64_t* pv_done = s_full + 6;             // [2] O_g += P_g(j) V_j retired
  uint64_t* baton = s_full + 8;               // [2] group g may run its exp2 loop (4 warp arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * ATT_BQ);
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int kvimg = img / p.kv_div;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_bar, 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 4);
      mbar_init(&p_full[g], 4);
      mbar_init(&pv_done[g], 1);
      mbar_init(&baton[g], 4);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {   // single-thread region known to the compiler: no ELECT/R2UR loop per TMA
      mbar_expect_tx(q_bar, 2 * Cfg::Q_TILE);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          tma_load_4d(sQ + t * Cfg::Q_TILE + c * ATT_BQ * 128, &p.tmQ, q_bar, c * 64, head,
                      q0 + t * ATT_BQ, img);
      const uint32_t stage_bytes =
          static_cast<uint32_t>(Cfg::K_STAGE + (BKV / 64) * p.dn * 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        mbar_expect_tx(&kv_full[stage], stage_bytes);
        const int kv0 = j * BKV;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          tma_load_4d(sK + stage * Cfg::K_STAGE + c * BKV * 128, &p.tmK, &kv_full[stage], c * 64,
                      head, kv0, kvimg);
#pragma unroll
        for (int c = 0; c < BKV / 64; ++c)
          tma_load_3d(sV + stage * Cfg::V_STAGE + c * Cfg::V_CHUNK, &p.tmV, &kv_full[stage],
                      kv0 + c * 64, head * p.vt_head_rows, kvimg);
        if (++stage == KST) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc_s = make_idesc_f16(ATT_BQ, BKV);
    const uint32_t idesc_o = make_idesc_f16(ATT_BQ, static_cast<uint32_t>(p.dn));
    auto issue_s = [&](int g, int stage) {
      if (elect_one()) {
        const uint32_t tS = tmem_base + static_cast<uint32_t>(g * BKV);
        for (int ks = 0; ks < p.dk16; ++ks) {
          const int c = ks >> 2, w = ks & 3;
          const uint64_t adesc =
              make_sdesc_sw128(smem_u32(sQ + g * Cfg::Q_TILE + c * ATT_BQ * 128)) + 2u * w;
          const uint64_t bdesc =
              make_sdesc_sw128(smem_u32(sK + stage * Cfg::K_STAGE + c * BKV * 128)) + 2u * w;
          tc_mma_f16_ss(tS, adesc, bdesc, idesc_s, ks > 0 ? 1u : 0u);
        }
        tc_commit(&s_full[g]);
      }
      __syncwarp();
    };
    mbar_wait(q_bar, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    issue_s(1, 0);
    if (n_tiles > 1) {
      mbar_wait(&kv_full[1 % KST], 0);
      for (int g = 0; g < 2; ++g) {
        mbar_wait(&s_free[g], 0);
        tc_fence_after();
        issue_s(g, 1 % KST);
      }
    }
    int stage = 0;                     // stage of tile j
    int stage2 = 2 % KST;              // stage of tile j + 2
    uint32_t phase2 = (2 / KST) & 1u;  // kv_full parity of tile j + 2
    for (int j = 0; j < n_tiles; ++j) {
      for (int g = 0; g < 2; ++g) {
        mbar_wait(&p_full[g], static_cast<uint32_t>(j & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tO = tmem_base + Cfg::O_COL0 + static_cast<uint32_t>(g * 64 * NCH);
          const int kv = p.lkv - j * BKV;
          const int ksteps = (kv >= BKV) ? (BKV / 16) : ((kv + 31) >> 5) * 2;
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks) {
            if (ks >= ksteps) break;
            const int c = ks >> 2, w = ks & 3;
            const uint64_t adesc =
                make_sdesc_sw128(smem_u32(sP + g * Cfg::P_TILE + c * ATT_BQ * 128)) + 2u * w;
            const uint64_t bdesc =
                make_sdesc_sw128(smem_u32(sV + stage * Cfg::V_STAGE + c * Cfg::V_CHUNK)) + 2u * w;
            tc_mma_f16_ss(tO, adesc, bdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(&pv_done[g]);
          if (g == 1) tc_commit(&kv_empty[stage]);   // both groups are through with K_j / V_j
        }
        __syncwarp();
      }
      if (j + 2 < n_tiles) {
        mbar_wait(&kv_full[stage2], phase2);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&s_free[g], static_cast<uint32_t>((j + 1) & 1));
          tc_fence_after();
          issue_s(g, stage2);
        }
      }
      if (++stage == KST) stage = 0;
      if (++stage2 == KST) {
        stage2 = 0;
        phase2 ^= 1u;
      }
    }
  } else {
    // ======================= softmax groups =======================
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // query row inside the group's tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + static_cast<uint32_t>(g * BKV);
    const uint32_t tO = tmem_base + lane_off + Cfg::O_COL0 + static_cast<uint32_t>(g * 64 * NCH);
    float m_used = -INFINITY;
    float l_sum = 0.f;
    const uint32_t prow = smem_u32(sP + g * Cfg::P_TILE) + static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    if (g == 1) {   // group 0 runs its first exp2 phase without waiting
      __syncwarp();
      if (lane == 0) mbar_arrive(&baton[0]);
    }

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[g], static_cast<uint32_t>(j & 1));
      tc_fence_after();
      // 32-column chunks of this tile that hold real keys (the last tile of a ragged sequence, e.g. the
      // 257 CLIP tokens, may need only one): the others are neither loaded, exponentiated nor fed to P V
      const int nvalid = p.lkv - j * BKV;  // columns >= nvalid are padding (last tile only)
      const int nch = (nvalid >= BKV) ? (BKV / 32) : ((nvalid + 31) >> 5);
      uint32_t v[BKV / 32][32];
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c)
        if (c < nch) tmem_ld_x32(tS + c * 32, v[c]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[g]);
      if (nvalid < BKV) {
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (c < nch && c * 32 + e >= nvalid) v[c][e] = 0xff800000u;  // -inf
          }
        }
      }
      float mx;
      {
        float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          if (c < nch) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mp[e & 3] = fmaxf(mp[e & 3], __uint_as_float(v[c][e]));
          }
        }
        mx = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
      }
      mx *= p.scale_log2;
      float alpha = 1.0f;
      bool rescale = false;
      if (j == 0) {
        m_used = mx;
      } else if (mx > m_used + ATT_RESCALE_THRESHOLD) {
        alpha = ex2_approx(m_used - mx);
        m_used = mx;
        if constexpr (!ONES) l_sum *= alpha;
        rescale = true;
      }
      if (j > 0) {
        mbar_wait(&pv_done[g], static_cast<uint32_t>((j - 1) & 1));
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, rescale)) {
        for (int c = 0; c < p.dn; c += 16) {
          uint32_t o[16];
          tmem_ld_x16(tO + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tmem_st_x16(tO + c, o);
        }
        tmem_wait_st();
      }
      // ---- exclusive exp2 phase: only one group at a time feeds the MUFU pipe ----
      mbar_wait(&baton[g], static_cast<uint32_t>(j & 1));
      float rsp[2] = {0.f, 0.f};
#pragma unroll
      for (int c = 0; c < BKV / 32; ++c) {
        if (c >= nch) break;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 =
                ex2_approx(fmaf(__uint_as_float(v[c][q4 * 8 + 2 * e]), p.scale_log2, -m_used));
            const float p1 =
                ex2_approx(fmaf(__uint_as_float(v[c][q4 * 8 + 2 * e + 1]), p.scale_log2, -m_used));
            if constexpr (!ONES) rsp[e & 1] += p0 + p1;
            pk[e] = pack_half2(p0, p1);
          }
          const uint32_t col8 = c * 4 + q4;
          const uint32_t cc = col8 >> 3, q = col8 & 7u;
          st_shared_v4(prow + cc * (ATT_BQ * 128) + ((q ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&baton[g ^ 1]);
      if constexpr (!ONES) l_sum += rsp[0] + rsp[1];
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }
    // ---- epilogue: O / l ----
    mbar_wait(&pv_done[g], static_cast<uint32_t>((n_tiles - 1) & 1));
    tc_fence_after();
    float inv;
    if constexpr (ONES) {
      uint32_t o[16];
      tmem_ld_x16(tO + static_cast<uint32_t>(p.d & ~15), o);
      tmem_wait_ld();
      inv = 1.0f / __uint_as_float(o[8]);
    } else {
      inv = 1.0f / l_sum;
    }
    const int qrow = q0 + g * ATT_BQ + row;
    __half* dst = p.out + (static_cast<long long>(blockIdx.z) * p.lq + qrow) * p.ldo + head * p.d;
    for (int c = 0; c < p.dn; c += 16) {
      uint32_t o[16];
      tmem_ld_x16(tO + c, o);
      tmem_wait_ld();
      if (qrow < p.lq) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (c + q * 8 < p.d) {
            uint4 val;
            val.x = pack_half2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            val.y = pack_half2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            val.z = pack_half2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            val.w = pack_half2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c + q * 8) = val;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int NCH, int BKV, int KST, bool ONES>
static int launch_attn_pp(const mdk_ctx* ctx, AttnParams& p, const mdk_attn_args* a,
                          cudaStream_t stream) {
  using Cfg = AttnPPCfg<NCH, BKV, KST>;
  static unsigned long long attr_set_mask = 0;   // bit d: attribute set on device d (it is a per-device property)
  const bool attr_set = ((attr_set_mask >> (ctx->device & 63)) & 1ull) != 0;
  if (!attr_set) {
    MDK_CHECK_CUDA(cudaFuncSetAttribute(attn_pp_kernel<NCH, BKV, KST, ONES>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set_mask |= 1ull << (ctx->device & 63);
  }
  if (encode_attn_maps(p, a, BKV)) return -1;
  p.n_kv_tiles = (a->lkv + BKV - 1) / BKV;
  dim3 grid((a->lq + 2 * ATT_BQ - 1) / (2 * ATT_BQ), a->heads, a->nimg);
  MDK_CHECK_CUDA(launch_pdl(attn_pp_kernel<NCH, BKV, KST, ONES>, grid, dim3(ATT_PP_THREADS), Cfg::SMEM_BYTES, stream, p));
  count_launch();
  (void)ctx;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Split-key variant for head_dim <= 64 (the MUFU-bound L0 self-attention): the 128 keys of every S
// tile are handled as two independent 64-key online-softmax streams by two softmax warp groups
// (warps 2-5: keys 0..63, warps 6-9: keys 64..127 of each tile), each with its own running max, its
// own lazy rescale, its own P half and its own accumulator O_h in TMEM (S 128 + O_0 64 + O_1 64 = 256
// columns).  Nothing is exchanged per tile; the two partial results are merged once at the end:
//     O = (2^(m0-m) O_0 + 2^(m1-m) O_1) / (2^(m0-m) l_0 + 2^(m1-m) l_1),  m = max(m0, m1).
// Why: the exponentials bound this kernel (16 MUFU/clk/SM) and the softmax warps run them in a
// dependent tmem-load -> max -> exp2 -> store chain; with 8 softmax warps per CTA and 2 CTAs per SM
// there are 4 such chains per SM sub-partition instead of 2 to keep the MUFU pipe fed.
// Needs lkv >= 128 (both streams see real keys in tile 0, so their running maxima are finite).
constexpr int ATT_SK_THREADS = 320;
constexpr int SK_BKV = 128;

template <int KST, bool ONES>
__global__ void __launch_bounds__(ATT_SK_THREADS, 2)
attn_sk_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<1, SK_BKV, KST>;
  constexpr int BKV = SK_BKV;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mdk attn: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + KST * Cfg::K_STAGE;
  uint8_t* sP = sV + KST * Cfg::V_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_bar = bars;                    // [1]
  uint64_t* kv_full = bars + 1;              // [KST]
  uint64_t* kv_empty = bars + 1 + KST;       // [KST]
  uint64_t* s_full = bars + 1 + 2 * KST;     // [1] S_j complete in TMEM
  uint64_t* s_free = s_full + 1;             // [1] S_j drained into registers (8 warp arrivals)
  uint64_t* p_full = s_full + 2;             // [2] P half h of tile j in shared memory (4 warp arrivals)
  uint64_t* pv_done = s_full + 4;            // [2] O_h += P_h V_h retired
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int kvimg = img / p.kv_div;
  const int n_tiles = p.n_kv_tiles;
  constexpr uint32_t S_COL = 0, O_COL = BKV;   // O_h at O_COL + 64 h

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_bar, 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);
    for (int h = 0; h < 2; ++h) {
      mbar_init(&p_full[h], 4);
      mbar_init(&pv_done[h], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      mbar_expect_tx(q_bar, Cfg::Q_BYTES);
      tma_load_4d(sQ, &p.tmQ, q_bar, 0, head, q0, img);
      const uint32_t stage_bytes = static_cast<uint32_t>(Cfg::K_STAGE + (BKV / 64) * p.dn * 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        mbar_expect_tx(&kv_full[stage], stage_bytes);
        const int kv0 = j * BKV;
        tma_load_4d(sK + stage * Cfg::K_STAGE, &p.tmK, &kv_full[stage], 0, head, kv0, kvimg);
#pragma unroll
        for (int c = 0; c < BKV / 64; ++c)
          tma_load_3d(sV + stage * Cfg::V_STAGE + c * Cfg::V_CHUNK, &p.tmV, &kv_full[stage],
                      kv0 + c * 64, head * p.vt_head_rows, kvimg);
        if (++stage == KST) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc_s = make_idesc_f16(ATT_BQ, BKV);
    const uint32_t idesc_o = make_idesc_f16(ATT_BQ, static_cast<uint32_t>(p.dn));
    const uint32_t tS = tmem_base + S_COL;
    auto issue_s = [&](int stage) {
      if (elect_one()) {
        for (int ks = 0; ks < p.dk16; ++ks) {
          const uint64_t adesc = make_sdesc_sw128(smem_u32(sQ)) + 2u * ks;
          const uint64_t bdesc = make_sdesc_sw128(smem_u32(sK + stage * Cfg::K_STAGE)) + 2u * ks;
          tc_mma_f16_ss(tS, adesc, bdesc, idesc_s, ks > 0 ? 1u : 0u);
        }
        tc_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == KST) {
        nstage = 0;
        nphase ^= 1u;
      }
      if (j + 1 < n_tiles) {
        mbar_wait(&kv_full[nstage], nphase);
        mbar_wait(s_free, static_cast<uint32_t>(j & 1));
        tc_fence_after();
        issue_s(nstage);
      }
      const int kv = p.lkv - j * BKV;   // keys in this tile
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&p_full[h], static_cast<uint32_t>(j & 1));
        tc_fence_after();
        if (elect_one()) {
          const int kvh = kv - 64 * h;                                   // keys of this half
          const int ksteps = (kvh >= 64) ? 4 : (kvh <= 0 ? 0 : ((kvh + 31) >> 5) * 2);
          const uint32_t tO = tmem_base + O_COL + static_cast<uint32_t>(64 * h);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks >= ksteps) break;
            const uint64_t adesc = make_sdesc_sw128(smem_u32(sP + h * ATT_BQ * 128)) + 2u * ks;
            const uint64_t bdesc =
                make_sdesc_sw128(smem_u32(sV + stage * Cfg::V_STAGE + h * Cfg::V_CHUNK)) + 2u * ks;
            tc_mma_f16_ss(tO, adesc, bdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(&pv_done[h]);
          if (h == 1) tc_commit(&kv_empty[stage]);
        }
        __syncwarp();
      }
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ======================= softmax streams =======================
    const int h = (warp - 2) >> 2;        // which 64-key half of every tile
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + S_COL + static_cast<uint32_t>(64 * h);
    const uint32_t tO = tmem_base + lane_off + O_COL + static_cast<uint32_t>(64 * h);
    float m_used = -INFINITY;
    float l_sum = 0.f;
    const uint32_t prow = smem_u32(sP) + static_cast<uint32_t>(h) * (ATT_BQ * 128) + static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, static_cast<uint32_t>(j & 1));
      tc_fence_after();
      const int nvalid = p.lkv - j * BKV - 64 * h;   // keys of this half in this tile (may be <= 0)
      const int nch = (nvalid >= 64) ? 2 : (nvalid <= 0 ? 0 : ((nvalid + 31) >> 5));
      uint32_t v[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c)
        if (c < nch) tmem_ld_x32(tS + c * 32, v[c]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      if (nvalid < 64) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (c < nch && c * 32 + e >= nvalid) v[c][e] = 0xff800000u;  // -inf
          }
        }
      }
      float mx;
      {
        float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < nch) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mp[e & 3] = fmaxf(mp[e & 3], __uint_as_float(v[c][e]));
          }
        }
        mx = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
      }
      mx *= p.scale_log2;
      float alpha = 1.0f;
      bool rescale = false;
      if (nch > 0) {
        if (j == 0) {
          m_used = mx;
        } else if (mx > m_used + ATT_RESCALE_THRESHOLD) {
          alpha = ex2_approx(m_used - mx);
          m_used = mx;
          if constexpr (!ONES) l_sum *= alpha;
          rescale = true;
        }
      }
      // P_h(j-1) V_h(j-1) must have retired before P_h (single buffer) or O_h may be touched
      if (j > 0) {
        mbar_wait(&pv_done[h], static_cast<uint32_t>((j - 1) & 1));
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, rescale)) {
        for (int c = 0; c < p.dn; c += 16) {
          uint32_t o[16];
          tmem_ld_x16(tO + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tmem_st_x16(tO + c, o);
        }
        tmem_wait_st();
      }
      float rsp[2] = {0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c >= nch) break;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 =
                ex2_approx(fmaf(__uint_as_float(v[c][q4 * 8 + 2 * e]), p.scale_log2, -m_used));
            const float p1 =
                ex2_approx(fmaf(__uint_as_float(v[c][q4 * 8 + 2 * e + 1]), p.scale_log2, -m_used));
            if constexpr (!ONES) rsp[e & 1] += p0 + p1;
            pk[e] = pack_half2(p0, p1);
          }
          const uint32_t q = static_cast<uint32_t>(c * 4 + q4);   // 16-byte piece inside the 64-key half
          st_shared_v4(prow + ((q ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      }
      if constexpr (!ONES) l_sum += rsp[0] + rsp[1];
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[h]);
    }
    // ---- merge the two streams: O / l ----
    mbar_wait(&pv_done[h], static_cast<uint32_t>((n_tiles - 1) & 1));
    tc_fence_after();
    float l_mine;
    if constexpr (ONES) {
      uint32_t o[16];
      tmem_ld_x16(tO + static_cast<uint32_t>(p.d & ~15), o);   // d % 16 == 8: the sums sit in column 8
      tmem_wait_ld();
      l_mine = __uint_as_float(o[8]);
    } else {
      l_mine = l_sum;
    }
    // every MMA has retired (both streams waited for their last P V): K / V / P shared memory is free
    // and serves as the exchange buffer: stream 1 publishes (m, l, O_1 row), stream 0 merges and stores
    named_bar_sync(1, 256);
    float* xch = reinterpret_cast<float*>(sK) + static_cast<size_t>(row) * 68;   // 68 floats per row
    if (h == 1) {
      xch[0] = m_used;
      xch[1] = l_mine;
      for (int c = 0; c < p.dn; c += 16) {
        uint32_t o[16];
        tmem_ld_x16(tO + c, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) xch[4 + c + e] = __uint_as_float(o[e]);
      }
    }
    named_bar_sync(1, 256);
    if (h == 0) {
      const float m1 = xch[0], l1 = xch[1];
      const float m = fmaxf(m_used, m1);
      const float a0 = ex2_approx(m_used - m), a1 = ex2_approx(m1 - m);
      const float inv = 1.0f / (a0 * l_mine + a1 * l1);
      const float w0 = a0 * inv, w1 = a1 * inv;
      const int qrow = q0 + row;
      __half* dst = p.out + (static_cast<long long>(blockIdx.z) * p.lq + qrow) * p.ldo + head * p.d;
      for (int c = 0; c < p.dn; c += 16) {
        uint32_t o[16];
        tmem_ld_x16(tO + c, o);
        tmem_wait_ld();
        if (qrow < p.lq) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (c + q * 8 < p.d) {
              float r[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                r[e] = __uint_as_float(o[q * 8 + e]) * w0 + xch[4 + c + q * 8 + e] * w1;
              uint4 val;
              val.x = pack_half2(r[0], r[1]);
              val.y = pack_half2(r[2], r[3]);
              val.z = pack_half2(r[4], r[5]);
              val.w = pack_half2(r[6], r[7]);
              *reinterpret_cast<uint4*>(dst + c + q * 8) = val;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

template <int KST, bool ONES>
static int launch_attn_sk(const mdk_ctx* ctx, AttnParams& p, const mdk_attn_args* a,
                          cudaStream_t stream) {
  using Cfg = AttnCfg<1, SK_BKV, KST>;
  static unsigned long long attr_set_mask = 0;   // bit d: attribute set on device d (it is a per-device property)
  const bool attr_set = ((attr_set_mask >> (ctx->device & 63)) & 1ull) != 0;
  if (!attr_set) {
    MDK_CHECK_CUDA(cudaFuncSetAttribute(attn_sk_kernel<KST, ONES>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set_mask |= 1ull << (ctx->device & 63);
  }
  if (encode_attn_maps(p, a, SK_BKV)) return -1;
  p.n_kv_tiles = (a->lkv + SK_BKV - 1) / SK_BKV;
  dim3 grid((a->lq + ATT_BQ - 1) / ATT_BQ, a->heads, a->nimg);
  MDK_CHECK_CUDA(launch_pdl(attn_sk_kernel<KST, ONES>, grid, dim3(ATT_SK_THREADS), Cfg::SMEM_BYTES, stream, p));
  count_launch();
  (void)ctx;
  return 0;
}

}  // namespace mdk

extern "C" int mdk_attn_debug_trace(void* buf, int32_t tiles) {
  mdk::g_attn_trace = static_cast<long long*>(buf);
  mdk::g_attn_trace_cap = buf ? tiles : 0;
  return 0;
}

extern "C" int mdk_attn_fwd_f16(mdk_ctx* ctx, const mdk_attn_args* a, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && a && a->q && a->k && a->vt && a->out, "mdk_attn_fwd_f16: null argument");
  MDK_REQUIRE(a->d % 8 == 0 && a->d >= 8 && a->d <= 192, "mdk_attn_fwd_f16: d=%d unsupported", a->d);
  MDK_REQUIRE(a->lq > 0 && a->lkv > 0 && a->heads > 0 && a->nimg > 0 && a->nkv > 0 && a->kv_div > 0,
              "mdk_attn_fwd_f16: empty problem");
  MDK_REQUIRE(a->heads <= 65535 && a->nimg <= 65535, "mdk_attn_fwd_f16: grid too large");
  MDK_REQUIRE((a->nimg + a->kv_div - 1) / a->kv_div <= a->nkv, "mdk_attn_fwd_f16: nkv too small");
  MDK_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldvt % 8 == 0 && a->ldo % 8 == 0,
              "mdk_attn_fwd_f16: leading dimensions must be multiples of 8");
  MDK_REQUIRE(a->ldvt >= a->lkv, "mdk_attn_fwd_f16: ldvt < lkv");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__half*>(a->out);
  p.ldo = a->ldo;
  p.lq = a->lq;
  p.lkv = a->lkv;
  p.heads = a->heads;
  p.d = a->d;
  p.dk16 = (a->d + 15) / 16;
  p.dn = p.dk16 * 16;
  p.kv_div = a->kv_div;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.vt_head_rows = a->vt_head_rows > 0 ? a->vt_head_rows : a->d;
  {
    const char* e = getenv("MDK_ATTN_POLY");
    p.poly = e ? atoi(e) : 0;
  }
  MDK_REQUIRE(p.vt_head_rows >= a->d, "mdk_attn_fwd_f16: vt_head_rows < d");
  if (a->vt_ones)
    MDK_REQUIRE(a->d % 16 == 8 && p.vt_head_rows >= a->d + 8 && a->d <= 64,
                "mdk_attn_fwd_f16: vt_ones needs d %% 16 == 8, d <= 64 and vt_head_rows >= d + 8");
  // ping-pong kernel (two query tiles per CTA) for the long self-attention sequences
  int pp;
  {
    const char* e = getenv("MDK_ATTN_PP");   // read per call: tests switch kernels in-process
    // bit 0: head_dim <= 64, bit 1: head_dim <= 128.  Measured (PERF.md): d = 80 0.818 vs 0.986 ms,
    // d = 40 2.007 vs 1.897 ms -> only the wide heads use it by default
    pp = e ? atoi(e) : 2;
  }
  if (a->lq >= 2 * ATT_BQ && a->lkv >= 2 * ATT_BQ) {
    if (a->d <= 64 && (pp & 1)) {
      if (a->vt_ones) return launch_attn_pp<1, 128, 3, true>(ctx, p, a, stream);
      return launch_attn_pp<1, 128, 3, false>(ctx, p, a, stream);
    }
    if (a->d > 64 && a->d <= 128 && (pp & 2)) return launch_attn_pp<2, 64, 3, false>(ctx, p, a, stream);
  }
  if (a->d <= 64) {
    // Two softmax streams per CTA, P in tensor memory, 32-key sub-tiles (attn_2s32.cu): the default for long key
    // sequences with the ones-row V^T (the L0 spatial self-attention: 1.81 ms vs 1.87 ms for attn_2s.cu and 1.94 ms
    // for the one-stream kernel, 8 images of 9216 tokens, d = 40, medians of alternating runs under the power
    // cap), with every fourth pair of exponentials on the FMA pipe.  The 257-token cross-attention stays on the
    // one-stream kernels (2.73 vs 2.93 ms per step).  MDK_ATTN_2S=0 switches it off; =1 attn_2s.cu (64-key
    // chunks), =2 its self-issuing variant, =3 attn_2s32.cu, each for every head_dim <= 64 problem;
    // MDK_ATTN_POLY overrides the polynomial share, MDK_ATTN_STAGGER de-phases the streams (no gain measured).
    const char* e = getenv("MDK_ATTN_2S");
    const int two = e ? atoi(e) : ((a->vt_ones && a->lkv >= 1024 && a->lq >= 256) ? 3 : 0);
    if (two) {
      if (!getenv("MDK_ATTN_POLY") && a->vt_ones) p.poly = 1;
      {
        const char* es = getenv("MDK_ATTN_STAGGER");
        p.stagger = es ? atoi(es) : 0;
      }
      if (two == 3) return launch_attn_2s32(ctx, p, a, stream);   // 32-key sub-tiles (attn_2s32.cu)
      const char* etr = getenv("MDK_ATTN_TRACE");
      const bool trace = etr && atoi(etr) && g_attn_trace != nullptr;
      return launch_attn_2s(ctx, p, a, stream, two, trace ? g_attn_trace : nullptr, g_attn_trace_cap);
    }
  }
  if (a->d <= 64 && a->lkv >= SK_BKV) {
    const char* e = getenv("MDK_ATTN_SK");
    const int sk = e ? atoi(e) : 0;   // measured 2.068 vs 1.989 ms (L0 self-attention): off by default
    if (sk) {
      if (a->vt_ones) return launch_attn_sk<2, true>(ctx, p, a, stream);
      return launch_attn_sk<2, false>(ctx, p, a, stream);
    }
  }
  if (a->d <= 64) {
    // 128-key tiles x 2 CTAs/SM (1.99 ms at L = 9216, n = 8) vs 64-key tiles x 3 CTAs/SM (2.13 ms); the
    // 64-key kernel wins when it pads the key sequence less (257 CLIP tokens: 320 vs 384 columns,
    // 0.391 vs 0.474 ms)
    const char* e = getenv("MDK_ATTN_BKV");
    int bkv = e ? atoi(e) : 0;
    const char* esp = getenv("MDK_ATTN_SPLITKV");   // separate K / V^T rings (written after round 1's GPU budget
    const int split_kv = esp ? atoi(esp) : 0;       // was spent: off until it has run on a B200)
    if (bkv == 0) bkv = ((a->lkv + 63) / 64 * 64 < (a->lkv + 127) / 128 * 128) ? 64 : 128;
    if (bkv == 64) {
      if (split_kv) return launch_attn<1, 64, 2, 0, true>(ctx, p, a, stream);
      return launch_attn<1, 64, 2, 0>(ctx, p, a, stream);   // 3 CTAs per SM
    }
    if (a->vt_ones) {
      const char* es = getenv("MDK_ATTN_STALE");   // read per call: tests switch kernels in-process
      const int stale = es ? atoi(es) : 0;         // off until measured on a B200
      if (stale && split_kv) return launch_attn<1, 128, 2, 2, true>(ctx, p, a, stream);
      if (stale) return launch_attn<1, 128, 2, 2>(ctx, p, a, stream);
      const char* etr = getenv("MDK_ATTN_TRACE");
      if (etr && atoi(etr) && g_attn_trace != nullptr) {   // timeline of one CTA (tests/gpu_diag.py trace_attn)
        if (split_kv) return launch_attn<1, 128, 2, 1, true, true>(ctx, p, a, stream);
        return launch_attn<1, 128, 2, 1, false, true>(ctx, p, a, stream);
      }
      if (split_kv) return launch_attn<1, 128, 2, 1, true>(ctx, p, a, stream);
      return launch_attn<1, 128, 2, 1>(ctx, p, a, stream);
    }
    if (split_kv) return launch_attn<1, 128, 2, 0, true>(ctx, p, a, stream);
    return launch_attn<1, 128, 2, 0>(ctx, p, a, stream);                  // 2 CTAs per SM
  }
  if (a->d <= 128) {
    const char* e = getenv("MDK_ATTN_BKV2");
    const int bkv2 = e ? atoi(e) : 128;
    const char* esp2 = getenv("MDK_ATTN_SPLITKV");
    const int split2 = esp2 ? atoi(esp2) : 0;
    if (bkv2 == 64) {
      if (split2) return launch_attn<2, 64, 2, 0, true>(ctx, p, a, stream);
      return launch_attn<2, 64, 2, 0>(ctx, p, a, stream);   // 2 CTAs per SM
    }
    if (split2) return launch_attn<2, 128, 2, 0, true>(ctx, p, a, stream);
    return launch_attn<2, 128, 2, 0>(ctx, p, a, stream);
  }
  {
    const char* esp3 = getenv("MDK_ATTN_SPLITKV");
    if (esp3 && atoi(esp3)) return launch_attn<3, 64, 2, 0, true>(ctx, p, a, stream);
  }
  return launch_attn<3, 64, 2, 0>(ctx, p, a, stream);
}
