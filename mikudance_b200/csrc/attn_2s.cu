// Two-stream flash attention for head_dim <= 64 (the L0 spatial self-attention, d = 40, is bound by the MUFU
// pipe: one ex2 per 160 tensor FLOPs).  One CTA = one (image, head, 128-query tile), 384 threads, 2 CTAs / SM:
//   warps 0-3 : softmax stream 0 — the FIRST 64 keys of every 128-key tile
//   warps 4-7 : softmax stream 1 — the SECOND 64 keys of every 128-key tile
//   warp 8    : TMA producer (Q once, then load units {K_u, V^T_(u-1)} through one KST-stage ring)
//   warp 9/10 : MMA issuer of stream 0 / stream 1          (warp 11: idle, completes the warpgroup for setmaxnreg)
// Each stream is a complete online softmax of its own key subset with its own S (128 x 64 fp32), P (128 x 64
// fp16) and O (128 x dn fp32) buffers in tensor memory, running maximum and row sum; the two partial results are
// merged once, in the epilogue (O = (w0 O0 + w1 O1) / (w0 l0 + w1 l1), w_g = 2^(m_g - max m)).
//
// Why two streams (round-2 timeline of the one-stream kernel, profiles/r02_first_call.log): per 128-key tile a
// softmax warp spends 1 700 cycles in the exp2 loop and 1 450 cycles in a serial chain that cannot use the MUFU
// pipe (barrier wake-ups, TMEM load, row maximum, fence + publish, the MMA round trip); with one stream per CTA
// and two CTAs per SM there are two softmax warps per SM sub-partition and the MUFU pipe idles 35 % of the time.
// Two streams per CTA put FOUR independent softmax warps on each sub-partition (64 scores per thread instead of
// 128, so the register file still holds two CTAs); nothing but the K / V^T ring couples the streams.
//
// Why P in tensor memory: ncu on the kernels that hand P to the tensor core through shared memory shows the
// shared-memory data pipe 58-65 % busy (P: 32 KB stored + 32 KB re-read per tile, out of 100 KB).  Here the
// softmax threads write their fp16 row back over the first 32 columns of the S tile they just drained
// (tcgen05.st) and P V is a TS MMA (A operand from TMEM); shared memory carries Q, K and V^T only and the 32 KB
// freed pay for a third ring stage.  S_g(t+1) overwrites P_g(t), so the MMA thread issues P_g(t) V then
// Q K_g(t+1)^T back to back (tcgen05.mma of one thread execute in issue order): one wake-up per stream and tile,
// no S-drained barrier, and S-full of tile t+1 also tells the softmax warps that P_g(t) V has retired.
// The ring is organised in load units {K_u, V^T_(u-1)}: exactly the operands of that back-to-back pair, so one
// commit after it releases the stage.
//
// Algorithmic FLOPs per launch: 4 * nimg * heads * lq * lkv * d.
#include "attn_common.cuh"

namespace mdk {

constexpr int A2S_THREADS = 384;
constexpr int A2S_BKV = 128;

struct A2SCfg {
  static constexpr int KST = 3;
  static constexpr int Q_BYTES = ATT_BQ * 128;
  static constexpr int K_STAGE = A2S_BKV * 128;
  static constexpr int V_CHUNK = 64 * 128;   // up to 64 rows (dn) of 64 keys
  static constexpr int V_STAGE = 2 * V_CHUNK;
  static constexpr int SMEM_BYTES = Q_BYTES + KST * (K_STAGE + V_STAGE) + 256;
  static constexpr uint32_t S_COL = 0;     // S_g at S_COL + 64 g; P_g = its first 32 columns
  static constexpr uint32_t O_COL = 128;   // O_g at O_COL + 64 g
  static constexpr uint32_t TMEM_COLS = 256;
  static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs must fit in one SM's shared memory");
};

// TRACE (diagnostic instantiation, MDK_ATTN_TRACE=1 + mdk_attn_debug_trace): the CTA in the middle of the grid's
// x range (head 0, image 0) records clock64() per tile t:
//   softmax stream 0 (warp 0 lane 0): 0 S full  1 S in registers  2 row max done  3 exponentials done  4 P published
//   softmax stream 1 (warp 4 lane 0): 10-14 likewise
//   MMA thread of stream 0:           5 P(t) full  6 P V(t) + Q K(t+1) issued and committed  7 next load unit landed
// POLY: 0 = every exponential on the MUFU pipe, 1 = every fourth pair of scores, 2 = every second pair through the
//       half2 polynomial on the FMA pipe (needs ONES: the row sums then come from the ones row of V^T).
// SELF: false = MMA issuer warps 9 / 10 (woken through the p_full mbarrier); true = the first warp of each softmax
//       stream issues its stream's MMAs itself after a 128-thread named barrier (one cross-warp hand-off less on
//       the P published -> S(t+1) full round trip, which is serial in every stream).
template <bool ONES, int POLY, bool SELF, bool TRACE>
__global__ void __launch_bounds__(A2S_THREADS, 2) attn_2s_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = A2SCfg;
  constexpr int KST = Cfg::KST;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mdk attn2s: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + KST * Cfg::K_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KST * Cfg::V_STAGE);
  uint64_t* q_bar = bars;                  // [1]
  uint64_t* u_full = bars + 1;             // [KST] load unit landed
  uint64_t* u_empty = bars + 1 + KST;      // [KST] both streams' MMAs on the unit have retired (2 commits)
  uint64_t* s_full = bars + 1 + 2 * KST;   // [stream]
  uint64_t* p_full = s_full + 2;           // [stream] (4 warp arrivals)
  uint64_t* pv_done = s_full + 4;          // [stream] last P V of the stream has retired
  uint64_t* x_full = s_full + 6;           // [1] stream 1 published {m, l} (4 warp arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 7);
  static_assert((1 + 2 * KST + 7) * 8 + 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int kvimg = img / p.kv_div;
  const int n0 = p.n_kv_tiles;                                       // tiles in which stream 0 has keys
  const int n1 = (p.lkv > 64) ? (p.lkv - 64 + A2S_BKV - 1) / A2S_BKV : 0;   // ... stream 1 (n0 or n0 - 1)
  const bool tr_on = TRACE && p.trace != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
  auto tr = [&](int tile, int slot) {
    if constexpr (TRACE) {
      if (tr_on && tile < p.trace_cap) p.trace[tile * 16 + slot] = clock64();
    }
  };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(q_bar, 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(&u_full[s], 1);
      mbar_init(&u_empty[s], 2);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&pv_done[g], 1);
    }
    mbar_init(x_full, 4);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_trigger();

  // Register reallocation, warpgroup-uniform and issued INSIDE each role's branch (ptxas budgets a region by
  // the setmaxnreg that dominates it): the driver warpgroup keeps 40 registers per thread, the two softmax
  // warpgroups (64 scores per thread live across the row-maximum and the exp2 passes) take 96.
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp == 8) {
      // ======================= TMA producer =======================
      // load unit u = {K_u (u < n0), V^T_(u-1) (u >= 1)}, u = 0 .. n0
      if (elect_one()) {   // (elect: a single-thread region the compiler knows — no ELECT / R2UR loop per TMA)
        mbar_expect_tx(q_bar, Cfg::Q_BYTES);
        tma_load_4d(sQ, &p.tmQ, q_bar, 0, head, q0, img);
        const uint32_t v_bytes = static_cast<uint32_t>(2 * p.dn * 128);
        const int vrow = head * p.vt_head_rows;
        int st = 0;
        uint32_t ph = 1;   // parity to wait for on u_empty: the first pass over the ring finds it free
#pragma unroll 1
        for (int u = 0; u <= n0; ++u) {
          mbar_wait(&u_empty[st], ph);
          mbar_expect_tx(&u_full[st], (u < n0 ? static_cast<uint32_t>(Cfg::K_STAGE) : 0u) + (u > 0 ? v_bytes : 0u));
          if (u < n0) tma_load_4d(sK + st * Cfg::K_STAGE, &p.tmK, &u_full[st], 0, head, u * A2S_BKV, kvimg);
          if (u > 0) {
            tma_load_3d(sV + st * Cfg::V_STAGE, &p.tmV, &u_full[st], (u - 1) * A2S_BKV, vrow, kvimg);
            tma_load_3d(sV + st * Cfg::V_STAGE + Cfg::V_CHUNK, &p.tmV, &u_full[st], (u - 1) * A2S_BKV + 64, vrow,
                        kvimg);
          }
          if (++st == KST) {
            st = 0;
            ph ^= 1u;
          }
        }
      }
    } else if (!SELF && (warp == 9 || warp == 10)) {
      // ======================= MMA issuer of stream g (one thread) =======================
      const int g = warp - 9;
      const int ng = g ? n1 : n0;
      if (ng > 0 && elect_one()) {
        const bool tr_m = (g == 0);
        const uint32_t idesc_s = make_idesc_f16(ATT_BQ, 64);
        const uint32_t idesc_o = make_idesc_f16(ATT_BQ, static_cast<uint32_t>(p.dn));
        const uint32_t tS = tmem_base + Cfg::S_COL + 64u * g;
        const uint32_t tO = tmem_base + Cfg::O_COL + 64u * g;
        const uint64_t qdesc = make_sdesc_sw128(smem_u32(sQ));
        const uint64_t kdesc0 = make_sdesc_sw128(smem_u32(sK + g * 64 * 128));
        const uint64_t vdesc0 = make_sdesc_sw128(smem_u32(sV + g * Cfg::V_CHUNK));
        const int dk16 = p.dk16;
        // unit 0: K_0
        mbar_wait(q_bar, 0);
        mbar_wait(&u_full[0], 0);
        if (p.stagger > 0) {
          // de-phase the streams: identical work per tile keeps streams that start together in lock-step (all in
          // their exp2 loops at once, then all waiting for the tensor core at once)
          const long long t_end = clock64() + static_cast<long long>(p.stagger) * (g + 2 * (blockIdx.x & 1));
          while (clock64() < t_end) {
          }
        }
        tc_fence_after();
        for (int ks = 0; ks < dk16; ++ks)
          tc_mma_f16_ss(tS, qdesc + 2u * ks, kdesc0 + 2u * ks, idesc_s, ks > 0 ? 1u : 0u);
        tc_commit(&s_full[g]);
        tc_commit(&u_empty[0]);
        int st = 1;         // stage of unit t + 1
        uint32_t ph = 0;
        int kv = p.lkv - g * 64;   // keys of this stream in tile t
#pragma unroll 1
        for (int t = 0; t < ng; ++t) {
          mbar_wait(&u_full[st], ph);   // unit t+1 = {K_(t+1), V^T_t}: normally long landed, off the critical path
          if (tr_m) tr(t, 7);
          mbar_wait(&p_full[g], static_cast<uint32_t>(t & 1));
          if (tr_m) tr(t, 5);
          tc_fence_after();
          const uint64_t vdesc = vdesc0 + static_cast<uint64_t>((st * Cfg::V_STAGE) >> 4);
          const int ksteps = (kv >= 64) ? 4 : ((kv + 31) >> 5) * 2;   // whole 32-key pieces of P
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps) tc_mma_f16_ts(tO, tS + 8u * ks, vdesc + 2u * ks, idesc_o, (t > 0 || ks > 0) ? 1u : 0u);
          }
          if (t + 1 < ng) {
            const uint64_t kdesc = kdesc0 + static_cast<uint64_t>((st * Cfg::K_STAGE) >> 4);
            for (int ks = 0; ks < dk16; ++ks)   // overwrites P_g(t): executes after P_g(t) V
              tc_mma_f16_ss(tS, qdesc + 2u * ks, kdesc + 2u * ks, idesc_s, ks > 0 ? 1u : 0u);
            tc_commit(&s_full[g]);
          } else {
            tc_commit(&pv_done[g]);
          }
          tc_commit(&u_empty[st]);
          if (tr_m) tr(t, 6);
          kv -= A2S_BKV;
          if (++st == KST) {
            st = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;\n");
    // ======================= softmax warps: stream g, lane quarter `quarter` =======================
    const int g = warp >> 2;
    const int ng = g ? n1 : n0;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + Cfg::S_COL + 64u * g;
    const uint32_t tO = tmem_base + lane_off + Cfg::O_COL + 64u * g;
    float m_used = -INFINITY;   // reference maximum (scaled, log2 domain) of this stream
    float l_sum = 0.f;
    const bool tr_sm = (quarter == 0 && lane == 0);
    const int tr0 = g * 10;
    const float scale_log2 = p.scale_log2;
    // SELF: MMA issue state of the stream's first warp
    const bool issuer = SELF && quarter == 0;
    const uint32_t idesc_s = make_idesc_f16(ATT_BQ, 64);
    const uint32_t idesc_o = make_idesc_f16(ATT_BQ, static_cast<uint32_t>(p.dn));
    const uint32_t mS = tmem_base + Cfg::S_COL + 64u * g;   // (lane field 0: MMA operand addresses)
    const uint32_t mO = tmem_base + Cfg::O_COL + 64u * g;
    int st = 1;          // stage of load unit t + 1 = {K_(t+1), V^T_t}
    uint32_t ph = 0;
    int kv = p.lkv - g * 64;   // keys of this stream in tile t
    if (issuer && ng > 0) {
      if (elect_one()) {
        const uint64_t qdesc = make_sdesc_sw128(smem_u32(sQ));
        const uint64_t kdesc0 = make_sdesc_sw128(smem_u32(sK + g * 64 * 128));
        mbar_wait(q_bar, 0);
        mbar_wait(&u_full[0], 0);
        tc_fence_after();
        for (int ks = 0; ks < p.dk16; ++ks)
          tc_mma_f16_ss(mS, qdesc + 2u * ks, kdesc0 + 2u * ks, idesc_s, ks > 0 ? 1u : 0u);
        tc_commit(&s_full[g]);
        tc_commit(&u_empty[0]);
      }
      __syncwarp();
    }

#pragma unroll 1
    for (int t = 0; t < ng; ++t) {
      if (issuer) {
        // unit t+1 = {K_(t+1), V^T_t}: normally long landed; waited for here, off the critical path
        if (elect_one()) mbar_wait(&u_full[st], ph);
        __syncwarp();
      }
      mbar_wait(&s_full[g], static_cast<uint32_t>(t & 1));
      if (tr_sm) tr(t, tr0 + 0);
      tc_fence_after();
      const int nvalid = p.lkv - (t * A2S_BKV + g * 64);   // >= 1; columns >= nvalid are padding
      const int nch = (nvalid >= 64) ? 2 : ((nvalid + 31) >> 5);
      uint32_t v[64];
      tmem_ld_x32p(tS, v);
      if (nch > 1) tmem_ld_x32p(tS + 32, v + 32);
      tmem_wait_ld();
      if (tr_sm) tr(t, tr0 + 1);
      if (__builtin_expect(nvalid < 64, 0)) {
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (e >= nch * 32 || e >= nvalid) v[e] = 0xff800000u;   // -inf
      }
      float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < 64; ++e) mp[e & 3] = fmaxf(mp[e & 3], __uint_as_float(v[e]));
      const float mx = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3])) * scale_log2;
      float alpha = 1.0f;
      bool rescale = false;
      if (t == 0) {
        m_used = mx;
      } else if (mx > m_used + ATT_RESCALE_THRESHOLD) {
        alpha = ex2_approx(m_used - mx);
        m_used = mx;
        if constexpr (!ONES) l_sum *= alpha;
        rescale = true;
      }
      if (tr_sm) tr(t, tr0 + 2);
      // O_g may be touched: S_g(t) was issued after P_g(t-1) V by the same thread, so it has retired
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
      // exp2 and fp16 pack, in place: word w of P (keys 2w, 2w+1) replaces v[w] (w <= 2w: already consumed)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c >= nch) break;
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const float x0 = fmaf(__uint_as_float(v[c * 32 + 2 * w]), scale_log2, -m_used);
          const float x1 = fmaf(__uint_as_float(v[c * 32 + 2 * w + 1]), scale_log2, -m_used);
          if ((POLY == 1 && (w & 3) == 3) || (POLY == 2 && (w & 1) == 1)) {
            v[c * 16 + w] = ex2_poly_h2(x0, x1);   // FMA pipe
          } else {
            const float p0 = ex2_approx(x0);
            const float p1 = ex2_approx(x1);
            if constexpr (!ONES) rsp[w & 1] += p0 + p1;
            v[c * 16 + w] = pack_half2(p0, p1);
          }
        }
      }
      if constexpr (!ONES) l_sum += rsp[0] + rsp[1];
      if (tr_sm) tr(t, tr0 + 3);
      if (nch > 1)
        tmem_st_x32(tS, v);
      else
        tmem_st_x16p(tS, v);
      tmem_wait_st();
      tc_fence_before();
      if constexpr (SELF) {
        named_bar_sync(1u + g, 128);   // every row of P_g(t) is in tensor memory
        if (tr_sm) tr(t, tr0 + 4);
        if (issuer) {
          tc_fence_after();
          if (elect_one()) {
            const uint64_t qdesc = make_sdesc_sw128(smem_u32(sQ));
            const uint64_t vdesc = make_sdesc_sw128(smem_u32(sV + st * Cfg::V_STAGE + g * Cfg::V_CHUNK));
            const int ksteps = (kv >= 64) ? 4 : ((kv + 31) >> 5) * 2;   // whole 32-key pieces of P
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) tc_mma_f16_ts(mO, mS + 8u * ks, vdesc + 2u * ks, idesc_o, (t > 0 || ks > 0) ? 1u : 0u);
            }
            if (t + 1 < ng) {
              const uint64_t kdesc = make_sdesc_sw128(smem_u32(sK + st * Cfg::K_STAGE + g * 64 * 128));
              for (int ks = 0; ks < p.dk16; ++ks)   // overwrites P_g(t): executes after P_g(t) V
                tc_mma_f16_ss(mS, qdesc + 2u * ks, kdesc + 2u * ks, idesc_s, ks > 0 ? 1u : 0u);
              tc_commit(&s_full[g]);
            } else {
              tc_commit(&pv_done[g]);
            }
            tc_commit(&u_empty[st]);
          }
          __syncwarp();
          if (tr_sm) tr(t, 6 + g);
        }
        kv -= A2S_BKV;
        if (++st == KST) {
          st = 0;
          ph ^= 1u;
        }
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        if (tr_sm) tr(t, tr0 + 4);
      }
    }

    // ---- epilogue: merge the two streams, O / l ----
    // {m, l} of stream 1 travel through two TMEM columns of its S tile that P does not cover (columns 32, 33 of
    // the lane = the query row both streams' threads share)
    if (g == 1) {
      if (ng > 0) {
        tmem_st_x2(tS + 32, __float_as_uint(m_used), __float_as_uint(l_sum));
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(x_full);
      }
    } else {
      mbar_wait(&pv_done[0], 0);
      float w0 = 1.0f, w1 = 0.0f, l1 = 0.0f;
      if (n1 > 0) {
        mbar_wait(&pv_done[1], 0);
        mbar_wait(x_full, 0);
        tc_fence_after();
        uint32_t um, ul;
        tmem_ld_x2(tS + 64 + 32, um, ul);
        tmem_wait_ld();
        const float m1 = __uint_as_float(um);
        l1 = __uint_as_float(ul);
        const float m = fmaxf(m_used, m1);
        w0 = ex2_approx(m_used - m);
        w1 = ex2_approx(m1 - m);
      }
      tc_fence_after();
      const uint32_t tO1 = tO + 64u;
      float inv;
      if constexpr (ONES) {
        uint32_t o[16];
        tmem_ld_x16(tO + static_cast<uint32_t>(p.d & ~15), o);   // d % 16 == 8: the sums sit in column 8
        tmem_wait_ld();
        float l = __uint_as_float(o[8]) * w0;
        if (n1 > 0) {
          tmem_ld_x16(tO1 + static_cast<uint32_t>(p.d & ~15), o);
          tmem_wait_ld();
          l += __uint_as_float(o[8]) * w1;
        }
        inv = 1.0f / l;
      } else {
        inv = 1.0f / (l_sum * w0 + l1 * w1);
      }
      w0 *= inv;
      w1 *= inv;
      const int qrow = q0 + row;
      __half* dst = p.out + (static_cast<long long>(blockIdx.z) * p.lq + qrow) * p.ldo + head * p.d;
      for (int c = 0; c < p.dn; c += 16) {
        uint32_t o[16], o1[16];
        tmem_ld_x16(tO + c, o);
        if (n1 > 0) tmem_ld_x16(tO1 + c, o1);
        tmem_wait_ld();
        float r[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          r[e] = __uint_as_float(o[e]) * w0;
          if (n1 > 0) r[e] = fmaf(__uint_as_float(o1[e]), w1, r[e]);
        }
        if (qrow < p.lq) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (c + q * 8 < p.d) {
              uint4 val;
              val.x = pack_half2(r[q * 8 + 0], r[q * 8 + 1]);
              val.y = pack_half2(r[q * 8 + 2], r[q * 8 + 3]);
              val.z = pack_half2(r[q * 8 + 4], r[q * 8 + 5]);
              val.w = pack_half2(r[q * 8 + 6], r[q * 8 + 7]);
              *reinterpret_cast<uint4*>(dst + c + q * 8) = val;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <bool ONES, int POLY, bool SELF, bool TRACE>
static int launch_attn_2s_t(AttnParams& p, const mdk_attn_args* a, cudaStream_t stream) {
  using Cfg = A2SCfg;
  MDK_CHECK_CUDA(cudaFuncSetAttribute(attn_2s_kernel<ONES, POLY, SELF, TRACE>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  if (encode_attn_maps(p, a, A2S_BKV)) return -1;
  p.n_kv_tiles = (a->lkv + A2S_BKV - 1) / A2S_BKV;
  dim3 grid((a->lq + ATT_BQ - 1) / ATT_BQ, a->heads, a->nimg);
  MDK_CHECK_CUDA(launch_pdl(attn_2s_kernel<ONES, POLY, SELF, TRACE>, grid, dim3(A2S_THREADS), Cfg::SMEM_BYTES, stream, p));
  count_launch();
  return 0;
}

// variant: 1 = MMA issuer warps, 2 = the softmax streams issue their own MMAs; trace != nullptr: timeline instantiation
int launch_attn_2s(const mdk_ctx* ctx, AttnParams& p, const mdk_attn_args* a, cudaStream_t stream, int variant,
                   long long* trace, int trace_cap) {
  (void)ctx;
  p.trace = trace;
  p.trace_cap = trace ? trace_cap : 0;
  const int poly = a->vt_ones ? p.poly : 0;
  const bool self = variant == 2;
  if (trace != nullptr && a->vt_ones)
    return self ? launch_attn_2s_t<true, 0, true, true>(p, a, stream) : launch_attn_2s_t<true, 0, false, true>(p, a, stream);
  if (!a->vt_ones)
    return self ? launch_attn_2s_t<false, 0, true, false>(p, a, stream)
                : launch_attn_2s_t<false, 0, false, false>(p, a, stream);
  if (self) {
    if (poly == 2) return launch_attn_2s_t<true, 2, true, false>(p, a, stream);
    if (poly == 1) return launch_attn_2s_t<true, 1, true, false>(p, a, stream);
    return launch_attn_2s_t<true, 0, true, false>(p, a, stream);
  }
  if (poly == 2) return launch_attn_2s_t<true, 2, false, false>(p, a, stream);
  if (poly == 1) return launch_attn_2s_t<true, 1, false, false>(p, a, stream);
  return launch_attn_2s_t<true, 0, false, false>(p, a, stream);
}

}  // namespace mdk
