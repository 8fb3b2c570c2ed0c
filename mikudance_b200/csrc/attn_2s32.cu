// Two-stream flash attention with 32-key sub-tiles (head_dim <= 64): the kernel of attn_2s.cu with the S tile of
// each stream split into two 32-column halves that are used as a DOUBLE BUFFER.
//
// In attn_2s.cu a stream is strictly serial: the softmax warps publish P(t), the MMA thread issues P(t) V and
// Q K(t+1)^T, and only when that S tile is complete can the softmax warps continue — the round trip (P published
// -> seen by the MMA thread 280, issue 420, S seen 190 = 900 cycles of a 2 400-cycle period, profiles/
// r02_attn2s_ab.log) is exposed in every stream, and with TMEM full (2 x (64 S + 64 O) columns per CTA, 2 CTAs per
// SM) there is no room for a second S buffer.  Here each stream alternates between the two halves of its 64
// columns: while the tensor core turns P_b(h) into O and computes S_b(h+2) into the same half, the softmax warps
// work on the other half, S_(1-b)(h+1), which has been ready for a whole sub-tile.  Per 32 keys a softmax thread
// loads 32 scores (not 64: ~64 registers), takes their maximum, exponentiates, writes 16 columns of P back over
// the first half of the S half-tile and publishes it; the MMA thread issues 2 K-steps of P V and 3 of Q K^T
// (N = 32) per sub-tile.
//   warps 0-3 / 4-7: softmax stream 0 / 1 (keys [0,64) / [64,128) of every 128-key tile, sub-tile b = 32-key half)
//   warp 8: TMA (load units {K_u, V^T_(u-1)}, 3-stage ring)   warp 9 / 10: MMA issuer of stream 0 / 1
// The lazy rescale of O (running maximum grew by more than 2^8) is the one place that needs P V of the PREVIOUS
// sub-tile retired (it is still in flight by design): that rare path waits on pv_done.
//
// Algorithmic FLOPs per launch: 4 * nimg * heads * lq * lkv * d.
#include "attn_common.cuh"

namespace mdk {

constexpr int A32_THREADS = 384;
constexpr int A32_BKV = 128;

struct A32Cfg {
  static constexpr int KST = 3;
  static constexpr int Q_BYTES = ATT_BQ * 128;
  static constexpr int K_STAGE = A32_BKV * 128;
  static constexpr int V_CHUNK = 64 * 128;
  static constexpr int V_STAGE = 2 * V_CHUNK;
  static constexpr int SMEM_BYTES = Q_BYTES + KST * (K_STAGE + V_STAGE) + 256;
  static constexpr uint32_t S_COL = 0;     // S_g^b at S_COL + 64 g + 32 b; P_g^b = its first 16 columns
  static constexpr uint32_t O_COL = 128;   // O_g at O_COL + 64 g
  static constexpr uint32_t TMEM_COLS = 256;
  static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs must fit in one SM's shared memory");
};

template <bool ONES, int POLY>
__global__ void __launch_bounds__(A32_THREADS, 2) attn_2s32_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = A32Cfg;
  constexpr int KST = Cfg::KST;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mdk attn2s32: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + KST * Cfg::K_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KST * Cfg::V_STAGE);
  uint64_t* q_bar = bars;                  // [1]
  uint64_t* u_full = bars + 1;             // [KST]
  uint64_t* u_empty = bars + 1 + KST;      // [KST] (2 commits: one per stream)
  uint64_t* s_full = bars + 1 + 2 * KST;   // [stream][half]
  uint64_t* p_full = s_full + 4;           // [stream][half] (4 warp arrivals)
  uint64_t* pv_done = s_full + 8;          // [stream] one phase per sub-tile (waited on by the stream itself only)
  uint64_t* o_done = s_full + 10;          // [stream] the stream's LAST P V has retired (single phase: safe to wait
                                           //          on from the other stream, whatever its progress)
  uint64_t* x_full = s_full + 12;          // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 13);
  static_assert((1 + 2 * KST + 13) * 8 + 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int kvimg = img / p.kv_div;
  const int n0 = p.n_kv_tiles;   // 128-key tiles (= load units - 1)
  // sub-tiles (32 keys) of stream g: those whose first key t*128 + g*64 + b*32 is < lkv
  auto n_sub = [&](int g) {
    const int rem = p.lkv - g * 64;            // keys from the stream's first key on
    if (rem <= 0) return 0;
    const int full = rem / A32_BKV;            // whole tiles: 2 sub-tiles each
    const int tail = rem - full * A32_BKV;     // keys of this stream's window in the last tile: (0, 128)
    return 2 * full + (tail > 32 ? 2 : (tail > 0 ? 1 : 0));
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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
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

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp == 8) {
      // ======================= TMA producer =======================
      if (elect_one()) {
        mbar_expect_tx(q_bar, Cfg::Q_BYTES);
        tma_load_4d(sQ, &p.tmQ, q_bar, 0, head, q0, img);
        const uint32_t v_bytes = static_cast<uint32_t>(2 * p.dn * 128);
        const int vrow = head * p.vt_head_rows;
        int st = 0;
        uint32_t ph = 1;
#pragma unroll 1
        for (int u = 0; u <= n0; ++u) {
          mbar_wait(&u_empty[st], ph);
          mbar_expect_tx(&u_full[st], (u < n0 ? static_cast<uint32_t>(Cfg::K_STAGE) : 0u) + (u > 0 ? v_bytes : 0u));
          if (u < n0) tma_load_4d(sK + st * Cfg::K_STAGE, &p.tmK, &u_full[st], 0, head, u * A32_BKV, kvimg);
          if (u > 0) {
            tma_load_3d(sV + st * Cfg::V_STAGE, &p.tmV, &u_full[st], (u - 1) * A32_BKV, vrow, kvimg);
            tma_load_3d(sV + st * Cfg::V_STAGE + Cfg::V_CHUNK, &p.tmV, &u_full[st], (u - 1) * A32_BKV + 64, vrow,
                        kvimg);
          }
          if (++st == KST) {
            st = 0;
            ph ^= 1u;
          }
        }
      }
    } else if (warp == 9 || warp == 10) {
      // ======================= MMA issuer of stream g (one thread) =======================
      const int g = warp - 9;
      const int nh = n_sub(g);
      if (nh > 0 && elect_one()) {
        const uint32_t idesc_s = make_idesc_f16(ATT_BQ, 32);
        const uint32_t idesc_o = make_idesc_f16(ATT_BQ, static_cast<uint32_t>(p.dn));
        const uint32_t tS = tmem_base + Cfg::S_COL + 64u * g;
        const uint32_t tO = tmem_base + Cfg::O_COL + 64u * g;
        const uint64_t qdesc = make_sdesc_sw128(smem_u32(sQ));
        // K rows [g*64 + b*32, +32) of a tile: 32 rows x 128 B = 4 KB per half; V^T chunk g, keys b*32.. : +64 B
        const uint64_t kdesc0 = make_sdesc_sw128(smem_u32(sK + g * 64 * 128));
        const uint64_t vdesc0 = make_sdesc_sw128(smem_u32(sV + g * Cfg::V_CHUNK));
        const int dk16 = p.dk16;
        auto issue_s = [&](int b, uint64_t kdesc_stage) {   // S_b = Q K_b^T  (N = 32)
          const uint64_t kd = kdesc_stage + static_cast<uint64_t>((b * 32 * 128) >> 4);
          for (int ks = 0; ks < dk16; ++ks)
            tc_mma_f16_ss(tS + 32u * b, qdesc + 2u * ks, kd + 2u * ks, idesc_s, ks > 0 ? 1u : 0u);
          tc_commit(&s_full[g * 2 + b]);
        };
        mbar_wait(q_bar, 0);
        mbar_wait(&u_full[0], 0);
        if (p.stagger > 0) {   // de-phase the streams (see attn_2s.cu)
          const long long t_end = clock64() + static_cast<long long>(p.stagger) * (g + 2 * (blockIdx.x & 1));
          while (clock64() < t_end) {
          }
        }
        tc_fence_after();
        issue_s(0, kdesc0);
        if (nh > 1) issue_s(1, kdesc0);
        tc_commit(&u_empty[0]);
        int st = 1;   // stage of unit t + 1 = {K_(t+1), V^T_t}
        uint32_t ph = 0;
#pragma unroll 1
        for (int h = 0; h < nh; ++h) {
          const int b = h & 1;
          if (b == 0) mbar_wait(&u_full[st], ph);   // normally long landed
          mbar_wait(&p_full[g * 2 + b], static_cast<uint32_t>((h >> 1) & 1));
          tc_fence_after();
          const uint64_t vdesc = vdesc0 + static_cast<uint64_t>((st * Cfg::V_STAGE) >> 4) + 4u * b;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc_mma_f16_ts(tO, tS + 32u * b + 8u * ks, vdesc + 2u * ks, idesc_o, (h > 0 || ks > 0) ? 1u : 0u);
          tc_commit(&pv_done[g]);
          if (h + 1 >= nh) tc_commit(&o_done[g]);
          if (h + 2 < nh) issue_s(b, kdesc0 + static_cast<uint64_t>((st * Cfg::K_STAGE) >> 4));
          if (b == 1 || h + 1 >= nh) {   // last sub-tile of this stream in tile t: unit t+1 is done with
            tc_commit(&u_empty[st]);
            if (++st == KST) {
              st = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;\n");
    // ======================= softmax warps: stream g, lane quarter `quarter` =======================
    const int g = warp >> 2;
    const int nh = n_sub(g);
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + Cfg::S_COL + 64u * g;
    const uint32_t tO = tmem_base + lane_off + Cfg::O_COL + 64u * g;
    float m_used = -INFINITY;
    float l_sum = 0.f;
    const float scale_log2 = p.scale_log2;

#pragma unroll 1
    for (int h = 0; h < nh; ++h) {
      const int b = h & 1;
      mbar_wait(&s_full[g * 2 + b], static_cast<uint32_t>((h >> 1) & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_x32p(tS + 32u * b, v);
      tmem_wait_ld();
      const int nvalid = p.lkv - ((h >> 1) * A32_BKV + g * 64 + b * 32);   // >= 1
      if (__builtin_expect(nvalid < 32, 0)) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e >= nvalid) v[e] = 0xff800000u;   // -inf
      }
      float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < 32; ++e) mp[e & 3] = fmaxf(mp[e & 3], __uint_as_float(v[e]));
      const float mx = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3])) * scale_log2;
      float alpha = 1.0f;
      bool rescale = false;
      if (h == 0) {
        m_used = mx;
      } else if (mx > m_used + ATT_RESCALE_THRESHOLD) {
        alpha = ex2_approx(m_used - mx);
        m_used = mx;
        if constexpr (!ONES) l_sum *= alpha;
        rescale = true;
      }
      if (__any_sync(0xffffffffu, rescale)) {
        // O_g is about to be read-modified-written: P V of the previous sub-tile (still in flight by design) must
        // have retired — the ones before it have (S of this half was issued after them)
        mbar_wait(&pv_done[g], static_cast<uint32_t>((h - 1) & 1));
        tc_fence_after();
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
      for (int w = 0; w < 16; ++w) {
        const float x0 = fmaf(__uint_as_float(v[2 * w]), scale_log2, -m_used);
        const float x1 = fmaf(__uint_as_float(v[2 * w + 1]), scale_log2, -m_used);
        if ((POLY == 1 && (w & 3) == 3) || (POLY == 2 && (w & 1) == 1)) {
          v[w] = ex2_poly_h2(x0, x1);   // FMA pipe
        } else {
          const float p0 = ex2_approx(x0);
          const float p1 = ex2_approx(x1);
          if constexpr (!ONES) rsp[w & 1] += p0 + p1;
          v[w] = pack_half2(p0, p1);
        }
      }
      if constexpr (!ONES) l_sum += rsp[0] + rsp[1];
      tmem_st_x16p(tS + 32u * b, v);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g * 2 + b]);
    }

    // ---- epilogue: merge the two streams, O / l ----
    if (g == 1) {
      if (nh > 0) {
        // columns 16, 17 of the stream's first half-tile: free once its last P has been consumed
        mbar_wait(&o_done[1], 0);
        tc_fence_after();
        tmem_st_x2(tS + 16, __float_as_uint(m_used), __float_as_uint(l_sum));
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(x_full);
      }
    } else {
      const int nh1 = n_sub(1);
      mbar_wait(&o_done[0], 0);
      float w0 = 1.0f, w1 = 0.0f, l1 = 0.0f;
      if (nh1 > 0) {
        mbar_wait(&o_done[1], 0);
        mbar_wait(x_full, 0);
        tc_fence_after();
        uint32_t um, ul;
        tmem_ld_x2(tS + 64 + 16, um, ul);
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
        tmem_ld_x16(tO + static_cast<uint32_t>(p.d & ~15), o);
        tmem_wait_ld();
        float l = __uint_as_float(o[8]) * w0;
        if (nh1 > 0) {
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
        if (nh1 > 0) tmem_ld_x16(tO1 + c, o1);
        tmem_wait_ld();
        float r[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          r[e] = __uint_as_float(o[e]) * w0;
          if (nh1 > 0) r[e] = fmaf(__uint_as_float(o1[e]), w1, r[e]);
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

template <bool ONES, int POLY>
static int launch_attn_2s32_t(AttnParams& p, const mdk_attn_args* a, cudaStream_t stream) {
  using Cfg = A32Cfg;
  MDK_CHECK_CUDA(cudaFuncSetAttribute(attn_2s32_kernel<ONES, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
  if (encode_attn_maps(p, a, A32_BKV)) return -1;
  p.n_kv_tiles = (a->lkv + A32_BKV - 1) / A32_BKV;
  dim3 grid((a->lq + ATT_BQ - 1) / ATT_BQ, a->heads, a->nimg);
  MDK_CHECK_CUDA(launch_pdl(attn_2s32_kernel<ONES, POLY>, grid, dim3(A32_THREADS), Cfg::SMEM_BYTES, stream, p));
  count_launch();
  return 0;
}

int launch_attn_2s32(const mdk_ctx* ctx, AttnParams& p, const mdk_attn_args* a, cudaStream_t stream) {
  (void)ctx;
  const int poly = a->vt_ones ? p.poly : 0;
  if (!a->vt_ones) return launch_attn_2s32_t<false, 0>(p, a, stream);
  if (poly == 2) return launch_attn_2s32_t<true, 2>(p, a, stream);
  if (poly == 1) return launch_attn_2s32_t<true, 1>(p, a, stream);
  return launch_attn_2s32_t<true, 0>(p, a, stream);
}

}  // namespace mdk
