// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]     fp16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM, 192 threads:
//   warp 0   : TMA producer  (A tile 128 x 64 and B tile BN x 64 per stage, 128B-swizzled)
//   warp 1   : MMA issuer    (one elected lane issues tcgen05.mma 128 x BN x 16, 4 per stage)
//   warps 2-9: epilogue      (tcgen05.ld of the accumulator, fused bias / row-bias / GEGLU /
//                             residual, fp16 stores through a swizzled smem transpose) — two warps
//                             per TMEM lane quarter, interleaved over the 32-column chunks; overlaps
//                             the next tile's main loop through two TMEM accumulator buffers.
// Convolution mode walks K as (tap, channel block): for every tap the A tile is one 4-D TMA box
// [bn images, bh rows, bw cols, 64 ch] shifted by (kh-1, kw-1); out-of-bounds pixels are
// zero-filled by TMA, which is exactly the conv's zero padding.  The skip concat is a second
// A source selected per channel block.
//
// CG = 2 (default for all but tiny problems): two CTAs of a cluster (one TPC) work as a pair on a
// 256 x BN tile with tcgen05.mma.cta_group::2 — each CTA stages its own 128 rows of A and only HALF of
// the B tile (BN/2 rows), the pair's tensor cores read both halves.  Per 128 x BN x 64 of MMA work an
// SM then pulls 16 KB + BN*64 B from L2 instead of 16 KB + BN*128 B: the 128-row kernel is bound by
// the ~64 B/clk L2->SM port, not by the tensor pipe (PERF.md).  The leader CTA (cluster rank 0) issues
// all MMAs; its commits are multicast to both CTAs' barriers; both CTAs run their own epilogue.
//
// Algorithmic bytes per launch: 2*(M*K + N*K + M*N [+ M*N residual]) ; FLOPs 2*M*N*K.
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"
#include "../../include/mdk.h"

namespace mdk {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int EPI_WARPS = 8;

struct GemmParams {
  CUtensorMap tmA0, tmA1, tmB;
  CUtensorMap tmO[3];  // output segments for the TMA-store epilogue (box 32 cols x 32 rows, 64B swizzle)
  int tma_store;       // 1: epilogue writes through TMA stores, 0: smem transpose + st.global
  int sbw, sbh, sbn;   // conv: the 32-row sub-box a warp stores
  int M, N;
  int kb0, kb1;  // 64-wide K blocks (per tap) from source 0 / 1
  int k0;        // K extent of source 0 (B column offset of source 1 inside a tap)
  int ktap;      // k0 + k1: B columns per tap
  int taps;
  int H, W, nimg, bw, bh, bn, tiles_w, tiles_h;
  int m_tiles, n_tiles;
  int mp_tiles;   // ceil(m_tiles / CG): M tiles per CTA of a pair
  const float* bias;
  const float* row_bias;
  int row_div, row_mod;
  const __half* residual;
  long long ldr;
  int geglu;
  int seg_cols;
  __half* out[3];
  long long ldo[3];
  int out_trans[3];
  int trans_rows;
  long long trans_ld;
  int trans_head_d, trans_head_dp;  // padded per-head rows of a transposed segment (0: dense)
  int dbg;  // MDK_GEMM_DEBUG bit 1 (perf triage only): skip the global stores of the epilogue
};

// BS ("B stationary", CG == 1, plain GEMM with K <= 320): the BN x K weight panel stays resident in shared memory and
// only the A tiles stream through the ring.  Tiles are walked n-major (all M tiles of one N tile before the next), so a
// persistent CTA reloads the panel once or twice per launch.  Why: the K = 320 linears of level 0 (M = 294 912) are
// bound by the L2 -> SM fill rate, not by HBM or the tensor pipe — ncu: 8.9 TB/s of L2 -> SM traffic, 28 % tensor,
// 2.7 TB/s DRAM — because every 128 x 160 tile re-stages its 100 KB weight panel next to 80 KB of A.
constexpr int BS_KB = 5;   // k-blocks of a resident panel (K <= 320)

template <int BN, int CG, bool BS = false>
struct GemmCfg {
  static constexpr int B_ROWS = BN / CG;                  // B rows this CTA stages
  static constexpr int B_TILE_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = BS ? A_TILE_BYTES : A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int EPI_STAGING = EPI_WARPS * 2 * 32 * 64;  // per epilogue warp: output + residual sub-tile
  static constexpr int EPI_BIAS = EPI_WARPS * 256;  // per epilogue warp: bias of the current chunk(s)
  static constexpr int PANEL_BYTES = BS ? BS_KB * B_TILE_BYTES : 0;
  static constexpr int RING_BUDGET = 232448 - EPI_STAGING - EPI_BIAS - 256 - PANEL_BYTES;
  static constexpr int STAGES_CG1 = (BN >= 192) ? 4 : (BN >= 160 ? 5 : 6);
  static constexpr int STAGES_FIT = RING_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = BS ? (STAGES_FIT > 8 ? 8 : STAGES_FIT)
                                   : ((CG == 1) ? STAGES_CG1 : (STAGES_FIT > 8 ? 8 : STAGES_FIT));
  static_assert(!BS || (CG == 1 && STAGES >= 3), "B-stationary: single CTA, at least 3 A stages");
  static constexpr int TMEM_COLS = (2 * BN <= 32)    ? 32
                                   : (2 * BN <= 64)  ? 64
                                   : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256
                                                     : 512;
  // dynamic shared memory is declared __align__(1024) (checked at run time): no alignment slack
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + PANEL_BYTES + EPI_STAGING + EPI_BIAS + 256 /*barriers*/;
  static_assert(B_TILE_BYTES % 1024 == 0, "B stage tiles must keep the 1024-byte swizzle alignment");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(2 * STAGES + 5 <= 30, "barrier block");
};

// GELU with the exact (erf) formulation of diffusers' GEGLU, x * Phi(x), evaluated as
//     x * sigmoid(2 u(x)),   u(x) = x (a1 + a3 x^2 + a5 x^4)   <=>   erf(x / sqrt 2) = tanh(u(x))
// with a minimax fit of the odd polynomial (scripts in PERF.md): max |error| of gelu over the whole real line
// 2.5e-5 — 20x below the fp16 rounding of the output — on 10 instructions (2 MUFU: ex2, rcp).  The previous
// Abramowitz-Stegun 7.1.26 form (1.5e-7) cost 19; the GEGLU epilogue evaluates 16 K of these per 128 x 256
// tile and, at K = 320, was issue-bound: 3 700 issue cycles per tile against 2 560 cycles of MMA.
// x is clamped to [-10, 10] inside u only (a5 < 0: the polynomial turns over beyond |x| = 10.4, where
// sigmoid(2u) is 0 or 1 to fp32 precision anyway).
__device__ __forceinline__ float gelu_erf(float x) {
  constexpr float K = -2.0f * 1.4426950408889634f;   // sigmoid(2u) = 1 / (1 + 2^(-2 log2(e) u))
  const float xc = fminf(fmaxf(x, -10.0f), 10.0f);
  const float x2 = xc * xc;
  float p = fmaf(x2, K * -0.0003515167879559536f, K * 0.037005646017822316f);
  p = fmaf(p, x2, K * 0.7975078842947918f);
  const float e = ex2_approx(p * xc);
  return __fdividef(x, 1.0f + e);
}

// Write a warp's 32 rows x 32 fp16 columns: every lane holds one row (o[32]).  Direct per-lane
// stores would be 16-byte pieces scattered over 32 rows (partial-sector writes: measured 5x slower
// than the rest of the kernel), so the sub-tile is transposed through a 2 KB warp-private staging
// buffer (XOR-swizzled, conflict free) and written as full 64-byte row segments: one store
// instruction covers 8 rows x 64 B.
__device__ __forceinline__ void store_chunk_coalesced(const float (&o)[32], uint32_t stage_addr,
                                                      int lane, int m_mine, __half* obase,
                                                      long long ldo, int col0, int nvalid) {
  const uint32_t my = stage_addr + static_cast<uint32_t>(lane) * 64u;
  const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    st_shared_v4(my + ((q ^ sw) << 4), pack_half2(o[q * 8 + 0], o[q * 8 + 1]),
                 pack_half2(o[q * 8 + 2], o[q * 8 + 3]), pack_half2(o[q * 8 + 4], o[q * 8 + 5]),
                 pack_half2(o[q * 8 + 6], o[q * 8 + 7]));
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = i * 32 + lane;
    const int row = t >> 2;
    const uint32_t q = static_cast<uint32_t>(t & 3);
    const int m_r = __shfl_sync(0xffffffffu, m_mine, row);
    uint4 val;
    ld_shared_v4(stage_addr + static_cast<uint32_t>(row) * 64u +
                     ((q ^ static_cast<uint32_t>((row >> 1) & 3)) << 4),
                 val);
    if (m_r >= 0 && static_cast<int>(q) * 8 < nvalid)
      *reinterpret_cast<uint4*>(obase + static_cast<long long>(m_r) * ldo + col0 + q * 8) = val;
  }
  __syncwarp();
}

// TMA-store variant: the sub-tile goes to the warp's staging buffer in the 64B-swizzled layout and
// one elected lane issues an asynchronous bulk tensor store; bounds are clipped by TMA.  The wait for
// the previous store sits at the top of the next chunk's store, i.e. behind that chunk's TMEM load,
// bias / residual adds and conversion.
__device__ __forceinline__ void store_chunk_tma(const float (&o)[32], uint32_t buf_addr, int lane, bool leader,
                                                const CUtensorMap* tm, bool conv, int c0, int c1, int c2,
                                                int c3) {
  if (leader) bulk_wait_read<0>();  // the previous store has finished reading the staging buffer
  __syncwarp();
  const uint32_t my = buf_addr + static_cast<uint32_t>(lane) * 64u;
  const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    st_shared_v4(my + ((q ^ sw) << 4), pack_half2(o[q * 8 + 0], o[q * 8 + 1]),
                 pack_half2(o[q * 8 + 2], o[q * 8 + 3]), pack_half2(o[q * 8 + 4], o[q * 8 + 5]),
                 pack_half2(o[q * 8 + 6], o[q * 8 + 7]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (leader) {
    if (conv)
      tma_store_4d(tm, buf_addr, c0, c1, c2, c3);
    else
      tma_store_2d(tm, buf_addr, c0, c1);
    bulk_commit();
  }
}

template <int BN, int CG, bool BS = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, CG, BS>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // 1024-byte alignment required by the 128B swizzle atoms
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mdk gemm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_TILE_BYTES;     // BS: the resident panel (BS_KB k-block slices)
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES + Cfg::PANEL_BYTES;
  uint8_t* smem_bias = smem_epi + Cfg::EPI_STAGING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + Cfg::EPI_BIAS);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint64_t* bfull_bar = bars + 2 * STAGES + 4;   // [1] BS: the weight panel has landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 is the leader (issues the MMAs, owns the full / tmem-empty barriers in use)
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_tile = (CG == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = (CG == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB);
    if (p.kb1 > 0) tma_prefetch_desc(&p.tmA1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_WARPS * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    mbar_init(bfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    if constexpr (CG == 2)
      tmem_alloc_cg2<Cfg::TMEM_COLS>(tmem_ptr_smem);
    else
      tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  }
  tc_fence_before();
  if constexpr (CG == 2)
    cluster_sync_all();   // the peer's barriers are initialised before anything may signal them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();      // everything above touched no global memory: it overlaps the previous kernel's tail
  pdl_trigger();

  const int total_tiles = p.mp_tiles * p.n_tiles;   // tiles per CTA (pair tiles for CG == 2)
  const int kblocks_per_tap = p.kb0 + p.kb1;
  const int num_kb = p.taps * kblocks_per_tap;

  if (warp == 0) {
    // ======================= TMA producer =======================
    // elect.sync (not `lane == 0`): the compiler then knows a single thread runs this region and
    // emits the uniform-datapath TMA / tcgen05 instructions directly instead of wrapping each one in
    // an ELECT + R2UR.BROADCAST loop (measured: that loop, not the tensor pipe, paced the main loop)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // pair mode: transaction bytes of both CTAs are counted on the leader's full barrier
      const uint32_t full0_cluster = (CG == 2) ? mapa_shared(smem_u32(&full_bar[0]), 0) : 0u;
      int cur_n = -1, last_stage = -1;   // BS: N tile of the resident panel; stage / phase of the last A load
      uint32_t last_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int n_tile = BS ? tile / p.mp_tiles : tile % p.n_tiles;
        const int m_tile = BS ? tile % p.mp_tiles : (tile / p.n_tiles) * CG + cta_rank;   // may be a phantom tile past M
        const int n0 = n_tile * BN + cta_rank * Cfg::B_ROWS;      // pair mode: this CTA's half of the B tile
        if constexpr (BS) {
          if (n_tile != cur_n) {
            // every MMA that reads the old panel has retired once the last A stage handed out is free again
            if (last_stage >= 0) mbar_wait(&empty_bar[last_stage], last_phase);
            mbar_expect_tx(bfull_bar, static_cast<uint32_t>(num_kb) * Cfg::B_TILE_BYTES);
            for (int kb = 0; kb < num_kb; ++kb)
              tma_load_2d(smem_b + kb * Cfg::B_TILE_BYTES, &p.tmB, bfull_bar, kb * BK, n0);
            cur_n = n_tile;
          }
        }
        int m0 = m_tile * BM;
        int img0 = 0, h0 = 0, w0 = 0;
        if (p.taps > 1) {
          const int tw = m_tile % p.tiles_w;
          const int th = (m_tile / p.tiles_w) % p.tiles_h;
          const int tn = m_tile / (p.tiles_w * p.tiles_h);
          w0 = tw * p.bw;
          h0 = th * p.bh;
          img0 = tn * p.bn;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dh = (p.taps > 1) ? (tap / 3 - 1) : 0;
          const int dw = (p.taps > 1) ? (tap % 3 - 1) : 0;
          for (int kb = 0; kb < kblocks_per_tap; ++kb) {
            const bool src1 = kb >= p.kb0;
            const int kk = (src1 ? (kb - p.kb0) : kb) * BK;
            const int bcol = tap * p.ktap + (src1 ? p.k0 + kk : kk);
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            const CUtensorMap* ta = src1 ? &p.tmA1 : &p.tmA0;
            if constexpr (CG == 2) {
              if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              const uint32_t fb = full0_cluster + static_cast<uint32_t>(stage) * 8u;
              if (p.taps > 1) {
                tma_load_4d_cg2(smem_a + stage * A_TILE_BYTES, ta, fb, kk, w0 + dw, h0 + dh, img0);
              } else {
                tma_load_2d_cg2(smem_a + stage * A_TILE_BYTES, ta, fb, kk, m0);
              }
              tma_load_2d_cg2(smem_b + stage * Cfg::B_TILE_BYTES, &p.tmB, fb, bcol, n0);
            } else {
              mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              if (p.taps > 1) {
                tma_load_4d(smem_a + stage * A_TILE_BYTES, ta, &full_bar[stage], kk, w0 + dw, h0 + dh,
                            img0);
              } else {
                tma_load_2d(smem_a + stage * A_TILE_BYTES, ta, &full_bar[stage], kk, m0);
              }
              if constexpr (!BS) tma_load_2d(smem_b + stage * Cfg::B_TILE_BYTES, &p.tmB, &full_bar[stage], bcol, n0);
              last_stage = stage;
              last_phase = phase;
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (pair mode: leader CTA only) =======================
    constexpr uint32_t idesc = make_idesc_f16(BM * CG, BN);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_phase[2] = {0, 0};
    int it = 0;
    int cur_n = -1;          // BS: N tile of the resident panel
    uint32_t bphase = 0;
    for (int tile = (cta_rank == 0) ? first_tile : total_tiles; tile < total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      if constexpr (BS) {
        if (tile / p.mp_tiles != cur_n) {
          mbar_wait(bfull_bar, bphase);
          bphase ^= 1u;
          cur_n = tile / p.mp_tiles;
        }
      }
      mbar_wait(&tempty_bar[acc], acc_phase[acc] ^ 1u);
      acc_phase[acc] ^= 1u;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_sdesc_sw128(smem_u32(smem_a + stage * A_TILE_BYTES));
          const uint64_t bdesc = make_sdesc_sw128(smem_u32(smem_b + (BS ? kb : stage) * Cfg::B_TILE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in 16-byte units
            if constexpr (CG == 2)
              tc_mma_f16_ss_cg2(d_tmem, adesc + static_cast<uint64_t>(2 * k),
                                bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else
              tc_mma_f16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k),
                            bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CG == 2) {
            tc_commit_cg2(&empty_bar[stage], 3);                       // frees the stage in BOTH CTAs
            if (kb == num_kb - 1) tc_commit_cg2(&tfull_bar[acc], 3);   // both epilogues may start
          } else {
            tc_commit(&empty_bar[stage]);                       // frees the smem stage when MMAs retire
            if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);   // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ======================= epilogue warps =======================
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int cgroup = (warp - 2) >> 2;     // 0/1: which half of the 32-column chunks this warp takes
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t stage_addr = smem_u32(smem_epi) + static_cast<uint32_t>(warp - 2) * 4096u;
    const uint32_t bias_addr = smem_u32(smem_bias) + static_cast<uint32_t>(warp - 2) * 256u;
    uint32_t acc_phase[2] = {0, 0};
    int it = 0;
    const bool leader = elect_one();   // the lane that issues this warp's TMA stores / barrier arrivals
    // pair mode: "accumulator drained" is reported to the leader's barrier (16 arrivals per tile)
    const uint32_t tempty0_cluster = (CG == 2) ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const int n_tile = BS ? tile / p.mp_tiles : tile % p.n_tiles;
      const int m_tile = BS ? tile % p.mp_tiles : (tile / p.n_tiles) * CG + cta_rank;
      const int n0 = n_tile * BN;
      long long m;  // global output row of this thread, -1 if out of range
      if (p.taps > 1) {
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int wi = row_in_tile % p.bw;
        const int hi = (row_in_tile / p.bw) % p.bh;
        const int ni = row_in_tile / (p.bw * p.bh);
        const int img = tn * p.bn + ni;
        m = (img < p.nimg)
                ? (static_cast<long long>(img) * p.H + (th * p.bh + hi)) * p.W + (tw * p.bw + wi)
                : -1;
      } else {
        m = static_cast<long long>(m_tile) * BM + row_in_tile;
        if (m >= p.M) m = -1;
      }
      // TMA-store coordinates of this warp's 32-row sub-tile
      int tc1, tc2 = 0, tc3 = 0;
      if (p.taps > 1) {
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int r0 = quarter * 32;
        tc1 = tw * p.bw;
        tc2 = th * p.bh + (r0 / p.bw) % p.bh;
        tc3 = tn * p.bn + r0 / (p.bw * p.bh);
      } else {
        tc1 = m_tile * BM + quarter * 32;
      }
      // per-row bias (temporal PE): when a warp's 32 rows share one group (row_div % 32 == 0 in plain
      // GEMM mode) it is folded into the staged per-column bias; otherwise every thread loads its own
      const float* rb = nullptr;       // per-thread path
      const float* rb_warp = nullptr;  // warp-uniform path
      if (p.row_bias != nullptr) {
        if (p.taps == 1 && (p.row_div & 31) == 0) {
          const long long m_first = static_cast<long long>(m_tile) * BM + quarter * 32;
          if (m_first < p.M) rb_warp = p.row_bias + static_cast<long long>((m_first / p.row_div) % p.row_mod) * p.N;
        } else if (m >= 0) {
          rb = p.row_bias + static_cast<long long>((m / p.row_div) % p.row_mod) * p.N;
        }
      }

      const uint32_t t_acc =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);

      const int m32 = static_cast<int>(m);  // M fits in int32 (checked on the host)
      if (BN == 256 && p.geglu) {   // (the host only launches GEGLU with 256-wide tiles)
        // tile columns [0, BN/2) are values, [BN/2, BN) the matching gates
        constexpr int HALF = BN / 2;
        constexpr int NCHUNK = (HALF + 63) / 64;   // 32-column chunks this warp handles per tile
        const int ocol0 = n_tile * HALF;
        // the bias of this warp's value / gate columns is fetched BEFORE the wait for the accumulator
        // (one column per lane); inside the chunk loop a global load would sit exposed between the
        // TMEM load and its use (measured: the GEGLU epilogue, not the main loop, set the tile time)
        float bpre_h[NCHUNK], bpre_g[NCHUNK];
#pragma unroll
        for (int i = 0; i < NCHUNK; ++i) {
          const int cc = cgroup * 32 + i * 64;
          const bool okc = p.bias != nullptr && cc < HALF && n0 + cc < p.N;
          bpre_h[i] = okc ? __ldg(p.bias + n0 + cc + lane) : 0.f;
          bpre_g[i] = okc ? __ldg(p.bias + n0 + HALF + cc + lane) : 0.f;
        }
        mbar_wait(&tfull_bar[acc], acc_phase[acc]);
        acc_phase[acc] ^= 1u;
        tc_fence_after();
        uint32_t vh[32], vg[32];
        if (cgroup * 32 < HALF && n0 + cgroup * 32 < p.N) {
          tmem_ld_x32(t_acc + cgroup * 32, vh);
          tmem_ld_x32(t_acc + HALF + cgroup * 32, vg);
        }
#pragma unroll
        for (int i = 0; i < NCHUNK; ++i) {
          const int c = cgroup * 32 + i * 64;
          if (c >= HALF || n0 + c >= p.N) break;
          if (p.bias) {
            sts32(bias_addr + static_cast<uint32_t>(lane) * 4u, __float_as_uint(bpre_h[i]));
            sts32(bias_addr + 128u + static_cast<uint32_t>(lane) * 4u, __float_as_uint(bpre_g[i]));
          }
          __syncwarp();
          tmem_wait_ld();
          float o[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 bh = make_float4(0.f, 0.f, 0.f, 0.f), bg = bh;
            if (p.bias) {
              uint4 r0, r1;
              ld_shared_v4(bias_addr + static_cast<uint32_t>(j4) * 16u, r0);
              ld_shared_v4(bias_addr + 128u + static_cast<uint32_t>(j4) * 16u, r1);
              bh = make_float4(__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z), __uint_as_float(r0.w));
              bg = make_float4(__uint_as_float(r1.x), __uint_as_float(r1.y), __uint_as_float(r1.z), __uint_as_float(r1.w));
            }
            o[j4 * 4 + 0] = (__uint_as_float(vh[j4 * 4 + 0]) + bh.x) * gelu_erf(__uint_as_float(vg[j4 * 4 + 0]) + bg.x);
            o[j4 * 4 + 1] = (__uint_as_float(vh[j4 * 4 + 1]) + bh.y) * gelu_erf(__uint_as_float(vg[j4 * 4 + 1]) + bg.y);
            o[j4 * 4 + 2] = (__uint_as_float(vh[j4 * 4 + 2]) + bh.z) * gelu_erf(__uint_as_float(vg[j4 * 4 + 2]) + bg.z);
            o[j4 * 4 + 3] = (__uint_as_float(vh[j4 * 4 + 3]) + bh.w) * gelu_erf(__uint_as_float(vg[j4 * 4 + 3]) + bg.w);
          }
          // the next chunk's accumulator load overlaps this chunk's residual add and store
          if (i + 1 < NCHUNK && c + 64 < HALF && n0 + c + 64 < p.N) {
            tmem_ld_x32(t_acc + c + 64, vh);
            tmem_ld_x32(t_acc + HALF + c + 64, vg);
          }
          __syncwarp();
          if (p.residual && m >= 0) {
            const __half* rsrc = p.residual + m * p.ldr + ocol0 + c;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 rv = *reinterpret_cast<const uint4*>(rsrc + q * 8);
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(rh[e]);
                o[q * 8 + 2 * e] += f.x;
                o[q * 8 + 2 * e + 1] += f.y;
              }
            }
          }
          if (!(p.dbg & 1)) {
            if (p.tma_store) {
              store_chunk_tma(o, stage_addr, lane, leader, &p.tmO[0], p.taps > 1, ocol0 + c, tc1, tc2, tc3);
            } else {
              store_chunk_coalesced(o, stage_addr, lane, m32, p.out[0], p.ldo[0], ocol0 + c, 32);
            }
          }
        }
      } else {
        const int seg = (p.seg_cols > 0) ? (n0 / p.seg_cols) : 0;
        const int seg_col0 = (p.seg_cols > 0) ? (n0 - seg * p.seg_cols) : n0;
        __half* obase = p.out[seg];
        const long long ldo = p.ldo[seg];
        const bool trans = p.out_trans[seg] != 0;
        // Everything the epilogue needs besides the accumulator is fetched ahead of use and handed
        // over through warp-private shared memory: the chunk's 32 bias values (one per lane, read
        // back as broadcasts) and the residual sub-tile (coalesced 64-byte row segments, read back
        // as this lane's own row).  ncu showed the epilogue stalled on exactly these global loads.
        const bool res_staged = p.residual != nullptr;
        const uint32_t res_buf = stage_addr + 2048u;
        uint32_t v[32];
        float bv = 0.f;
        uint4 rres[4];
        int c = cgroup * 32;
        auto prefetch_aux = [&](int cc) {
          const int col = n0 + cc;
          bv = (p.bias != nullptr && col + lane < p.N) ? __ldg(p.bias + col + lane) : 0.f;
          if (rb_warp != nullptr && col + lane < p.N) bv += __ldg(rb_warp + col + lane);
          if (res_staged) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int t = i * 32 + lane;
              const int row = t >> 2, q = t & 3;
              const int m_r = __shfl_sync(0xffffffffu, static_cast<int>(m), row);
              rres[i] = (m_r >= 0 && col + q * 8 < p.N)
                            ? *reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(m_r) * p.ldr +
                                                              col + q * 8)
                            : make_uint4(0, 0, 0, 0);
            }
          }
        };
        const bool any_chunk = c < BN && n0 + c < p.N;
        if (any_chunk) prefetch_aux(c);          // overlaps the wait for the accumulator
        mbar_wait(&tfull_bar[acc], acc_phase[acc]);
        acc_phase[acc] ^= 1u;
        tc_fence_after();
        if (any_chunk) tmem_ld_x32(t_acc + c, v);
        for (; c < BN; c += 64) {
          if (n0 + c >= p.N) break;  // warp-uniform
          tmem_wait_ld();
          float acc_f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc_f[j] = __uint_as_float(v[j]);
          sts32(bias_addr + static_cast<uint32_t>(lane) * 4u, __float_as_uint(bv));
          if (res_staged) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int t = i * 32 + lane;
              const uint32_t row = static_cast<uint32_t>(t >> 2), q = static_cast<uint32_t>(t & 3);
              st_shared_v4(res_buf + row * 64u + ((q ^ ((row >> 1) & 3u)) << 4), rres[i].x, rres[i].y,
                           rres[i].z, rres[i].w);
            }
          }
          __syncwarp();
          // prefetch this warp's next chunk while the current one is converted and stored
          const bool has_next = c + 64 < BN && n0 + c + 64 < p.N;
          if (has_next) {
            tmem_ld_x32(t_acc + c + 64, v);
            prefetch_aux(c + 64);
          }
          const int nvalid = min(32, p.N - (n0 + c));   // multiple of 8
          float o[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            uint4 braw;
            ld_shared_v4(bias_addr + static_cast<uint32_t>(j4) * 16u, braw);
            float4 b = make_float4(__uint_as_float(braw.x), __uint_as_float(braw.y), __uint_as_float(braw.z),
                                   __uint_as_float(braw.w));
            if (rb && j4 * 4 < nvalid) {
              const float4 r = __ldg(reinterpret_cast<const float4*>(rb + n0 + c) + j4);
              b.x += r.x; b.y += r.y; b.z += r.z; b.w += r.w;
            }
            o[j4 * 4 + 0] = acc_f[j4 * 4 + 0] + b.x;
            o[j4 * 4 + 1] = acc_f[j4 * 4 + 1] + b.y;
            o[j4 * 4 + 2] = acc_f[j4 * 4 + 2] + b.z;
            o[j4 * 4 + 3] = acc_f[j4 * 4 + 3] + b.w;
          }
          if (res_staged) {
            const uint32_t myrow = res_buf + static_cast<uint32_t>(lane) * 64u;
            const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
              uint4 rv;
              ld_shared_v4(myrow + ((q ^ sw) << 4), rv);
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(rh[e]);
                o[q * 8 + 2 * e] += f.x;
                o[q * 8 + 2 * e + 1] += f.y;
              }
            }
          } else if (p.residual && m >= 0) {
            const __half* rsrc = p.residual + m * p.ldr + n0 + c;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q * 8 < nvalid) {
                uint4 rv = *reinterpret_cast<const uint4*>(rsrc + q * 8);
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __half22float2(rh[e]);
                  o[q * 8 + 2 * e] += f.x;
                  o[q * 8 + 2 * e + 1] += f.y;
                }
              }
            }
          }
          if (p.dbg & 1) continue;
          if (!trans) {
            if (p.tma_store) {
              store_chunk_tma(o, stage_addr, lane, leader, &p.tmO[seg], p.taps > 1, seg_col0 + c, tc1, tc2, tc3);
            } else {
              store_chunk_coalesced(o, stage_addr, lane, m32, obase, ldo, seg_col0 + c, nvalid);
            }
          } else {
            // per-image transposed store (V^T for the attention kernels).  Fast path — the warp's 32 rows are 32
            // consecutive, 8-aligned positions of ONE image: the 32 x 32 sub-tile is transposed through the warp's
            // staging buffer ([column][row], 64 B per column) and written as 16-byte pieces, 8 columns x 64 B per
            // store instruction (the direct path below issues 32 two-byte stores per lane: 207 TFLOP/s on the
            // level-0 k|v^T projection where the plain GEMM reaches 620).
            const int m_first = __shfl_sync(0xffffffffu, m32, 0);
            const int m_last = __shfl_sync(0xffffffffu, m32, 31);
            const int l_first = (m_first >= 0) ? m_first % p.trans_rows : -1;
            const bool fast = m_first >= 0 && m_last == m_first + 31 && (l_first & 7) == 0 &&
                              l_first + 32 <= p.trans_rows && (p.trans_ld & 7) == 0 && nvalid == 32 &&
                              (reinterpret_cast<uintptr_t>(obase) & 15) == 0;
            if (fast) {
              if (p.tma_store && leader) bulk_wait_read<0>();   // an earlier TMA store may still read the buffer
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const __half hv = __float2half_rn(o[j]);
                asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(stage_addr + static_cast<uint32_t>(j) * 64u +
                                                               static_cast<uint32_t>(lane) * 2u),
                             "h"(*reinterpret_cast<const unsigned short*>(&hv))
                             : "memory");
              }
              __syncwarp();
              const long long img = m_first / p.trans_rows;
              const int hd = p.trans_head_d, hdp = p.trans_head_dp;
              const long long rows_per_img = (hdp > 0) ? static_cast<long long>(p.seg_cols / hd) * hdp : p.seg_cols;
              __half* ibase = obase + img * rows_per_img * p.trans_ld + l_first;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int col = i * 8 + (lane >> 2);      // column of the sub-tile
                const int piece = lane & 3;               // 8 rows = 16 bytes
                const int ch = seg_col0 + c + col;        // channel inside the segment
                const long long vrow = (hdp > 0) ? static_cast<long long>(ch / hd) * hdp + (ch % hd) : ch;
                uint4 val;
                ld_shared_v4(stage_addr + static_cast<uint32_t>(col) * 64u + static_cast<uint32_t>(piece) * 16u, val);
                *reinterpret_cast<uint4*>(ibase + vrow * p.trans_ld + piece * 8) = val;
              }
              __syncwarp();
            } else if (m >= 0) {
              // direct path: lanes hold consecutive rows -> 64-byte runs per column
              const long long img = m / p.trans_rows;
              const long long l = m % p.trans_rows;
              if (p.trans_head_dp > 0) {
                // channel (seg_col0 + c + j) = head * d + w  ->  row head * dp + w of this image
                const int hd = p.trans_head_d, hdp = p.trans_head_dp;
                int head = (seg_col0 + c) / hd, wch = (seg_col0 + c) % hd;
                __half* ibase = obase + img * (p.seg_cols / hd) * hdp * p.trans_ld + l;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j < nvalid)
                    ibase[static_cast<long long>(head * hdp + wch) * p.trans_ld] = __float2half_rn(o[j]);
                  if (++wch == hd) {
                    wch = 0;
                    ++head;
                  }
                }
              } else {
                __half* dst = obase + (img * p.seg_cols + seg_col0 + c) * p.trans_ld + l;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j < nvalid) dst[static_cast<long long>(j) * p.trans_ld] = __float2half_rn(o[j]);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (leader) {
        if constexpr (CG == 2)
          mbar_arrive_cluster(tempty0_cluster + static_cast<uint32_t>(acc) * 8u);
        else
          mbar_arrive(&tempty_bar[acc]);
      }
    }
    if (p.tma_store && leader) bulk_wait<0>();
  }

  tc_fence_before();
  if constexpr (CG == 2)
    cluster_sync_all();   // neither CTA exits (or frees TMEM) while its peer may still signal / read it
  else
    __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (CG == 2)
      tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    else
      tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Tile width: among the widths that divide the segment width, minimise (waves on this GPU) x
// (tile width ~ tile time) with the padded columns charged in full; ties go to the wider tile
// (fewer A re-reads from L2).  E.g. M=4608, N=1280: 128x256 tiles need 2 waves of 180 tiles
// (cost 512) while 128x160 tiles fill 1.95 waves of 288 (cost 320).
static int pick_bn(int n, int seg_cols, int nseg, int m_tiles, int num_sms) {
  const int cands[] = {256, 192, 160, 128, 64, 32};
  int best = 0;
  long long best_cost = 0;
  for (int bn : cands) {
    if (nseg > 1 && (seg_cols % bn) != 0) continue;
    const long long n_tiles = (n + bn - 1) / bn;
    const long long waves = (n_tiles * m_tiles + num_sms - 1) / num_sms;
    const long long cost = waves * (bn < 128 ? 128 : bn);  // narrow tiles do not get cheaper: smem-bound MMA
    if (best == 0 || cost < best_cost) {
      best = bn;
      best_cost = cost;
    }
  }
  return best;
}

// largest power of two <= cap that divides x
static int pow2_div(int x, int cap) {
  int r = 1;
  while (r * 2 <= cap && (x % (r * 2)) == 0) r *= 2;
  return r;
}

template <int BN, int CG, bool BS = false>
static int launch_gemm(const mdk_ctx* ctx, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG, BS>;
  static unsigned long long attr_set_mask = 0;   // bit d: attribute set on device d (it is a per-device property)
  const bool attr_set = ((attr_set_mask >> (ctx->device & 63)) & 1ull) != 0;
  static int max_clusters = 0;
  if (!attr_set) {
    MDK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CG, BS>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    if (CG == 2) {
      // how many CTA pairs fit on the device at once (persistent kernel: launch no more than that)
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(static_cast<unsigned>(ctx->num_sms / 2 * 2));
      qc.blockDim = dim3(GEMM_THREADS);
      qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int n = 0;
      MDK_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, CG, BS>, &qc));
      MDK_REQUIRE(n > 0, "mdk_gemm_f16: no CTA pair of the 2-CTA GEMM fits on this device");
      max_clusters = n < ctx->num_sms / 2 ? n : ctx->num_sms / 2;
    }
    attr_set_mask |= 1ull << (ctx->device & 63);
  }
  const int total = p.mp_tiles * p.n_tiles;
  if (CG == 2) {
    const int clusters = total < max_clusters ? total : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MDK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CG, BS>, p));
    count_launch();
    return 0;
  }
  const int grid = total < ctx->num_sms ? total : ctx->num_sms;
  MDK_CHECK_CUDA(launch_pdl(gemm_tc_kernel<BN, CG, BS>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, p));
  count_launch();
  return 0;
}

}  // namespace mdk

extern "C" int mdk_gemm_geglu_block(void) { return 256; }

extern "C" int mdk_gemm_f16(mdk_ctx* ctx, const mdk_gemm_args* a, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && a, "mdk_gemm_f16: null ctx/args");
  MDK_REQUIRE(a->a0 && a->b && a->out[0], "mdk_gemm_f16: null operand");
  MDK_REQUIRE(a->m > 0 && a->n > 0 && a->k0 > 0, "mdk_gemm_f16: empty problem m=%d n=%d k0=%d",
              a->m, a->n, a->k0);
  MDK_REQUIRE(a->k0 % 8 == 0 && a->k1 % 8 == 0 && a->n % 8 == 0,
              "mdk_gemm_f16: k0=%d k1=%d n=%d must be multiples of 8", a->k0, a->k1, a->n);
  MDK_REQUIRE(a->ldb % 8 == 0, "mdk_gemm_f16: ldb=%lld must be a multiple of 8", (long long)a->ldb);
  MDK_REQUIRE(a->conv_taps == 1 || a->conv_taps == 9, "mdk_gemm_f16: conv_taps must be 1 or 9");
  MDK_REQUIRE((reinterpret_cast<uintptr_t>(a->a0) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->b) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->a1) & 15) == 0,
              "mdk_gemm_f16: operands must be 16-byte aligned");
  const int ktap = a->k0 + a->k1;
  const long long K = static_cast<long long>(a->conv_taps) * ktap;
  MDK_REQUIRE(a->ldb >= K, "mdk_gemm_f16: ldb=%lld < K=%lld", (long long)a->ldb, K);

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->m;
  p.N = a->n;
  p.k0 = a->k0;
  p.ktap = ktap;
  p.kb0 = (a->k0 + BK - 1) / BK;
  p.kb1 = (a->k1 + BK - 1) / BK;
  p.taps = a->conv_taps;
  p.bias = a->bias;
  p.row_bias = a->row_bias;
  p.row_div = a->row_div > 0 ? a->row_div : 1;
  p.row_mod = a->row_mod > 0 ? a->row_mod : 1;
  p.residual = static_cast<const __half*>(a->residual);
  p.ldr = a->ldr;
  p.geglu = a->geglu;
  p.seg_cols = a->seg_cols;
  p.trans_rows = a->trans_rows > 0 ? a->trans_rows : 1;
  p.trans_ld = a->trans_ld;
  p.trans_head_d = 0;
  p.trans_head_dp = 0;
  if (a->trans_head_dp > 0) {
    MDK_REQUIRE(a->trans_head_d > 0 && a->trans_head_dp >= a->trans_head_d && a->seg_cols > 0 &&
                    a->seg_cols % a->trans_head_d == 0,
                "mdk_gemm_f16: bad trans_head_d=%d trans_head_dp=%d", a->trans_head_d, a->trans_head_dp);
    p.trans_head_d = a->trans_head_d;
    p.trans_head_dp = a->trans_head_dp;
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("MDK_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.dbg = dbg;
  }
  int nseg = 1;
  const int ncols_out = a->geglu ? a->n / 2 : a->n;
  if (a->seg_cols > 0) {
    MDK_REQUIRE(!a->geglu, "mdk_gemm_f16: seg_cols with geglu is not supported");
    MDK_REQUIRE(ncols_out % a->seg_cols == 0 && ncols_out / a->seg_cols <= 3,
                "mdk_gemm_f16: n=%d is not 1..3 segments of %d", ncols_out, a->seg_cols);
    nseg = ncols_out / a->seg_cols;
  }
  for (int s = 0; s < 3; ++s) {
    p.out[s] = static_cast<__half*>(a->out[s]);
    p.ldo[s] = a->ldo[s];
    p.out_trans[s] = a->out_trans[s];
    if (s < nseg) {
      MDK_REQUIRE(a->out[s] != nullptr, "mdk_gemm_f16: out[%d] is NULL", s);
      if (a->out_trans[s])
        MDK_REQUIRE(a->seg_cols > 0 && a->trans_ld > 0,
                    "mdk_gemm_f16: transposed output needs seg_cols and trans_ld");
      else
        MDK_REQUIRE(a->ldo[s] % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out[s]) & 15) == 0,
                    "mdk_gemm_f16: out[%d] must be 16-byte aligned with ldo %% 8 == 0", s);
    }
  }
  MDK_REQUIRE((reinterpret_cast<uintptr_t>(a->bias) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->row_bias) & 15) == 0,
              "mdk_gemm_f16: bias / row_bias must be 16-byte aligned");
  if (a->residual)
    MDK_REQUIRE(a->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0,
                "mdk_gemm_f16: residual must be 16-byte aligned with ldr %% 8 == 0");

  // CTA pairs (256-row tiles, tcgen05.mma.cta_group::2) for the main-loop-bound problems: measured
  // (PERF.md) +18..21 % on the K >= 2880 convolutions, but -5..-20 % on the epilogue-bound K <= 640
  // linears (both CTAs' epilogues gate the pair's next tile).  MDK_GEMM_CG=1 / =2 forces one kernel.
  int cg_env;
  {
    const char* e = getenv("MDK_GEMM_CG");   // read per call: tests switch kernels in-process
    cg_env = e ? atoi(e) : 0;
  }
  const int m_tiles_est = (a->m + BM - 1) / BM;   // (conv tiles are also 128 pixels)
  const long long num_kb_est = static_cast<long long>(a->conv_taps) * ((a->k0 + BK - 1) / BK + (a->k1 + BK - 1) / BK);
  int cg = (num_kb_est >= 16) ? 2 : 1;
  if (cg_env == 1 || cg_env == 2) cg = cg_env;
  if (m_tiles_est < 2) cg = 1;
  int bn;
  if (a->geglu) {
    bn = 256;
    MDK_REQUIRE(a->n % 256 == 0, "mdk_gemm_f16: geglu needs n %% 256 == 0 (n=%d)", a->n);
  } else {
    bn = pick_bn(a->n, a->seg_cols, nseg, (m_tiles_est + cg - 1) / cg, ctx->num_sms / cg);
    MDK_REQUIRE(bn > 0, "mdk_gemm_f16: no tile width divides seg_cols=%d", a->seg_cols);
  }
  p.n_tiles = (a->n + bn - 1) / bn;

  // ---- tensor maps ----
  if (a->conv_taps == 1) {
    MDK_REQUIRE(a->lda0 % 8 == 0 && a->lda0 >= a->k0, "mdk_gemm_f16: bad lda0=%lld",
                (long long)a->lda0);
    p.m_tiles = (a->m + BM - 1) / BM;
    uint64_t dims[2] = {static_cast<uint64_t>(a->k0), static_cast<uint64_t>(a->m)};
    uint64_t str[2] = {0, static_cast<uint64_t>(a->lda0) * 2};
    uint32_t box[2] = {BK, BM};
    if (encode_tmap_f16(&p.tmA0, a->a0, 2, dims, str, box)) return -1;
    if (a->k1 > 0) {
      MDK_REQUIRE(a->a1 && a->lda1 % 8 == 0 && a->lda1 >= a->k1, "mdk_gemm_f16: bad a1/lda1");
      uint64_t dims1[2] = {static_cast<uint64_t>(a->k1), static_cast<uint64_t>(a->m)};
      uint64_t str1[2] = {0, static_cast<uint64_t>(a->lda1) * 2};
      if (encode_tmap_f16(&p.tmA1, a->a1, 2, dims1, str1, box)) return -1;
    }
  } else {
    MDK_REQUIRE(a->nimg > 0 && a->h > 0 && a->w > 0 &&
                    static_cast<long long>(a->nimg) * a->h * a->w == a->m,
                "mdk_gemm_f16: conv needs m == nimg*h*w");
    p.H = a->h;
    p.W = a->w;
    p.nimg = a->nimg;
    p.bw = pow2_div(a->w, 16);
    p.bh = pow2_div(a->h, BM / p.bw);
    p.bn = BM / (p.bw * p.bh);
    p.tiles_w = a->w / p.bw;
    p.tiles_h = a->h / p.bh;
    const int tiles_n = (a->nimg + p.bn - 1) / p.bn;
    p.m_tiles = p.tiles_w * p.tiles_h * tiles_n;
    uint32_t box[4] = {BK, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh),
                       static_cast<uint32_t>(p.bn)};
    {
      const uint64_t c = static_cast<uint64_t>(a->k0);
      uint64_t dims[4] = {c, static_cast<uint64_t>(a->w), static_cast<uint64_t>(a->h),
                          static_cast<uint64_t>(a->nimg)};
      uint64_t str[4] = {0, c * 2, c * 2 * a->w, c * 2 * a->w * a->h};
      if (encode_tmap_f16(&p.tmA0, a->a0, 4, dims, str, box)) return -1;
    }
    if (a->k1 > 0) {
      MDK_REQUIRE(a->a1 != nullptr, "mdk_gemm_f16: a1 is NULL with k1 > 0");
      const uint64_t c = static_cast<uint64_t>(a->k1);
      uint64_t dims[4] = {c, static_cast<uint64_t>(a->w), static_cast<uint64_t>(a->h),
                          static_cast<uint64_t>(a->nimg)};
      uint64_t str[4] = {0, c * 2, c * 2 * a->w, c * 2 * a->w * a->h};
      if (encode_tmap_f16(&p.tmA1, a->a1, 4, dims, str, box)) return -1;
    }
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(a->n)};
    uint64_t str[2] = {0, static_cast<uint64_t>(a->ldb) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(bn / cg)};   // pair mode: each CTA loads half of the B tile
    if (encode_tmap_f16(&p.tmB, a->b, 2, dims, str, box)) return -1;
  }

  {
    static int tma_store = -1;
    if (tma_store < 0) {
      const char* e = getenv("MDK_GEMM_TMA_STORE");
      tma_store = e ? atoi(e) : 1;   // default: TMA-store epilogue (MDK_GEMM_TMA_STORE=0 selects st.global)
    }
    p.tma_store = tma_store;
  }
  if (p.tma_store) {
    const int segw = a->seg_cols > 0 ? a->seg_cols : ncols_out;
    uint32_t obox4[4] = {32, 1, 1, 1};
    if (a->conv_taps == 9) {
      p.sbw = p.bw;
      p.sbh = p.bh < 32 / p.bw ? p.bh : 32 / p.bw;
      p.sbn = 32 / (p.sbw * p.sbh);
      obox4[1] = static_cast<uint32_t>(p.sbw);
      obox4[2] = static_cast<uint32_t>(p.sbh);
      obox4[3] = static_cast<uint32_t>(p.sbn);
    }
    for (int s = 0; s < nseg; ++s) {
      if (a->out_trans[s]) continue;
      const uint64_t ld = static_cast<uint64_t>(a->ldo[s]);
      if (a->conv_taps == 9) {
        uint64_t dims[4] = {static_cast<uint64_t>(segw), static_cast<uint64_t>(a->w),
                            static_cast<uint64_t>(a->h), static_cast<uint64_t>(a->nimg)};
        uint64_t str[4] = {0, ld * 2, ld * 2 * a->w, ld * 2 * a->w * a->h};
        if (encode_tmap_f16(&p.tmO[s], a->out[s], 4, dims, str, obox4, 64)) return -1;
      } else {
        uint64_t dims[2] = {static_cast<uint64_t>(segw), static_cast<uint64_t>(a->m)};
        uint64_t str[2] = {0, ld * 2};
        uint32_t obox[2] = {32, 32};
        if (encode_tmap_f16(&p.tmO[s], a->out[s], 2, dims, str, obox, 64)) return -1;
      }
    }
  }

  p.mp_tiles = (p.m_tiles + cg - 1) / cg;
  {
    // B-stationary kernel: plain single-source GEMM, K <= 320, single CTA, and enough M tiles per SM that the panel
    // (re)load amortises (level 0: 2 304 M tiles).  MDK_GEMM_BS=1 switches it on (2: for every eligible M).
    const char* e = getenv("MDK_GEMM_BS");   // read per call: tests switch kernels in-process
    const int bs = e ? atoi(e) : 0;   // measured (profiles/r02_gemm_bstationary.log): no gain — 0.098 vs 0.090 ms for
                                      // M = 294 912, N = K = 320; the step is unchanged — so it is off by default
    if (bs && cg == 1 && a->conv_taps == 1 && a->k1 == 0 && !a->geglu && K <= BS_KB * BK &&
        (bs == 2 || p.mp_tiles >= 4 * ctx->num_sms)) {   // (= 2: regardless of M, for the parity tests)
      switch (bn) {
        case 192: return launch_gemm<192, 1, true>(ctx, p, stream);
        case 160: return launch_gemm<160, 1, true>(ctx, p, stream);
        case 128: return launch_gemm<128, 1, true>(ctx, p, stream);
        default: break;
      }
    }
  }
  if (cg == 2) {
    switch (bn) {
      case 256: return launch_gemm<256, 2>(ctx, p, stream);
      case 192: return launch_gemm<192, 2>(ctx, p, stream);
      case 160: return launch_gemm<160, 2>(ctx, p, stream);
      case 128: return launch_gemm<128, 2>(ctx, p, stream);
      case 64: return launch_gemm<64, 2>(ctx, p, stream);
      case 32: return launch_gemm<32, 2>(ctx, p, stream);
    }
  } else {
    switch (bn) {
      case 256: return launch_gemm<256, 1>(ctx, p, stream);
      case 192: return launch_gemm<192, 1>(ctx, p, stream);
      case 160: return launch_gemm<160, 1>(ctx, p, stream);
      case 128: return launch_gemm<128, 1>(ctx, p, stream);
      case 64: return launch_gemm<64, 1>(ctx, p, stream);
      case 32: return launch_gemm<32, 1>(ctx, p, stream);
    }
  }
  return set_error("mdk_gemm_f16: internal: bad tile width %d", bn);
}
