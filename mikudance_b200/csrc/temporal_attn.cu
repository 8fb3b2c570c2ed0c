// Temporal attention of the motion module: softmax(Q K^T / sqrt(d)) V over the frame axis at a
// fixed pixel and head; sequence length f <= 32.  0.1 % of the step's FLOPs — the kernel is bound
// by HBM traffic (read Q,K,V, write O once), so it is a SIMT kernel with 16-byte vector IO:
//   * one (batch, pixel, head) problem per QPW-lane group (QPW = pow2 >= f_q), lane = query frame
//   * K and V rows of the problem staged in shared memory (broadcast reads), scores in registers
//   * the reference's "(b f) d c -> (b d) f c" transposes become strided addressing
//   * PE is added to the query only: q += pe_q[frame]  with pe_q = pe @ Wq^T (fp32)
//
// Algorithmic bytes per launch: 2 * nb*npix*C * (2*f_q + 2*f_kv) (Q + O over f_q frames, K + V
// over f_kv frames).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"
#include "../../include/mdk.h"

namespace mdk {

struct TattnParams {
  const __half* q;
  const __half* kv;
  const float* pe_q;
  __half* out;
  long long q_ld, kv_ld, out_ld;
  int q_off, k_off, v_off;
  int nb, f_q, f_kv, f_kv_rank, f_q_offset, npix, heads, d;
  int qpw;      // lanes per problem (pow2 >= f_q)
  int dp;       // padded row length in smem (halves)
  long long nprob;
  float scale_log2;
};

template <int FKV_MAX>
__global__ void temporal_attn_kernel(const TattnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int ppw = 32 / p.qpw;  // problems per warp
  const int dv = p.d >> 3;     // 16-byte vectors per head row
  const int per_prob = 2 * p.f_kv * p.dp;  // halves (K then V)
  __half* wsm = reinterpret_cast<__half*>(smem_raw) + static_cast<long long>(warp) * ppw * per_prob;

  const long long groups = (p.nprob + ppw - 1) / ppw;
  for (long long grp = static_cast<long long>(blockIdx.x) * warps + warp; grp < groups;
       grp += static_cast<long long>(gridDim.x) * warps) {
    const long long pid0 = grp * ppw;
    // ---- stage K and V of the warp's problems ----
    const int nvec = ppw * p.f_kv * dv * 2;
    for (int idx = lane; idx < nvec; idx += 32) {
      int t = idx;
      const int v = t % dv;
      t /= dv;
      const int j = t % p.f_kv;
      t /= p.f_kv;
      const int which = t & 1;  // 0 = K, 1 = V
      const int pr = t >> 1;
      const long long pid = pid0 + pr;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (pid < p.nprob) {
        const int h = static_cast<int>(pid % p.heads);
        const long long bp = pid / p.heads;
        const int px = static_cast<int>(bp % p.npix);
        const int b = static_cast<int>(bp / p.npix);
        const int g = j / p.f_kv_rank, fl = j % p.f_kv_rank;
        const long long row =
            ((static_cast<long long>(g) * p.nb + b) * p.f_kv_rank + fl) * p.npix + px;
        val = *reinterpret_cast<const uint4*>(p.kv + row * p.kv_ld + (which ? p.v_off : p.k_off) +
                                              h * p.d + v * 8);
      }
      *reinterpret_cast<uint4*>(wsm + pr * per_prob + which * p.f_kv * p.dp + j * p.dp + v * 8) = val;
    }
    __syncwarp();

    const int pr = lane / p.qpw;
    const int i = lane % p.qpw;
    const long long pid = pid0 + pr;
    const bool active = (i < p.f_q) && (pid < p.nprob);
    if (active) {
      const int h = static_cast<int>(pid % p.heads);
      const long long bp = pid / p.heads;
      const int px = static_cast<int>(bp % p.npix);
      const int b = static_cast<int>(bp / p.npix);
      const long long qrow = (static_cast<long long>(b) * p.f_q + i) * p.npix + px;
      const __half* qptr = p.q + qrow * p.q_ld + p.q_off + h * p.d;
      const float* peptr =
          p.pe_q ? p.pe_q + static_cast<long long>(p.f_q_offset + i) * (p.heads * p.d) + h * p.d
                 : nullptr;
      const __half* ks = wsm + pr * per_prob;
      const __half* vs = ks + p.f_kv * p.dp;

      float s[FKV_MAX];
#pragma unroll
      for (int j = 0; j < FKV_MAX; ++j) s[j] = 0.f;
      for (int v = 0; v < dv; ++v) {
        float qv[8];
        {
          uint4 raw = *reinterpret_cast<const uint4*>(qptr + v * 8);
          const __half2* hq = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __half22float2(hq[e]);
            qv[2 * e] = f.x;
            qv[2 * e + 1] = f.y;
          }
          if (peptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) qv[e] += peptr[v * 8 + e];
          }
        }
#pragma unroll
        for (int j = 0; j < FKV_MAX; ++j) {
          if (j < p.f_kv) {
            uint4 raw = *reinterpret_cast<const uint4*>(ks + j * p.dp + v * 8);
            const __half2* hk = reinterpret_cast<const __half2*>(&raw);
            float a = s[j];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __half22float2(hk[e]);
              a = fmaf(qv[2 * e], f.x, a);
              a = fmaf(qv[2 * e + 1], f.y, a);
            }
            s[j] = a;
          }
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < FKV_MAX; ++j)
        if (j < p.f_kv) mx = fmaxf(mx, s[j]);
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < FKV_MAX; ++j) {
        if (j < p.f_kv) {
          s[j] = exp2f((s[j] - mx) * p.scale_log2);
          l += s[j];
        }
      }
      const float inv = 1.0f / l;
      __half* optr = p.out + qrow * p.out_ld + h * p.d;
      for (int v = 0; v < dv; ++v) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
        for (int j = 0; j < FKV_MAX; ++j) {
          if (j < p.f_kv) {
            uint4 raw = *reinterpret_cast<const uint4*>(vs + j * p.dp + v * 8);
            const __half2* hv = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __half22float2(hv[e]);
              o[2 * e] = fmaf(s[j], f.x, o[2 * e]);
              o[2 * e + 1] = fmaf(s[j], f.y, o[2 * e + 1]);
            }
          }
        }
        uint4 pk;
        pk.x = pack_half2(o[0] * inv, o[1] * inv);
        pk.y = pack_half2(o[2] * inv, o[3] * inv);
        pk.z = pack_half2(o[4] * inv, o[5] * inv);
        pk.w = pack_half2(o[6] * inv, o[7] * inv);
        *reinterpret_cast<uint4*>(optr + v * 8) = pk;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Tiled variant for the single-GPU layout (fused q|k|v rows [M, 3C], f_q == f_kv, PE already folded
// into Q by the projection GEMM's row bias): a CTA owns PG
// consecutive pixels of one batch entry, stages their f x 3C rows with fully coalesced 16-byte loads
// (consecutive pixels of one frame are contiguous in memory), computes every (pixel, head) problem
// from shared memory, overwrites the Q slots with O and writes the outputs back coalesced.
// ------------------------------------------------------------------------------------------------
struct TattnTileParams {
  const __half* qkv;
  const float* pe_q;
  __half* out;
  long long ld, out_ld;
  int nb, f, npix, heads, d, C;
  int hg;        // heads per CTA (one warp each); the CTA stages only their q|k|v column segments
  int seg;       // hg * d: columns of one segment
  int rs;        // smem row stride in halves (3*seg + 8)
  float scale_log2;
};

__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                             uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];\n"
               : "=r"(r0), "=r"(r1)
               : "r"(addr));
}

// One warp per (pixel, head) problem.  The tiny GEMMs run on the legacy mma.sync path on purpose:
// a 16 x 16 x 40 problem cannot fill a tcgen05 128-row tile, the kernel is HBM-bound, and
// m16n8k16 fragments can be fed straight from the staged rows (row stride 3C+8 halves makes every
// fragment load bank-conflict free).  MT = number of 16-row query tiles (frames padded to 16*MT).
// The (pixel, head) problems of one staged item: warp w < hg computes head w of the group from the rows at `sm`
// (q | k | v segments per frame row, row stride p.rs halves) and overwrites the Q slots with O.
template <int MT>
__device__ __forceinline__ void tattn_compute_item(const TattnTileParams& p, __half* sm) {
  constexpr int FP = 16 * MT;
  (void)FP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ksteps = (p.d + 15) >> 4;
  const bool half_last = (p.d & 15) != 0;  // d % 16 == 8: upper half of the last K step is padding
  const int ntile_o = p.d >> 3;
  const uint32_t rsb = static_cast<uint32_t>(p.rs) * 2u;  // row stride in bytes
  if (warp < p.hg) {
    const uint32_t qb = smem_u32(sm) + static_cast<uint32_t>(warp * p.d) * 2u;
    const uint32_t kb = qb + static_cast<uint32_t>(p.seg) * 2u;
    const uint32_t vb = qb + static_cast<uint32_t>(p.seg) * 4u;
    // ---- S = Q K^T ----
    float sacc[MT][2 * MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[mt][nt][e] = 0.f;
    for (int ks = 0; ks < ksteps; ++ks) {
      const bool pad_hi = half_last && ks == ksteps - 1;
      const uint32_t col = static_cast<uint32_t>(ks * 16 + 2 * t) * 2u;
      uint32_t bfr[2 * MT][2];
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt) {
        const uint32_t ra = kb + static_cast<uint32_t>(nt * 8 + g) * rsb + col;
        bfr[nt][0] = lds32(ra);
        bfr[nt][1] = pad_hi ? 0u : lds32(ra + 16);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t a[4];
        const uint32_t r0 = qb + static_cast<uint32_t>(mt * 16 + g) * rsb + col;
        a[0] = lds32(r0);
        a[1] = lds32(r0 + 8 * rsb);
        a[2] = pad_hi ? 0u : lds32(r0 + 16);
        a[3] = pad_hi ? 0u : lds32(r0 + 8 * rsb + 16);
#pragma unroll
        for (int nt = 0; nt < 2 * MT; ++nt) mma_m16n8k16(sacc[mt][nt], a, bfr[nt][0], bfr[nt][1]);
      }
    }
    // ---- softmax over the keys (columns), rows g and g+8 of each query tile ----
    uint32_t pfr[MT][MT][4];  // P as the A operand of the P V MMAs: [query tile][key k-step]
    float inv_l[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool valid = nt * 8 + 2 * t + e < p.f;
          if (!valid) {
            sacc[mt][nt][e] = -INFINITY;
            sacc[mt][nt][2 + e] = -INFINITY;
          }
          mx0 = fmaxf(mx0, sacc[mt][nt][e]);
          mx1 = fmaxf(mx1, sacc[mt][nt][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p0 = exp2f((sacc[mt][nt][e] - mx0) * p.scale_log2);
          const float p1 = exp2f((sacc[mt][nt][2 + e] - mx1) * p.scale_log2);
          sacc[mt][nt][e] = p0;
          sacc[mt][nt][2 + e] = p1;
          l0 += p0;
          l1 += p1;
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      inv_l[mt][0] = 1.0f / l0;
      inv_l[mt][1] = 1.0f / l1;
#pragma unroll
      for (int kk = 0; kk < MT; ++kk) {
        pfr[mt][kk][0] = pack_half2(sacc[mt][2 * kk][0], sacc[mt][2 * kk][1]);
        pfr[mt][kk][1] = pack_half2(sacc[mt][2 * kk][2], sacc[mt][2 * kk][3]);
        pfr[mt][kk][2] = pack_half2(sacc[mt][2 * kk + 1][0], sacc[mt][2 * kk + 1][1]);
        pfr[mt][kk][3] = pack_half2(sacc[mt][2 * kk + 1][2], sacc[mt][2 * kk + 1][3]);
      }
    }
    // ---- O = P V, written over the (dead) query slot of this head ----
    __syncwarp();
    for (int nt = 0; nt < ntile_o; ++nt) {
      float oacc[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[mt][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < MT; ++kk) {
        uint32_t b0, b1;
        // lanes 0-15 address the 16 key rows of this k-step (8 channels starting at nt*8)
        ldmatrix_x2_trans(vb + static_cast<uint32_t>(kk * 16 + (lane & 15)) * rsb +
                              static_cast<uint32_t>(nt * 8) * 2u,
                          b0, b1);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma_m16n8k16(oacc[mt], pfr[mt][kk], b0, b1);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t o0 = qb + static_cast<uint32_t>(mt * 16 + g) * rsb +
                            static_cast<uint32_t>(nt * 8 + 2 * t) * 2u;
        sts32(o0, pack_half2(oacc[mt][0] * inv_l[mt][0], oacc[mt][1] * inv_l[mt][0]));
        sts32(o0 + 8 * rsb, pack_half2(oacc[mt][2] * inv_l[mt][1], oacc[mt][3] * inv_l[mt][1]));
      }
    }
  }
}

template <int MT>
__global__ void __launch_bounds__(256) temporal_attn_tile_kernel(const TattnTileParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __half* sm = reinterpret_cast<__half*>(smem_raw);
  constexpr int FP = 16 * MT;
  // CTA = (batch entry, pixel, head group): small CTAs (31 KB of smem at any level) so that 6-7 of
  // them overlap their load / compute / store phases on one SM
  const int ngroups = p.heads / p.hg;
  const int hgi = blockIdx.x % ngroups;
  const long long bp = blockIdx.x / ngroups;
  const int px = static_cast<int>(bp % p.npix);
  const int b = static_cast<int>(bp / p.npix);
  const int segv = p.seg >> 3;           // 16-byte vectors per segment
  const int col0 = hgi * p.seg;          // first channel of this head group
  // ---- load the q | k | v segments of every frame (zero rows for padded frames) ----
  for (int i = threadIdx.x; i < FP * 3 * segv; i += blockDim.x) {
    const int v = i % segv;
    const int which = (i / segv) % 3;
    const int j = i / (3 * segv);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (j < p.f)
      val = *reinterpret_cast<const uint4*>(
          p.qkv + ((static_cast<long long>(b) * p.f + j) * p.npix + px) * p.ld + which * p.C + col0 + v * 8);
    *reinterpret_cast<uint4*>(sm + static_cast<long long>(j) * p.rs + which * p.seg + v * 8) = val;
  }
  __syncthreads();
  tattn_compute_item<MT>(p, sm);
  __syncthreads();
  // ---- store the O segment of every frame (first `seg` columns of every staged row) ----
  for (int i = threadIdx.x; i < p.f * segv; i += blockDim.x) {
    const int v = i % segv, j = i / segv;
    *reinterpret_cast<uint4*>(p.out + ((static_cast<long long>(b) * p.f + j) * p.npix + px) * p.out_ld +
                              col0 + v * 8) =
        *reinterpret_cast<const uint4*>(sm + static_cast<long long>(j) * p.rs + v * 8);
  }
}

// Pipelined variant (default): a CTA walks a strided list of (batch entry, pixel, head group) items with TWO
// staging buffers — the 16-byte cp.async loads of item i+1 are in flight while item i is computed and stored — so
// every resident CTA always has ~30 KB of loads outstanding (the one-item-per-CTA kernel above alternates
// load / compute / store phases and measured 2.3 TB/s of the 6.5 TB/s copy bandwidth).  Rows of padded frames
// (f < 16 MT) are zeroed once; loads never touch them (only their dead Q slots are overwritten by O).
template <int MT>
__global__ void __launch_bounds__(256) temporal_attn_pipe_kernel(const TattnTileParams p, long long nitems) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int FP = 16 * MT;
  const int stage_halves = FP * p.rs;
  __half* st0 = reinterpret_cast<__half*>(smem_raw);
  const int ngroups = p.heads / p.hg;
  const int segv = p.seg >> 3;
  for (int i = threadIdx.x; i < 2 * stage_halves / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(st0)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  auto issue_loads = [&](long long item, __half* sm) {
    const int hgi = static_cast<int>(item % ngroups);
    const long long bp = item / ngroups;
    const int px = static_cast<int>(bp % p.npix);
    const int b = static_cast<int>(bp / p.npix);
    const int col0 = hgi * p.seg;
    for (int i = threadIdx.x; i < p.f * 3 * segv; i += blockDim.x) {
      const int v = i % segv;
      const int which = (i / segv) % 3;
      const int j = i / (3 * segv);
      const __half* src = p.qkv + ((static_cast<long long>(b) * p.f + j) * p.npix + px) * p.ld + which * p.C + col0 + v * 8;
      const uint32_t dst = smem_u32(sm + static_cast<long long>(j) * p.rs + which * p.seg + v * 8);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  long long item = blockIdx.x;
  if (item < nitems) issue_loads(item, st0);
  int s = 0;
  for (; item < nitems; item += gridDim.x, s ^= 1) {
    __half* sm = st0 + static_cast<long long>(s) * stage_halves;
    const long long next = item + gridDim.x;
    if (next < nitems) {
      issue_loads(next, st0 + static_cast<long long>(s ^ 1) * stage_halves);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    tattn_compute_item<MT>(p, sm);
    __syncthreads();
    const int hgi = static_cast<int>(item % ngroups);
    const long long bp = item / ngroups;
    const int px = static_cast<int>(bp % p.npix);
    const int b = static_cast<int>(bp / p.npix);
    const int col0 = hgi * p.seg;
    for (int i = threadIdx.x; i < p.f * segv; i += blockDim.x) {
      const int v = i % segv, j = i / segv;
      *reinterpret_cast<uint4*>(p.out + ((static_cast<long long>(b) * p.f + j) * p.npix + px) * p.out_ld + col0 + v * 8) =
          *reinterpret_cast<const uint4*>(sm + static_cast<long long>(j) * p.rs + v * 8);
    }
    __syncthreads();   // the buffer is free for the loads of item i+2
  }
}

}  // namespace mdk

extern "C" int mdk_temporal_attn_f16(mdk_ctx* ctx, const mdk_tattn_args* a, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && a && a->q && a->kv && a->out, "mdk_temporal_attn_f16: null argument");
  MDK_REQUIRE(a->d % 8 == 0 && a->d > 0, "mdk_temporal_attn_f16: d=%d must be a multiple of 8", a->d);
  MDK_REQUIRE(a->f_q >= 1 && a->f_q <= 32 && a->f_kv >= 1 && a->f_kv <= 32,
              "mdk_temporal_attn_f16: f_q=%d f_kv=%d must be in 1..32 (PE table max_len 32)", a->f_q,
              a->f_kv);
  const int f_kv_rank = a->f_kv_rank > 0 ? a->f_kv_rank : a->f_kv;
  MDK_REQUIRE(a->f_kv % f_kv_rank == 0, "mdk_temporal_attn_f16: f_kv %% f_kv_rank != 0");
  MDK_REQUIRE(a->q_ld % 8 == 0 && a->kv_ld % 8 == 0 && a->out_ld % 8 == 0 && a->q_off % 8 == 0 &&
                  a->k_off % 8 == 0 && a->v_off % 8 == 0,
              "mdk_temporal_attn_f16: leading dimensions / offsets must be multiples of 8");
  // ---- tiled fast path: fused q|k|v rows, all frames local ----
  {
    const int C = a->heads * a->d;
    const bool fused = a->q == a->kv && a->q_off == 0 && a->k_off == C && a->v_off == 2 * C &&
                       a->q_ld == a->kv_ld && a->q_ld >= 3 * C && a->f_q == a->f_kv &&
                       f_kv_rank == a->f_kv && a->f_q_offset == 0;
    // heads per CTA: keep one q/k/v segment around 640 B (good coalescing, ~31 KB of smem per CTA)
    int hg = a->heads;
    while (hg > 1 && hg % 2 == 0 && (hg / 2) * a->d >= 320) hg /= 2;
    if (hg > 8) hg = 8;
    while (a->heads % hg != 0) --hg;
    const int seg = hg * a->d;
    const int rs = 3 * seg + 8;
    const int fp = a->f_q <= 16 ? 16 : 32;
    const size_t smem = static_cast<size_t>(fp) * rs * sizeof(__half);
    if (fused && a->pe_q == nullptr && smem <= 200 * 1024) {
      TattnTileParams t;
      t.qkv = static_cast<const __half*>(a->q);
      t.pe_q = nullptr;
      t.out = static_cast<__half*>(a->out);
      t.ld = a->q_ld;
      t.out_ld = a->out_ld;
      t.nb = a->nb;
      t.f = a->f_q;
      t.npix = a->npix;
      t.heads = a->heads;
      t.d = a->d;
      t.C = C;
      t.hg = hg;
      t.seg = seg;
      t.rs = rs;
      t.scale_log2 = a->scale * 1.4426950408889634f;
      static unsigned long long tile_attr_mask = 0;
      const bool tile_attr = ((tile_attr_mask >> (ctx->device & 63)) & 1ull) != 0;
      if (!tile_attr) {
        MDK_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_tile_kernel<1>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MDK_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_tile_kernel<2>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        tile_attr_mask |= 1ull << (ctx->device & 63);
      }
      const long long grid = static_cast<long long>(a->nb) * a->npix * (a->heads / hg);
      MDK_REQUIRE(grid < (1ll << 31), "mdk_temporal_attn_f16: grid too large");
      const int threads = 32 * hg < 64 ? 64 : 32 * hg;
      static int pipe = -1;
      if (pipe < 0) {
        const char* e = getenv("MDK_TATTN_PIPE");
        pipe = e ? atoi(e) : 1;
      }
      if (pipe && 2 * smem <= 200 * 1024) {
        MDK_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_pipe_kernel<1>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MDK_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_pipe_kernel<2>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int per_sm = static_cast<int>((220 * 1024) / (2 * smem + 1024));
        if (per_sm > 2048 / threads) per_sm = 2048 / threads;
        if (per_sm < 1) per_sm = 1;
        long long ctas = static_cast<long long>(ctx->num_sms) * per_sm;
        if (ctas > grid) ctas = grid;
        if (a->f_q <= 16)
          MDK_CHECK_CUDA(launch_pdl(temporal_attn_pipe_kernel<1>, dim3(static_cast<unsigned>(ctas)), dim3(threads), 2 * smem,
                                    stream, t, grid));
        else
          MDK_CHECK_CUDA(launch_pdl(temporal_attn_pipe_kernel<2>, dim3(static_cast<unsigned>(ctas)), dim3(threads), 2 * smem,
                                    stream, t, grid));
      } else if (a->f_q <= 16) {
        temporal_attn_tile_kernel<1><<<static_cast<unsigned>(grid), threads, smem, stream>>>(t);
      } else {
        temporal_attn_tile_kernel<2><<<static_cast<unsigned>(grid), threads, smem, stream>>>(t);
      }
      count_launch();
      MDK_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  TattnParams p;
  p.q = static_cast<const __half*>(a->q);
  p.kv = static_cast<const __half*>(a->kv);
  p.pe_q = a->pe_q;
  p.out = static_cast<__half*>(a->out);
  p.q_ld = a->q_ld;
  p.kv_ld = a->kv_ld;
  p.out_ld = a->out_ld;
  p.q_off = a->q_off;
  p.k_off = a->k_off;
  p.v_off = a->v_off;
  p.nb = a->nb;
  p.f_q = a->f_q;
  p.f_kv = a->f_kv;
  p.f_kv_rank = f_kv_rank;
  p.f_q_offset = a->f_q_offset;
  p.npix = a->npix;
  p.heads = a->heads;
  p.d = a->d;
  int qpw = 1;
  while (qpw < a->f_q) qpw *= 2;
  p.qpw = qpw;
  p.dp = a->d + 8;
  p.nprob = static_cast<long long>(a->nb) * a->npix * a->heads;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  const int ppw = 32 / qpw;
  const size_t per_warp = static_cast<size_t>(ppw) * 2 * a->f_kv * p.dp * sizeof(__half);
  int warps = 8;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps /= 2;
  const size_t smem = per_warp * warps;
  MDK_REQUIRE(smem <= 200 * 1024, "mdk_temporal_attn_f16: problem too large for shared memory");
  const long long groups = (p.nprob + ppw - 1) / ppw;
  long long ctas = (groups + warps - 1) / warps;
  const long long cap = static_cast<long long>(ctx->num_sms) * 8;
  if (ctas > cap) ctas = cap;
  static unsigned long long attr_set_mask = 0;   // bit d: attribute set on device d (it is a per-device property)
  const bool attr_set = ((attr_set_mask >> (ctx->device & 63)) & 1ull) != 0;
  if (!attr_set) {
    MDK_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_kernel<16>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MDK_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_kernel<32>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set_mask |= 1ull << (ctx->device & 63);
  }
  if (a->f_kv <= 16)
    temporal_attn_kernel<16><<<static_cast<unsigned>(ctas), warps * 32, smem, stream>>>(p);
  else
    temporal_attn_kernel<32><<<static_cast<unsigned>(ctas), warps * 32, smem, stream>>>(p);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
