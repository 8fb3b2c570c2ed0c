// GroupNorm (per image, NHWC, one or two channel-concatenated sources, optional SiLU) and
// LayerNorm (+ optional reference-bank add) — HBM-bound kernels, fp32 math, 16-byte vector IO.
//
// Algorithmic bytes: GroupNorm reads the input twice (stats pass + apply pass) and writes once:
// 3 * 2 * nimg*hw*C bytes; LayerNorm reads once, writes once (twice with the bank output).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"
#include "../../include/mdk.h"

namespace mdk {

constexpr int GN_MAX_CHUNKS = 64;

struct GnParams {
  const __half* x0;
  const __half* x1;
  int c0, c1, C;
  int nimg, hw, groups, cpg;
  float eps;
  const __half* gamma;
  const __half* beta;
  int silu;
  __half* out;
  float* ws;  // [nimg, GN_MAX_CHUNKS, groups, 2] per-CTA partial (sum, sumsq)
  int pix_per_cta;
  int nchunks;
  int stage_pix;       // bulk kernels: pixels per shared-memory stage
  int out_chunk_pix;   // > 0: exchange layout [chunk, nimg_total, out_chunk_pix, C] (see mdk.h)
  int nimg_total;      // images of the whole call (the launch may cover a sub-range starting at img0)
  int img0;
};

__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f = __half22float2(h[e]);
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  uint4 pk;
  pk.x = pack_half2(v[0], v[1]);
  pk.y = pack_half2(v[2], v[3]);
  pk.z = pack_half2(v[4], v[5]);
  pk.w = pack_half2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = pk;
}

// blockDim = (vx, vy): thread x owns channel vector cv = threadIdx.x (8 channels), thread y strides
// over the pixels of this CTA's slab.
// Deterministic (no atomics): per-thread partials -> fixed-order reduction over threadIdx.y ->
// fixed-order reduction over the channels of each group -> one partial per (image, chunk, group).
__global__ void gn_stats_kernel(const GnParams p) {
  extern __shared__ float s_red[];  // [vy][vx][16] then col[2][C]
  pdl_wait();
  pdl_trigger();
  const int img = blockIdx.y;
  const int cv = threadIdx.x;
  const int V = p.C >> 3;
  const int vx = blockDim.x, vy = blockDim.y;
  float* col = s_red + vy * vx * 16;
  const int tid = threadIdx.y * vx + threadIdx.x;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (cv < V) {
    const int ch = cv * 8;
    const bool from1 = ch >= p.c0;
    const __half* src = from1 ? p.x1 : p.x0;
    const int cs = from1 ? p.c1 : p.c0;
    const int coff = from1 ? ch - p.c0 : ch;
    const int pbeg = blockIdx.x * p.pix_per_cta;
    const int pend = min(p.hw, pbeg + p.pix_per_cta);
    for (int px = pbeg + threadIdx.y; px < pend; px += 4 * vy) {
      // four independent 16-byte loads in flight per thread
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = px + u * vy;
        raw[u] = (pp < pend) ? *reinterpret_cast<const uint4*>(
                                   src + (static_cast<long long>(img) * p.hw + pp) * cs + coff)
                             : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          s[2 * e] += f.x;
          q[2 * e] += f.x * f.x;
          s[2 * e + 1] += f.y;
          q[2 * e + 1] += f.y * f.y;
        }
      }
    }
  }
  float* mine = s_red + (threadIdx.y * vx + cv) * 16;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mine[e] = s[e];
    mine[8 + e] = q[e];
  }
  __syncthreads();
  if (threadIdx.y == 0 && cv < V) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = 0.f, b = 0.f;
      for (int y = 0; y < vy; ++y) {
        a += s_red[(y * vx + cv) * 16 + e];
        b += s_red[(y * vx + cv) * 16 + 8 + e];
      }
      col[cv * 8 + e] = a;
      col[p.C + cv * 8 + e] = b;
    }
  }
  __syncthreads();
  if (tid < p.groups) {
    float a = 0.f, b = 0.f;
    for (int c = tid * p.cpg; c < (tid + 1) * p.cpg; ++c) {
      a += col[c];
      b += col[p.C + c];
    }
    float* dst = p.ws + ((static_cast<long long>(img) * GN_MAX_CHUNKS + blockIdx.x) * p.groups + tid) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

__global__ void gn_apply_kernel(const GnParams p) {
  pdl_wait();
  pdl_trigger();
  const int img = blockIdx.y;
  const int cv = threadIdx.x;
  const int V = p.C >> 3;
  const int ch = cv * 8;
  const bool from1 = ch >= p.c0;
  const __half* src = from1 ? p.x1 : p.x0;
  const int cs = from1 ? p.c1 : p.c0;
  const int coff = from1 ? ch - p.c0 : ch;
  // mean / rstd of every group: combined once per CTA from the per-chunk partials (fixed order, fp64).  The
  // partials are fetched by ALL threads first (one coalesced wave of loads) and summed from shared memory: a
  // serial loop of nchunks dependent global loads per group thread kept the whole CTA at the barrier below for
  // ~10 us of a ~20 us lifetime.
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float s_part[GN_MAX_CHUNKS * 64 * 2];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  {
    const float* part = p.ws + static_cast<long long>(img) * GN_MAX_CHUNKS * p.groups * 2;
    const int n = p.nchunks * p.groups * 2;      // chunk-major, contiguous: [k][group][2]
    for (int i = tid; i < n; i += nthr) s_part[i] = part[i];
  }
  __syncthreads();
  if (tid < p.groups) {
    double sum = 0.0, sq = 0.0;
    for (int k = 0; k < p.nchunks; ++k) {
      sum += static_cast<double>(s_part[(k * p.groups + tid) * 2]);
      sq += static_cast<double>(s_part[(k * p.groups + tid) * 2 + 1]);
    }
    const double cnt = static_cast<double>(p.hw) * p.cpg;
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
  }
  __syncthreads();
  if (cv >= V) return;
  float scale[8], shift[8];
  {
    float gm[8], bt[8];
    load8(p.gamma + ch, gm);
    load8(p.beta + ch, bt);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (ch + e) / p.cpg;
      scale[e] = gm[e] * s_rstd[g];
      shift[e] = bt[e] - s_mean[g] * scale[e];
    }
  }
  const int pbeg = blockIdx.x * p.pix_per_cta;
  const int pend = min(p.hw, pbeg + p.pix_per_cta);
  const int stepy = blockDim.y;
  for (int px = pbeg + threadIdx.y; px < pend; px += 4 * stepy) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = px + u * stepy;
      if (pp < pend)
        raw[u] = *reinterpret_cast<const uint4*>(src + (static_cast<long long>(img) * p.hw + pp) * cs + coff);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = px + u * stepy;
      if (pp < pend) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
        float v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          v[2 * e] = f.x;
          v[2 * e + 1] = f.y;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = v[e] * scale[e] + shift[e];
          if (p.silu) y = __fdividef(y, 1.0f + __expf(-y));
          v[e] = y;
        }
        long long orow = static_cast<long long>(img) * p.hw + pp;
        if (p.out_chunk_pix > 0) {
          const int d = pp / p.out_chunk_pix;
          orow = (static_cast<long long>(d) * p.nimg_total + (p.img0 + img)) * p.out_chunk_pix + (pp - d * p.out_chunk_pix);
        }
        store8(p.out + orow * p.C + ch, v);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Bulk-copy variants (MDK_GN_BULK=1; measured time-neutral, so off by default): the CTA's pixel slab is contiguous in memory per source, so one elected thread
// streams it through a GN_STAGES-deep shared-memory ring with cp.async.bulk (1-D bulk copies signalled on
// mbarriers) and the threads read their channel vectors from shared memory.  Why: ncu on the register-staged
// kernels above (profiles/r02_ncu_norms_metrics.txt) — 54-56 registers x 480 threads = 2 CTAs per SM, 4 loads of
// 16 bytes in flight per thread = 61 KB per SM: 3.3 TB/s (stats) and 4.0 TB/s (apply) of the 6.5 TB/s copy
// bandwidth; 8 loads per thread need 88 registers and halve the occupancy (measured: slower).  With the ring the
// bytes in flight (GN_STAGES x ~16 KB per CTA) do not live in registers.  Result: parity-green, same time — bytes
// in flight were not the limit either; what helps the small levels is fewer, longer CTAs (MDK_GN_OVERSUB).
// ------------------------------------------------------------------------------------------------
constexpr int GN_STAGES = 4;
constexpr int GN_STAGE_BYTES = 16384;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// stage layout: [stage_pix x c0 halves | stage_pix x c1 halves]
struct GnRing {
  uint8_t* base;
  uint64_t* full;
  int pbeg, pend, nst;   // slab [pbeg, pend), number of stages of stage_pix pixels
};

__device__ __forceinline__ void gn_issue_stage(const GnParams& p, const GnRing& r, int img, int it) {
  const int px0 = r.pbeg + it * p.stage_pix;
  const int npx = min(p.stage_pix, r.pend - px0);
  const int slot = it % GN_STAGES;
  uint8_t* dst = r.base + slot * GN_STAGE_BYTES;
  const uint32_t b0 = static_cast<uint32_t>(npx) * p.c0 * 2u;
  const uint32_t b1 = static_cast<uint32_t>(npx) * p.c1 * 2u;
  mbar_expect_tx(&r.full[slot], b0 + b1);
  bulk_load_1d(smem_u32(dst), p.x0 + (static_cast<long long>(img) * p.hw + px0) * p.c0, b0, &r.full[slot]);
  if (p.c1 > 0)
    bulk_load_1d(smem_u32(dst) + static_cast<uint32_t>(p.stage_pix) * p.c0 * 2u,
                 p.x1 + (static_cast<long long>(img) * p.hw + px0) * p.c1, b1, &r.full[slot]);
}

__global__ void gn_stats_bulk_kernel(const GnParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  float* s_red = reinterpret_cast<float*>(gsm + GN_STAGES * GN_STAGE_BYTES);   // [vy][vx][16] then col[2][C]
  __shared__ uint64_t full[GN_STAGES];
  const int img = blockIdx.y;
  const int cv = threadIdx.x;
  const int V = p.C >> 3;
  const int vx = blockDim.x, vy = blockDim.y;
  float* col = s_red + vy * vx * 16;
  const int tid = threadIdx.y * vx + threadIdx.x;
  GnRing r;
  r.base = gsm;
  r.full = full;
  r.pbeg = blockIdx.x * p.pix_per_cta;
  r.pend = min(p.hw, r.pbeg + p.pix_per_cta);
  r.nst = (r.pend - r.pbeg + p.stage_pix - 1) / p.stage_pix;
  if (tid == 0) {
    for (int i = 0; i < GN_STAGES; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  if (tid == 0)
    for (int i = 0; i < GN_STAGES && i < r.nst; ++i) gn_issue_stage(p, r, img, i);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  const int ch = cv * 8;
  const bool from1 = ch >= p.c0;
  const int cs = from1 ? p.c1 : p.c0;
  const int soff = from1 ? p.stage_pix * p.c0 * 2 + (ch - p.c0) * 2 : ch * 2;
  for (int it = 0; it < r.nst; ++it) {
    const int slot = it % GN_STAGES;
    mbar_wait(&full[slot], static_cast<uint32_t>((it / GN_STAGES) & 1));
    const int npx = min(p.stage_pix, r.pend - (r.pbeg + it * p.stage_pix));
    if (cv < V) {
      const uint8_t* st = gsm + slot * GN_STAGE_BYTES + soff;
      for (int px = threadIdx.y; px < npx; px += vy) {
        const uint4 raw = *reinterpret_cast<const uint4*>(st + static_cast<long long>(px) * cs * 2);
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          s[2 * e] += f.x;
          q[2 * e] += f.x * f.x;
          s[2 * e + 1] += f.y;
          q[2 * e + 1] += f.y * f.y;
        }
      }
    }
    __syncthreads();   // every thread is done with the slot
    if (tid == 0 && it + GN_STAGES < r.nst) gn_issue_stage(p, r, img, it + GN_STAGES);
  }
  float* mine = s_red + (threadIdx.y * vx + cv) * 16;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mine[e] = s[e];
    mine[8 + e] = q[e];
  }
  __syncthreads();
  if (threadIdx.y == 0 && cv < V) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = 0.f, b = 0.f;
      for (int y = 0; y < vy; ++y) {
        a += s_red[(y * vx + cv) * 16 + e];
        b += s_red[(y * vx + cv) * 16 + 8 + e];
      }
      col[cv * 8 + e] = a;
      col[p.C + cv * 8 + e] = b;
    }
  }
  __syncthreads();
  if (tid < p.groups) {
    float a = 0.f, b = 0.f;
    for (int c = tid * p.cpg; c < (tid + 1) * p.cpg; ++c) {
      a += col[c];
      b += col[p.C + c];
    }
    float* dst = p.ws + ((static_cast<long long>(img) * GN_MAX_CHUNKS + blockIdx.x) * p.groups + tid) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

__global__ void gn_apply_bulk_kernel(const GnParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  __shared__ uint64_t full[GN_STAGES];
  __shared__ float s_mean[64], s_rstd[64];
  float* s_part = reinterpret_cast<float*>(gsm + GN_STAGES * GN_STAGE_BYTES);   // [GN_MAX_CHUNKS * 64 * 2]
  const int img = blockIdx.y;
  const int cv = threadIdx.x;
  const int V = p.C >> 3;
  const int ch = cv * 8;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  GnRing r;
  r.base = gsm;
  r.full = full;
  r.pbeg = blockIdx.x * p.pix_per_cta;
  r.pend = min(p.hw, r.pbeg + p.pix_per_cta);
  r.nst = (r.pend - r.pbeg + p.stage_pix - 1) / p.stage_pix;
  if (tid == 0) {
    for (int i = 0; i < GN_STAGES; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  if (tid == 0)
    for (int i = 0; i < GN_STAGES && i < r.nst; ++i) gn_issue_stage(p, r, img, i);
  {
    const float* part = p.ws + static_cast<long long>(img) * GN_MAX_CHUNKS * p.groups * 2;
    const int n = p.nchunks * p.groups * 2;
    for (int i = tid; i < n; i += nthr) s_part[i] = part[i];
  }
  __syncthreads();
  if (tid < p.groups) {
    double sum = 0.0, sq = 0.0;
    for (int k = 0; k < p.nchunks; ++k) {
      sum += static_cast<double>(s_part[(k * p.groups + tid) * 2]);
      sq += static_cast<double>(s_part[(k * p.groups + tid) * 2 + 1]);
    }
    const double cnt = static_cast<double>(p.hw) * p.cpg;
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
  }
  __syncthreads();
  float scale[8], shift[8];
  if (cv < V) {
    float gm[8], bt[8];
    load8(p.gamma + ch, gm);
    load8(p.beta + ch, bt);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (ch + e) / p.cpg;
      scale[e] = gm[e] * s_rstd[g];
      shift[e] = bt[e] - s_mean[g] * scale[e];
    }
  }
  const bool from1 = ch >= p.c0;
  const int cs = from1 ? p.c1 : p.c0;
  const int soff = from1 ? p.stage_pix * p.c0 * 2 + (ch - p.c0) * 2 : ch * 2;
  const int vy = blockDim.y;
  for (int it = 0; it < r.nst; ++it) {
    const int slot = it % GN_STAGES;
    mbar_wait(&full[slot], static_cast<uint32_t>((it / GN_STAGES) & 1));
    const int px0 = r.pbeg + it * p.stage_pix;
    const int npx = min(p.stage_pix, r.pend - px0);
    if (cv < V) {
      const uint8_t* st = gsm + slot * GN_STAGE_BYTES + soff;
      for (int px = threadIdx.y; px < npx; px += vy) {
        const uint4 raw = *reinterpret_cast<const uint4*>(st + static_cast<long long>(px) * cs * 2);
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
        float v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          v[2 * e] = f.x;
          v[2 * e + 1] = f.y;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = v[e] * scale[e] + shift[e];
          if (p.silu) y = __fdividef(y, 1.0f + __expf(-y));
          v[e] = y;
        }
        const int pp = px0 + px;
        long long orow = static_cast<long long>(img) * p.hw + pp;
        if (p.out_chunk_pix > 0) {
          const int d = pp / p.out_chunk_pix;
          orow = (static_cast<long long>(d) * p.nimg_total + (p.img0 + img)) * p.out_chunk_pix + (pp - d * p.out_chunk_pix);
        }
        store8(p.out + orow * p.C + ch, v);
      }
    }
    __syncthreads();
    if (tid == 0 && it + GN_STAGES < r.nst) gn_issue_stage(p, r, img, it + GN_STAGES);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per R rows, rows held in registers (C <= 8*32*LN_MAXV), two-pass variance.
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 6;  // C <= 1536

struct LnParams {
  const __half* x;
  long long rows;
  int c;
  float eps;
  const __half* gamma;
  const __half* beta;
  __half* out;
  const __half* add;
  __half* out2;
  long long add_row0;
};

// NV = 16-byte vectors per lane per row, R = rows a warp processes together (R*NV loads in flight
// per lane: the kernel is latency-bound otherwise).
template <int NV, int R>
__global__ void layernorm_kernel(const LnParams p) {
  pdl_wait();
  pdl_trigger();
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int V = p.c >> 3;
  const long long w0 = static_cast<long long>(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5);
  const long long wstride = static_cast<long long>(gridDim.x) * warps_per_cta;
  const float inv_c = 1.0f / static_cast<float>(p.c);
  for (long long rb = w0 * R; rb < p.rows; rb += wstride * R) {
    uint4 raw[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int cv = lane + i * 32;
        raw[r][i] = (rb + r < p.rows && cv < V)
                        ? *reinterpret_cast<const uint4*>(p.x + (rb + r) * p.c + cv * 8)
                        : make_uint4(0, 0, 0, 0);
      }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[r][i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          sum += f.x + f.y;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      mean[r] = sum * inv_c;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < V) {
          const __half2* h = reinterpret_cast<const __half2*>(&raw[r][i]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            const float d0 = f.x - mean[r], d1 = f.y - mean[r];
            sq += d0 * d0 + d1 * d1;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      rstd[r] = rsqrtf(sq * inv_c + p.eps);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int cv = lane + i * 32;
      if (cv < V) {
        float gm[8], bt[8];
        load8(p.gamma + cv * 8, gm);
        load8(p.beta + cv * 8, bt);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const long long row = rb + r;
          if (row < p.rows) {
            const __half2* h = reinterpret_cast<const __half2*>(&raw[r][i]);
            float y[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h[e]);
              y[2 * e] = (f.x - mean[r]) * rstd[r] * gm[2 * e] + bt[2 * e];
              y[2 * e + 1] = (f.y - mean[r]) * rstd[r] * gm[2 * e + 1] + bt[2 * e + 1];
            }
            store8(p.out + row * p.c + cv * 8, y);
            if (p.add != nullptr && row >= p.add_row0) {
              float ad[8];
              const long long r2 = row - p.add_row0;
              load8(p.add + r2 * p.c + cv * 8, ad);
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] += ad[e];
              store8(p.out2 + r2 * p.c + cv * 8, y);
            }
          }
        }
      }
    }
  }
}

// Rows of c = 40 * LPR halves (320 / 640 channels: LPR = 8 / 16): a group of LPR lanes owns a row and
// every lane holds exactly 5 16-byte vectors, so all 32 lanes of every load carry data and each
// group reads whole 128-byte lines (the generic kernel above runs 40- and 80-vector rows on 64 / 96
// lane slots: 62 % / 83 % of the lanes busy).  RR rows per group in flight.
template <int LPR, int RR>
__global__ void layernorm_lpr_kernel(const LnParams p) {
  pdl_wait();
  pdl_trigger();
  constexpr int GPW = 32 / LPR;  // row groups per warp
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, grp = lane / LPR;
  const long long w0 = static_cast<long long>(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5);
  const long long wstride = static_cast<long long>(gridDim.x) * warps_per_cta;
  const float inv_c = 1.0f / static_cast<float>(p.c);
  for (long long rb = w0 * (GPW * RR); rb < p.rows; rb += wstride * (GPW * RR)) {
    uint4 raw[RR][5];
#pragma unroll
    for (int r = 0; r < RR; ++r) {
      const long long row = rb + r * GPW + grp;
#pragma unroll
      for (int i = 0; i < 5; ++i)
        raw[r][i] = (row < p.rows) ? *reinterpret_cast<const uint4*>(p.x + row * p.c + (sub + i * LPR) * 8)
                                   : make_uint4(0, 0, 0, 0);
    }
    float mean[RR], rstd[RR];
#pragma unroll
    for (int r = 0; r < RR; ++r) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[r][i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          sum += f.x + f.y;
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      mean[r] = sum * inv_c;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[r][i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          const float d0 = f.x - mean[r], d1 = f.y - mean[r];
          sq += d0 * d0 + d1 * d1;
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      rstd[r] = rsqrtf(sq * inv_c + p.eps);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int cv = sub + i * LPR;
      float gm[8], bt[8];
      load8(p.gamma + cv * 8, gm);
      load8(p.beta + cv * 8, bt);
#pragma unroll
      for (int r = 0; r < RR; ++r) {
        const long long row = rb + r * GPW + grp;
        if (row < p.rows) {
          const __half2* h = reinterpret_cast<const __half2*>(&raw[r][i]);
          float y[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            y[2 * e] = (f.x - mean[r]) * rstd[r] * gm[2 * e] + bt[2 * e];
            y[2 * e + 1] = (f.y - mean[r]) * rstd[r] * gm[2 * e + 1] + bt[2 * e + 1];
          }
          store8(p.out + row * p.c + cv * 8, y);
          if (p.add != nullptr && row >= p.add_row0) {
            float ad[8];
            const long long r2 = row - p.add_row0;
            load8(p.add + r2 * p.c + cv * 8, ad);
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] += ad[e];
            store8(p.out2 + r2 * p.c + cv * 8, y);
          }
        }
      }
    }
  }
}

template <int LPR, int RR>
static void launch_ln_lpr(const LnParams& p, int num_sms, cudaStream_t stream) {
  const int warps = 8;
  const long long per_warp = (32 / LPR) * RR;
  long long groups = (p.rows + per_warp - 1) / per_warp;
  long long ctas = (groups + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (ctas > cap) ctas = cap;
  launch_pdl(layernorm_lpr_kernel<LPR, RR>, dim3(static_cast<unsigned>(ctas)), dim3(warps * 32), 0, stream, p);
}

template <int NV, int R>
static void launch_ln(const LnParams& p, int num_sms, cudaStream_t stream) {
  const int warps = 8;
  long long groups = (p.rows + R - 1) / R;
  long long ctas = (groups + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (ctas > cap) ctas = cap;
  launch_pdl(layernorm_kernel<NV, R>, dim3(static_cast<unsigned>(ctas)), dim3(warps * 32), 0, stream, p);
}

}  // namespace mdk

extern "C" int64_t mdk_groupnorm_ws_bytes(int32_t nimg, int32_t groups) {
  return static_cast<int64_t>(nimg) * mdk::GN_MAX_CHUNKS * groups * 2 * sizeof(float);
}

extern "C" int mdk_groupnorm_f16(mdk_ctx* ctx, const mdk_gn_args* a, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && a && a->x0 && a->out && a->ws && a->gamma && a->beta,
              "mdk_groupnorm_f16: null argument");
  const int C = a->c0 + a->c1;
  MDK_REQUIRE(a->c0 % 8 == 0 && a->c1 % 8 == 0 && C > 0, "mdk_groupnorm_f16: c0=%d c1=%d must be %%8",
              a->c0, a->c1);
  MDK_REQUIRE(a->groups <= 64, "mdk_groupnorm_f16: at most 64 groups");
  MDK_REQUIRE(a->groups > 0 && C % a->groups == 0, "mdk_groupnorm_f16: C=%d not divisible by groups=%d",
              C, a->groups);
  MDK_REQUIRE(C / 8 <= 1024, "mdk_groupnorm_f16: C=%d too large", C);
  MDK_REQUIRE(a->c1 == 0 || a->x1 != nullptr, "mdk_groupnorm_f16: x1 is NULL");
  GnParams p;
  p.x0 = static_cast<const __half*>(a->x0);
  p.x1 = static_cast<const __half*>(a->x1);
  p.c0 = a->c0;
  p.c1 = a->c1;
  p.C = C;
  p.nimg = a->nimg;
  p.hw = a->hw;
  p.groups = a->groups;
  p.cpg = C / a->groups;
  p.eps = a->eps;
  p.gamma = static_cast<const __half*>(a->gamma);
  p.beta = static_cast<const __half*>(a->beta);
  p.silu = a->silu;
  p.out = static_cast<__half*>(a->out);
  p.ws = static_cast<float*>(a->ws);
  p.out_chunk_pix = a->out_chunk_pix;
  p.nimg_total = a->nimg;
  p.img0 = 0;
  MDK_REQUIRE(a->out_chunk_pix == 0 || (a->out_chunk_pix > 0 && a->out_chunks > 0 &&
                                        static_cast<long long>(a->out_chunk_pix) * a->out_chunks >= a->hw),
              "mdk_groupnorm_f16: exchange layout %d x %d does not cover hw=%d", a->out_chunks, a->out_chunk_pix, a->hw);
  const int V = C / 8;
  // blockDim.x = V exactly (not rounded up to a warp multiple): thread (x, y) reads vector x of pixel
  // y, so with a single source the linear thread id walks memory contiguously and every lane of
  // every warp carries data (C = 320: 40 vectors per pixel; a 64-wide block left 37 % of the lanes
  // idle).  MDK_GN_VX32=1 restores the rounded width.
  static int vx32 = -1;
  if (vx32 < 0) {
    const char* e = getenv("MDK_GN_VX32");
    vx32 = e ? atoi(e) : 0;
  }
  const int vx = vx32 ? ((V + 31) / 32) * 32 : V;
  int vy = 512 / vx;
  if (vy < 1) vy = 1;
  MDK_REQUIRE(a->groups <= vx * vy, "mdk_groupnorm_f16: too many groups");
  const size_t stats_smem = (static_cast<size_t>(vx) * vy * 16 + 2 * static_cast<size_t>(C)) * sizeof(float);
  MDK_REQUIRE(stats_smem <= 48 * 1024, "mdk_groupnorm_f16: C=%d too large", C);
  // The apply pass re-reads what the statistics pass just read.  Running the two passes over groups of
  // images small enough to stay in L2 (126 MB) turns that second read into L2 hits: HBM traffic
  // 2 passes instead of 3.  Measured on B200: SLOWER (0.246 vs 0.169 ms at 32 x 9216 x 320 with
  // 48 MB groups, 0.278 ms with 24 MB): the extra launches cost more than the L2 hits save, so the
  // default is MDK_GN_CHUNK_MB=0 (whole batch in one go); the switch stays for experiments.
  static int chunk_mb = -1;
  if (chunk_mb < 0) {
    const char* e = getenv("MDK_GN_CHUNK_MB");
    chunk_mb = e ? atoi(e) : 0;
  }
  const long long img_bytes = static_cast<long long>(a->hw) * C * 2;
  int img_per_group = a->nimg;
  if (chunk_mb > 0) {
    long long g = (static_cast<long long>(chunk_mb) << 20) / (img_bytes > 0 ? img_bytes : 1);
    if (g < 1) g = 1;
    if (g < img_per_group) img_per_group = static_cast<int>(g);
  }
  for (int i0 = 0; i0 < a->nimg; i0 += img_per_group) {
    const int ni = (a->nimg - i0 < img_per_group) ? (a->nimg - i0) : img_per_group;
    GnParams q = p;
    q.nimg = ni;
    q.x0 = p.x0 + static_cast<long long>(i0) * a->hw * a->c0;
    if (p.x1) q.x1 = p.x1 + static_cast<long long>(i0) * a->hw * a->c1;
    q.out = (p.out_chunk_pix > 0) ? p.out : p.out + static_cast<long long>(i0) * a->hw * C;
    q.img0 = i0;
    q.ws = p.ws + static_cast<long long>(i0) * GN_MAX_CHUNKS * a->groups * 2;
    // CTAs per SM over the whole launch (MDK_GN_OVERSUB): both kernels fit 2 CTAs per SM; every CTA pays a fixed
    // ~5 us (launch, partial-sum fetch / smem reduction tail) that is not overlapped by its only neighbour
    static int oversub = -1;
    if (oversub < 0) {
      const char* e = getenv("MDK_GN_OVERSUB");
      oversub = e ? atoi(e) : 3;   // measured (profiles/r02_gn_sweep.log): 3 -> 0.141 / 0.080 / 0.045 ms at L0 / L1 / L2,
                                   // 8 (round 1) -> 0.140 / 0.090 / 0.063 ms
      if (oversub < 1) oversub = 1;
    }
    int chunks = (ctx->num_sms * oversub + ni - 1) / ni;
    const int max_chunks = (a->hw + vy - 1) / vy;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
    if (chunks < 1) chunks = 1;
    q.pix_per_cta = (a->hw + chunks - 1) / chunks;
    chunks = (a->hw + q.pix_per_cta - 1) / q.pix_per_cta;
    q.nchunks = chunks;
    dim3 block(vx, vy);
    dim3 grid(chunks, ni);
    int bulk = -1;   // read per call: tests switch kernels in-process
    {
      const char* e = getenv("MDK_GN_BULK");
      bulk = e ? atoi(e) : 0;   // measured: no gain over the register-staged kernels (0.149 vs 0.145 ms at L0): off
    }
    // bulk (cp.async.bulk ring) kernels: sources 16-byte aligned, one pixel of both sources fits a stage
    q.stage_pix = GN_STAGE_BYTES / (C * 2);
    const bool can_bulk = bulk && q.stage_pix >= 1 && (reinterpret_cast<uintptr_t>(q.x0) & 15) == 0 &&
                          (q.x1 == nullptr || (reinterpret_cast<uintptr_t>(q.x1) & 15) == 0);
    if (can_bulk) {
      const size_t ring = static_cast<size_t>(GN_STAGES) * GN_STAGE_BYTES;
      const size_t smem_s = ring + stats_smem;
      const size_t smem_a = ring + static_cast<size_t>(GN_MAX_CHUNKS) * 64 * 2 * sizeof(float);
      static unsigned long long gn_attr_mask = 0;
      if (((gn_attr_mask >> (ctx->device & 63)) & 1ull) == 0) {
        MDK_CHECK_CUDA(cudaFuncSetAttribute(gn_stats_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        MDK_CHECK_CUDA(cudaFuncSetAttribute(gn_apply_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        gn_attr_mask |= 1ull << (ctx->device & 63);
      }
      MDK_CHECK_CUDA(launch_pdl(gn_stats_bulk_kernel, grid, block, smem_s, stream, q));
      count_launch();
      MDK_CHECK_CUDA(launch_pdl(gn_apply_bulk_kernel, grid, block, smem_a, stream, q));
      count_launch();
      continue;
    }
    MDK_CHECK_CUDA(launch_pdl(gn_stats_kernel, grid, block, stats_smem, stream, q));
    count_launch();
    MDK_CHECK_CUDA(launch_pdl(gn_apply_kernel, grid, block, 0, stream, q));
    count_launch();
  }
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_layernorm_f16(mdk_ctx* ctx, const mdk_ln_args* a, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && a && a->x && a->out && a->gamma && a->beta, "mdk_layernorm_f16: null argument");
  MDK_REQUIRE(a->c % 8 == 0 && a->c > 0 && a->c <= 8 * 32 * LN_MAXV,
              "mdk_layernorm_f16: c=%d must be a multiple of 8 and <= %d", a->c, 8 * 32 * LN_MAXV);
  MDK_REQUIRE((a->add == nullptr) == (a->out2 == nullptr), "mdk_layernorm_f16: add/out2 mismatch");
  if (a->rows <= 0) return 0;
  LnParams p;
  p.x = static_cast<const __half*>(a->x);
  p.rows = a->rows;
  p.c = a->c;
  p.eps = a->eps;
  p.gamma = static_cast<const __half*>(a->gamma);
  p.beta = static_cast<const __half*>(a->beta);
  p.out = static_cast<__half*>(a->out);
  p.add = static_cast<const __half*>(a->add);
  p.out2 = static_cast<__half*>(a->out2);
  p.add_row0 = a->add_row0;
  static int lpr_on = -1;
  if (lpr_on < 0) {
    const char* e = getenv("MDK_LN_LPR");
    lpr_on = e ? atoi(e) : 1;
  }
  // measured: c = 320 0.087 vs 0.105 ms (294912 rows), c = 640 0.049 vs 0.045 ms -> 320 only (2 = both)
  if ((lpr_on && a->c == 320) || (lpr_on == 2 && a->c == 640)) {
    if (a->c == 320)
      launch_ln_lpr<8, 2>(p, ctx->num_sms, stream);
    else
      launch_ln_lpr<16, 2>(p, ctx->num_sms, stream);
    count_launch();
    MDK_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  switch ((a->c / 8 + 31) / 32) {
    case 1: launch_ln<1, 4>(p, ctx->num_sms, stream); break;
    case 2: launch_ln<2, 3>(p, ctx->num_sms, stream); break;
    case 3: launch_ln<3, 2>(p, ctx->num_sms, stream); break;
    case 4: launch_ln<4, 1>(p, ctx->num_sms, stream); break;
    case 5: launch_ln<5, 1>(p, ctx->num_sms, stream); break;
    default: launch_ln<6, 1>(p, ctx->num_sms, stream); break;
  }
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
