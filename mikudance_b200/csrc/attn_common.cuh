// Definitions shared by the attention kernels (attn_tc.cu: one softmax stream per CTA and its variants;
// attn_2s.cu: two softmax streams per CTA).
#pragma once
#include "host_common.h"
#include "ptx.cuh"
#include "../../include/mdk.h"

namespace mdk {

constexpr int ATT_BQ = 128;
constexpr int ATT_THREADS = 192;
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units
// exp2 of two scores at once on the FMA pipe, in half2 (P is rounded to fp16 anyway): round-to-nearest
// split x = n + f through the magic constant 1551 = 0x660F (the low 5 bits of the sum are n + 15, the
// fp16 exponent field of 2^n), degree-3 polynomial for 2^f on [-0.5, 0.5], product with 2^n.  13
// issue slots per pair, none of them on the MUFU pipe / MIO queue that bound this kernel.
__device__ __forceinline__ uint32_t ex2_poly_h2(float x0, float x1) {
  const __half2 lo = __floats2half2_rn(-15.0f, -15.0f);
  const __half2 magic = __floats2half2_rn(1551.0f, 1551.0f);
  __half2 x = __hmax2(__floats2half2_rn(x0, x1), lo);
  const __half2 t = __hadd2(x, magic);
  const __half2 n = __hsub2(t, magic);
  const __half2 f = __hsub2(x, n);
  __half2 pl = __hfma2(__floats2half2_rn(0.05517165f, 0.05517165f), f, __floats2half2_rn(0.24261113f, 0.24261113f));
  pl = __hfma2(pl, f, __floats2half2_rn(0.69326097f, 0.69326097f));
  pl = __hfma2(pl, f, __floats2half2_rn(0.99992806f, 0.99992806f));
  const uint32_t tb = *reinterpret_cast<const uint32_t*>(&t);
  const uint32_t eb = (tb & 0x001F001Fu) << 10;
  const __half2 r = __hmul2(pl, *reinterpret_cast<const __half2*>(&eb));
  return *reinterpret_cast<const uint32_t*>(&r);
}

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  __half* out;
  long long ldo;
  int lq, lkv, heads, d;
  int dk16;  // ceil(d / 16): K steps of Q K^T
  int dn;    // ceil16(d): N of the P V MMA
  int kv_div;
  int n_kv_tiles;
  float scale_log2;
  int vt_head_rows;  // rows per head in V^T (>= d)
  int poly;          // 1: every second pair of exponentials on the FMA pipe (ex2_poly_h2)
  long long* trace;  // TRACE instantiation only: clock64 timeline of one CTA, [tile][16 slots]
  int trace_cap;     // tiles the trace buffer holds
  int stagger;       // two-stream kernels: cycles by which stream 1 (and odd CTAs, x2) start late (MDK_ATTN_STAGGER)
};


// Q / K: 4-D [d, heads, L, image] (box 64 x 1 x rows x 1); V^T: 3-D [L, heads * vt_head_rows, image]
inline int encode_attn_maps(AttnParams& p, const mdk_attn_args* a, int bkv) {
  const uint64_t d = static_cast<uint64_t>(a->d);
  {
    uint64_t dims[4] = {d, static_cast<uint64_t>(a->heads), static_cast<uint64_t>(a->lq),
                        static_cast<uint64_t>(a->nimg)};
    uint64_t str[4] = {0, d * 2, static_cast<uint64_t>(a->ldq) * 2,
                       static_cast<uint64_t>(a->ldq) * 2 * a->lq};
    uint32_t box[4] = {64, 1, ATT_BQ, 1};
    if (encode_tmap_f16(&p.tmQ, a->q, 4, dims, str, box)) return -1;
  }
  {
    uint64_t dims[4] = {d, static_cast<uint64_t>(a->heads), static_cast<uint64_t>(a->lkv),
                        static_cast<uint64_t>(a->nkv)};
    uint64_t str[4] = {0, d * 2, static_cast<uint64_t>(a->ldk) * 2,
                       static_cast<uint64_t>(a->ldk) * 2 * a->lkv};
    uint32_t box[4] = {64, 1, static_cast<uint32_t>(bkv), 1};
    if (encode_tmap_f16(&p.tmK, a->k, 4, dims, str, box)) return -1;
  }
  {
    const uint64_t C = static_cast<uint64_t>(p.vt_head_rows) * a->heads;   // V^T rows per image
    uint64_t dims[3] = {static_cast<uint64_t>(a->lkv), C, static_cast<uint64_t>(a->nkv)};
    uint64_t str[3] = {0, static_cast<uint64_t>(a->ldvt) * 2, static_cast<uint64_t>(a->ldvt) * 2 * C};
    uint32_t box[3] = {64, static_cast<uint32_t>(p.dn), 1};
    if (encode_tmap_f16(&p.tmV, a->vt, 3, dims, str, box)) return -1;
  }
  return 0;
}


// attn_2s.cu: two independent softmax streams per CTA (head_dim <= 64)
// variant 1: P through shared memory, 2: P in tensor memory; trace != nullptr: timeline instantiation
int launch_attn_2s(const mdk_ctx* ctx, AttnParams& p, const mdk_attn_args* a, cudaStream_t stream, int variant,
                   long long* trace, int trace_cap);

// attn_2s32.cu: the two-stream kernel with 32-key sub-tiles (S half-tiles as a double buffer)
int launch_attn_2s32(const mdk_ctx* ctx, AttnParams& p, const mdk_attn_args* a, cudaStream_t stream);

}  // namespace mdk
