// Small HBM/latency-bound kernels around the tensor-core path: nearest x2 upsample, stride-2
// im2col, latent layout changes, window accumulate, fused CFG + DDIM update, timestep embedding.
#include "host_common.h"
#include "ptx.cuh"
#include "../../include/mdk.h"

namespace mdk {

// ---------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int nimg,
                                  int h, int w, int cv) {
  const long long total = static_cast<long long>(nimg) * (2 * h) * (2 * w) * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    long long t = i / cv;
    const int ox = static_cast<int>(t % (2 * w));
    t /= (2 * w);
    const int oy = static_cast<int>(t % (2 * h));
    const long long img = t / (2 * h);
    out[i] = x[((img * h + (oy >> 1)) * w + (ox >> 1)) * cv + v];
  }
}

__global__ void im2col3x3_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int nimg,
                                 int h, int w, int cv, int stride, int ho, int wo, int kpadv) {
  // one thread per 16-byte vector of the output matrix [nimg*ho*wo, kpadv]
  const long long total = static_cast<long long>(nimg) * ho * wo * kpadv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kv = static_cast<int>(i % kpadv);
    long long m = i / kpadv;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (kv < 9 * cv) {
      const int tap = kv / cv;
      const int v = kv % cv;
      const int ox = static_cast<int>(m % wo);
      const int oy = static_cast<int>((m / wo) % ho);
      const long long img = m / (static_cast<long long>(wo) * ho);
      const int iy = oy * stride + tap / 3 - 1;
      const int ix = ox * stride + tap % 3 - 1;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) val = x[((img * h + iy) * w + ix) * cv + v];
    }
    out[i] = val;
  }
}

// sample [b_src, c, f_total, hw] fp16 -> out [(b fl), hw, cpad] fp16 (channels >= c zeroed)
__global__ void latents_to_nhwc_kernel(const __half* __restrict__ sample, __half* __restrict__ out,
                                       int b, int b_src, int c, int f_total,
                                       const int* __restrict__ frame_idx, int fl, int hw, int cpad) {
  const long long total = static_cast<long long>(b) * fl * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % hw);
    const int j = static_cast<int>((i / hw) % fl);
    const int bi = static_cast<int>(i / (static_cast<long long>(hw) * fl));
    const int fr = frame_idx ? frame_idx[j] : j;
    const int bs = bi % b_src;
    for (int ch = 0; ch < cpad; ++ch) {
      __half v = __float2half(0.f);
      if (ch < c) v = sample[((static_cast<long long>(bs) * c + ch) * f_total + fr) * hw + px];
      out[i * cpad + ch] = v;
    }
  }
}

// pred [(b fl), hw, cpad] fp16 -> acc[b, c, frame, hw] += ; counter[frame] += 1
__global__ void pred_accumulate_kernel(const __half* __restrict__ pred, float* __restrict__ acc,
                                       float* __restrict__ counter, int b, int c, int f_total,
                                       const int* __restrict__ frame_idx, int fl, int hw, int cpad) {
  const long long total = static_cast<long long>(b) * fl * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % hw);
    const int j = static_cast<int>((i / hw) % fl);
    const int bi = static_cast<int>(i / (static_cast<long long>(hw) * fl));
    const int fr = frame_idx ? frame_idx[j] : j;
    for (int ch = 0; ch < c; ++ch) {
      const long long o = ((static_cast<long long>(bi) * c + ch) * f_total + fr) * hw + px;
      acc[o] += __half2float(pred[i * cpad + ch]);
    }
    if (counter != nullptr && px == 0 && bi == 0) counter[fr] += 1.0f;
  }
}

__global__ void cfg_ddim_kernel(const float* __restrict__ acc, const float* __restrict__ counter,
                                __half* __restrict__ latents, const float* __restrict__ coef,
                                float gs, int nb, int c, int f, int hw, int vpred) {
  const long long total = static_cast<long long>(c) * f * hw;
  const float sa = coef[0], sb = coef[1], sap = coef[2], sbp = coef[3];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int fr = static_cast<int>((i / hw) % f);
    const float inv = 1.0f / counter[fr];
    float g;
    if (nb == 2) {
      const float eu = acc[i] * inv;
      const float ec = acc[total + i] * inv;
      g = eu + gs * (ec - eu);
    } else {
      g = acc[i] * inv;
    }
    const float x = __half2float(latents[i]);
    float x0, e;
    if (vpred) {
      x0 = sa * x - sb * g;
      e = sa * g + sb * x;
    } else {
      x0 = (x - sb * g) / sa;
      e = g;
    }
    latents[i] = __float2half_rn(sap * x0 + sbp * e);
  }
}

// ---------------------------------------------------------------------------------------------
// GEMV: out[r] = act(W[r, :] . x + b[r]); one warp per row; x staged in shared memory as fp32.
// mode 0: x = sinusoid(timestep) computed in place; mode 1: x read from xin (fp32).
// ---------------------------------------------------------------------------------------------
struct GemvParams {
  const long long* timestep;
  int flip;
  float freq_shift;
  const float* xin;
  int cols;
  const __half* w;
  const __half* b16;  // fp16 bias or NULL
  const float* b32;   // fp32 bias or NULL
  int rows;
  int silu_out;
  float* out;
  int mode;
};

__global__ void gemv_kernel(const GemvParams p) {
  extern __shared__ float sx[];
  if (p.mode == 0) {
    const float t = static_cast<float>(*p.timestep);
    const int half = p.cols / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const float freq = expf(-logf(10000.0f) * static_cast<float>(i) /
                              (static_cast<float>(half) - p.freq_shift));
      const float arg = t * freq;
      const float s = sinf(arg), c = cosf(arg);
      if (p.flip) {
        sx[i] = c;
        sx[half + i] = s;
      } else {
        sx[i] = s;
        sx[half + i] = c;
      }
    }
  } else {
    for (int i = threadIdx.x; i < p.cols; i += blockDim.x) sx[i] = p.xin[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < p.rows; r += gridDim.x * warps) {
    const __half* wr = p.w + static_cast<long long>(r) * p.cols;
    float acc = 0.f;
    for (int c0 = lane * 8; c0 < p.cols; c0 += 32 * 8) {
      uint4 raw = *reinterpret_cast<const uint4*>(wr + c0);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(h[e]);
        acc += f.x * sx[c0 + 2 * e] + f.y * sx[c0 + 2 * e + 1];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (p.b16) acc += __half2float(p.b16[r]);
      if (p.b32) acc += p.b32[r];
      if (p.silu_out) acc = acc / (1.0f + expf(-acc));
      p.out[r] = acc;
    }
  }
}

static inline int grid_for(long long total, int block, int num_sms) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}


// out[i, px, :] = back[px / pp, i, px % pp, :] (+ x[i, px, :])   (16-byte vectors, fp32 add; x may be NULL)
__global__ void unshard_add_kernel(const __half* __restrict__ back, const __half* __restrict__ x,
                                   __half* __restrict__ out, int nimg, int hw, int pp, int cv) {
  const long long total = static_cast<long long>(nimg) * hw * cv;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(t % cv);
    const long long row = t / cv;
    const int px = static_cast<int>(row % hw);
    const int img = static_cast<int>(row / hw);
    const int d = px / pp;
    const long long brow = (static_cast<long long>(d) * nimg + img) * pp + (px - d * pp);
    const uint4 a = *reinterpret_cast<const uint4*>(back + (brow * cv + v) * 8);
    if (x == nullptr) {   // pure layout change (bit-exact)
      *reinterpret_cast<uint4*>(out + (row * cv + v) * 8) = a;
      continue;
    }
    const uint4 b = *reinterpret_cast<const uint4*>(x + (row * cv + v) * 8);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
      const __half2 r = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
      ow[e] = *reinterpret_cast<const uint32_t*>(&r);
    }
    *reinterpret_cast<uint4*>(out + (row * cv + v) * 8) = o;
  }
}

}  // namespace mdk

using namespace mdk;

extern "C" int mdk_upsample2x_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t h,
                                  int32_t w, int32_t c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && x && out, "mdk_upsample2x_f16: null argument");
  MDK_REQUIRE(c % 8 == 0, "mdk_upsample2x_f16: c=%d must be a multiple of 8", c);
  const long long total = static_cast<long long>(nimg) * 4 * h * w * (c / 8);
  upsample2x_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, stream>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), nimg, h, w, c / 8);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_im2col3x3_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t h,
                                 int32_t w, int32_t c, int32_t stride, int32_t kpad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && x && out, "mdk_im2col3x3_f16: null argument");
  MDK_REQUIRE(c % 8 == 0 && kpad % 8 == 0 && kpad >= 9 * c && (stride == 1 || stride == 2),
              "mdk_im2col3x3_f16: bad c=%d kpad=%d stride=%d", c, kpad, stride);
  const int ho = (h + 2 - 3) / stride + 1;
  const int wo = (w + 2 - 3) / stride + 1;
  const long long total = static_cast<long long>(nimg) * ho * wo * (kpad / 8);
  im2col3x3_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, stream>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), nimg, h, w, c / 8, stride, ho, wo,
      kpad / 8);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_latents_to_nhwc(mdk_ctx* ctx, const void* sample, void* out, int32_t b,
                                   int32_t b_src, int32_t c, int32_t f_total,
                                   const int32_t* frame_idx, int32_t fl, int32_t hw, int32_t cpad,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && sample && out && b_src > 0 && cpad >= c, "mdk_latents_to_nhwc: bad argument");
  const long long total = static_cast<long long>(b) * fl * hw;
  latents_to_nhwc_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, stream>>>(
      static_cast<const __half*>(sample), static_cast<__half*>(out), b, b_src, c, f_total, frame_idx,
      fl, hw, cpad);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_unshard_add_f16(mdk_ctx* ctx, const void* back, const void* x, void* out, int32_t nimg,
                                   int32_t hw, int32_t chunk_pix, int32_t c, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && back && out, "mdk_unshard_add_f16: null argument");
  MDK_REQUIRE(c % 8 == 0 && c > 0 && chunk_pix > 0 && nimg > 0 && hw > 0, "mdk_unshard_add_f16: bad shape");
  const long long total = static_cast<long long>(nimg) * hw * (c / 8);
  unshard_add_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, stream>>>(
      static_cast<const __half*>(back), static_cast<const __half*>(x), static_cast<__half*>(out), nimg, hw,
      chunk_pix, c / 8);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_pred_accumulate(mdk_ctx* ctx, const void* pred, float* acc, float* counter,
                                   int32_t b, int32_t c, int32_t f_total, const int32_t* frame_idx,
                                   int32_t fl, int32_t hw, int32_t cpad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && pred && acc, "mdk_pred_accumulate: null argument");
  const long long total = static_cast<long long>(b) * fl * hw;
  pred_accumulate_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, stream>>>(
      static_cast<const __half*>(pred), acc, counter, b, c, f_total, frame_idx, fl, hw, cpad);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_cfg_ddim_step(mdk_ctx* ctx, const float* acc, const float* counter,
                                 void* latents, const float* coef, float guidance_scale, int32_t nb,
                                 int32_t c, int32_t f, int32_t hw, int32_t v_prediction,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && acc && counter && latents && coef, "mdk_cfg_ddim_step: null argument");
  MDK_REQUIRE(nb == 1 || nb == 2, "mdk_cfg_ddim_step: nb must be 1 or 2");
  const long long total = static_cast<long long>(c) * f * hw;
  cfg_ddim_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, stream>>>(
      acc, counter, static_cast<__half*>(latents), coef, guidance_scale, nb, c, f, hw, v_prediction);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_time_embed_f16(mdk_ctx* ctx, const mdk_temb_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && a && a->timestep && a->w1 && a->w2 && a->scratch, "mdk_time_embed_f16: null arg");
  MDK_REQUIRE(a->dim % 16 == 0 && a->edim % 8 == 0, "mdk_time_embed_f16: bad dims %d %d", a->dim,
              a->edim);
  float* h1 = a->scratch;            // [edim]  silu(linear_1(sinusoid))
  float* h2 = a->scratch + a->edim;  // [edim]  silu(linear_2(h1))
  const int warps = 8;
  GemvParams p;
  memset(&p, 0, sizeof(p));
  p.timestep = reinterpret_cast<const long long*>(a->timestep);
  p.flip = a->flip_sin_to_cos;
  p.freq_shift = a->freq_shift;
  // stage 1
  p.mode = 0;
  p.cols = a->dim;
  p.w = static_cast<const __half*>(a->w1);
  p.b16 = static_cast<const __half*>(a->b1);
  p.rows = a->edim;
  p.silu_out = 1;
  p.out = h1;
  gemv_kernel<<<(a->edim + warps - 1) / warps, warps * 32, a->dim * sizeof(float), stream>>>(p);
  count_launch();
  // stage 2
  p.mode = 1;
  p.xin = h1;
  p.cols = a->edim;
  p.w = static_cast<const __half*>(a->w2);
  p.b16 = static_cast<const __half*>(a->b2);
  p.rows = a->edim;
  p.silu_out = 1;
  p.out = h2;
  gemv_kernel<<<(a->edim + warps - 1) / warps, warps * 32, a->edim * sizeof(float), stream>>>(p);
  count_launch();
  // stage 3: every resnet's time_emb_proj in one launch
  if (a->nrows > 0) {
    MDK_REQUIRE(a->proj_w && a->temb_out, "mdk_time_embed_f16: null proj_w/temb_out");
    p.xin = h2;
    p.w = static_cast<const __half*>(a->proj_w);
    p.b16 = nullptr;
    p.b32 = a->proj_b;
    p.rows = a->nrows;
    p.silu_out = 0;
    p.out = a->temb_out;
    int g = (a->nrows + warps - 1) / warps;
    if (g > ctx->num_sms * 8) g = ctx->num_sms * 8;
    gemv_kernel<<<g, warps * 32, a->edim * sizeof(float), stream>>>(p);
    count_launch();
  }
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
