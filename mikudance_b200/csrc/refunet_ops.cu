// Kernels of the stages around the denoising loop.  CLIP image encoder: the in-place
// QuickGELU of its MLP.  Reference UNet (writer):
//   * condition-latent layout change: channel slice of an NCHW batch -> zero-padded NHWC, with the
//     nearest-neighbour resize MANModule applies to the scene-motion map,
//   * in-place ReLU (MANModule.mlp_shared),
//   * MANModule's parameter-free InstanceNorm2d + (1 + gamma) / beta modulation.
// All three are HBM-bound SIMT kernels: fp32 math, 16-byte vector IO on the channel axis.
//
// Algorithmic bytes (fp16): cond_to_nhwc 2*nimg*(h*w*c + ho*wo*cpad); relu 2*2*n;
// man_modulate 2*nimg*hw*C*(2 reads of x + gamma + beta + 1 write) = 10*nimg*hw*C.
#include "host_common.h"
#include "ptx.cuh"
#include "../../include/mdk.h"

namespace mdk {

// x [nimg, ctot, h, w] fp16 (NCHW) -> out [nimg, ho, wo, cpad] fp16: channels [c_first, c_first + c)
// sampled at (floor(oy*h/ho), floor(ox*w/wo)); channels >= c are zero.
__global__ void cond_to_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ out, int nimg,
                                    int ctot, int c_first, int c, int h, int w, int ho, int wo,
                                    int cpad) {
  const long long total = static_cast<long long>(nimg) * ho * wo;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % wo);
    const int oy = static_cast<int>((i / wo) % ho);
    const long long img = i / (static_cast<long long>(wo) * ho);
    const int iy = min(h - 1, static_cast<int>((static_cast<long long>(oy) * h) / ho));
    const int ix = min(w - 1, static_cast<int>((static_cast<long long>(ox) * w) / wo));
    const __half* src = x + ((img * ctot + c_first) * h + iy) * w + ix;
    __half* dst = out + i * cpad;
    for (int ch = 0; ch < cpad; ++ch)
      dst[ch] = ch < c ? src[static_cast<long long>(ch) * h * w] : __float2half(0.f);
  }
}

// CLIP's MLP activation, in place: x * sigmoid(1.702 x)  (transformers QuickGELUActivation), fp32 math
__global__ void quick_gelu_kernel(uint4* __restrict__ x, long long nvec) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 v = x[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(h[e]);
      f.x = __fdividef(f.x, 1.0f + __expf(-1.702f * f.x));
      f.y = __fdividef(f.y, 1.0f + __expf(-1.702f * f.y));
      h[e] = __floats2half2_rn(f.x, f.y);
    }
    x[i] = v;
  }
}

// VAE mid-block attention (one head of 512 channels: beyond the flash kernel's head size, run as two GEMMs):
// in-place row softmax of the fp16 score matrix, fp32 math, one CTA per row (three passes over a row that
// stays in L1/L2: max, sum of exponentials, normalised write).
__global__ void softmax_rows_kernel(__half* __restrict__ x, long long ld, int cols) {
  __shared__ float s_red[32];
  __half* row = x + static_cast<long long>(blockIdx.x) * ld;
  const int nvec = cols >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  auto block_reduce = [&](float v, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? fmaxf(v, t) : v + t;
    }
    __syncthreads();           // s_red may still be read from the previous reduction
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float r = s_red[0];
    for (int w = 1; w < nwarp; ++w) r = is_max ? fmaxf(r, s_red[w]) : r + s_red[w];
    return r;
  };
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(row)[i];
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      mx = fmaxf(mx, fmaxf(f.x, f.y));
    }
  }
  mx = block_reduce(mx, true);
  float sum = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(row)[i];
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      sum += __expf(f.x - mx) + __expf(f.y - mx);
    }
  }
  sum = block_reduce(sum, false);
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    uint4 v = reinterpret_cast<const uint4*>(row)[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      h[e] = __floats2half2_rn(__expf(f.x - mx) * inv, __expf(f.y - mx) * inv);
    }
    reinterpret_cast<uint4*>(row)[i] = v;
  }
}

// im2col of a 3x3 convolution with `pad_lo` zero rows / columns in front and one behind (the VAE encoder's
// Downsample2D(padding=0) pads right / bottom only: pad_lo = 0, stride 2).  Same output layout as
// im2col3x3_kernel (misc_ops.cu): [nimg*ho*wo, kpadv vectors], column (kh*3+kw)*c + ci.
__global__ void im2col3x3_ex_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int nimg, int h, int w,
                                    int cv, int stride, int pad_lo, int ho, int wo, int kpadv) {
  const long long total = static_cast<long long>(nimg) * ho * wo * kpadv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kv = static_cast<int>(i % kpadv);
    const long long m = i / kpadv;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (kv < 9 * cv) {
      const int tap = kv / cv;
      const int v = kv % cv;
      const int ox = static_cast<int>(m % wo);
      const int oy = static_cast<int>((m / wo) % ho);
      const long long img = m / (static_cast<long long>(wo) * ho);
      const int iy = oy * stride + tap / 3 - pad_lo;
      const int ix = ox * stride + tap % 3 - pad_lo;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) val = x[((img * h + iy) * w + ix) * cv + v];
    }
    out[i] = val;
  }
}

__global__ void relu_kernel(uint4* __restrict__ x, long long nvec) {
  const __half2 zero = __float2half2_rn(0.f);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 v = x[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __hmax2(h[e], zero);
    x[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// MANModule: per (image, channel) statistics over the hw pixels, then
//   out = (x - mean) * rstd * (1 + gamma) + beta        (gamma, beta per pixel and channel)
// Deterministic like the GroupNorm pair: per-CTA fp32 partial sums in a workspace, fixed-order fp64
// combine in the apply kernel, no atomics.
// ------------------------------------------------------------------------------------------------
constexpr int MAN_MAX_CHUNKS = 32;

struct ManParams {
  const __half* x;
  const __half* gb;
  long long ldgb;
  int nimg, hw, C;
  float eps;
  __half* out;
  float* ws;  // [nimg, MAN_MAX_CHUNKS, 2, C] per-CTA partial (sum | sumsq) per channel
  int pix_per_cta, nchunks;
};

// blockDim = (V, vy): thread x owns channel vector x (8 channels), thread y strides over the pixels
__global__ void man_stats_kernel(const ManParams p) {
  extern __shared__ float s_part[];  // [vy][V][16]
  const int img = blockIdx.y;
  const int cv = threadIdx.x;
  const int V = blockDim.x, vy = blockDim.y;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  const int pbeg = blockIdx.x * p.pix_per_cta;
  const int pend = min(p.hw, pbeg + p.pix_per_cta);
  for (int px = pbeg + threadIdx.y; px < pend; px += vy) {
    const uint4 raw = *reinterpret_cast<const uint4*>(
        p.x + (static_cast<long long>(img) * p.hw + px) * p.C + cv * 8);
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
  float* mine = s_part + (threadIdx.y * V + cv) * 16;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mine[e] = s[e];
    mine[8 + e] = q[e];
  }
  __syncthreads();
  if (threadIdx.y == 0) {
    float* dst = p.ws + (static_cast<long long>(img) * MAN_MAX_CHUNKS + blockIdx.x) * 2 * p.C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = 0.f, b = 0.f;
      for (int y = 0; y < vy; ++y) {
        a += s_part[(y * V + cv) * 16 + e];
        b += s_part[(y * V + cv) * 16 + 8 + e];
      }
      dst[cv * 8 + e] = a;
      dst[p.C + cv * 8 + e] = b;
    }
  }
}

__global__ void man_apply_kernel(const ManParams p) {
  extern __shared__ float s_stat[];  // mean[C] | rstd[C]
  const int img = blockIdx.y;
  const int cv = threadIdx.x;
  const int V = blockDim.x, vy = blockDim.y;
  const int tid = threadIdx.y * V + cv;
  for (int c = tid; c < p.C; c += V * vy) {
    double sum = 0.0, sq = 0.0;
    for (int k = 0; k < p.nchunks; ++k) {
      const float* part = p.ws + (static_cast<long long>(img) * MAN_MAX_CHUNKS + k) * 2 * p.C;
      sum += static_cast<double>(part[c]);
      sq += static_cast<double>(part[p.C + c]);
    }
    const double mean = sum / p.hw;
    double var = sq / p.hw - mean * mean;  // biased variance, as nn.InstanceNorm2d uses
    if (var < 0.0) var = 0.0;
    s_stat[c] = static_cast<float>(mean);
    s_stat[p.C + c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
  }
  __syncthreads();
  float scale[8], shift[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    scale[e] = s_stat[p.C + cv * 8 + e];
    shift[e] = -s_stat[cv * 8 + e] * scale[e];
  }
  const int pbeg = blockIdx.x * p.pix_per_cta;
  const int pend = min(p.hw, pbeg + p.pix_per_cta);
  for (int px = pbeg + threadIdx.y; px < pend; px += vy) {
    const long long row = static_cast<long long>(img) * p.hw + px;
    const uint4 rx = *reinterpret_cast<const uint4*>(p.x + row * p.C + cv * 8);
    const uint4 rg = *reinterpret_cast<const uint4*>(p.gb + row * p.ldgb + cv * 8);
    const uint4 rb = *reinterpret_cast<const uint4*>(p.gb + row * p.ldgb + p.C + cv * 8);
    const __half2* hx = reinterpret_cast<const __half2*>(&rx);
    const __half2* hg = reinterpret_cast<const __half2*>(&rg);
    const __half2* hb = reinterpret_cast<const __half2*>(&rb);
    uint4 o;
    uint32_t* po = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fx = __half22float2(hx[e]);
      const float2 fg = __half22float2(hg[e]);
      const float2 fb = __half22float2(hb[e]);
      const float y0 = (fx.x * scale[2 * e] + shift[2 * e]) * (1.0f + fg.x) + fb.x;
      const float y1 = (fx.y * scale[2 * e + 1] + shift[2 * e + 1]) * (1.0f + fg.y) + fb.y;
      po[e] = pack_half2(y0, y1);
    }
    *reinterpret_cast<uint4*>(p.out + row * p.C + cv * 8) = o;
  }
}

}  // namespace mdk

extern "C" int mdk_cond_to_nhwc_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t ctot,
                                    int32_t c_first, int32_t c, int32_t h, int32_t w, int32_t ho,
                                    int32_t wo, int32_t cpad, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && x && out, "mdk_cond_to_nhwc_f16: null argument");
  MDK_REQUIRE(nimg > 0 && h > 0 && w > 0 && ho > 0 && wo > 0, "mdk_cond_to_nhwc_f16: empty problem");
  MDK_REQUIRE(c > 0 && c_first >= 0 && c_first + c <= ctot && cpad >= c && cpad % 8 == 0,
              "mdk_cond_to_nhwc_f16: bad channel slice [%d, %d) of %d, cpad=%d", c_first, c_first + c,
              ctot, cpad);
  const long long total = static_cast<long long>(nimg) * ho * wo;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(ctx->num_sms) * 8;
  if (blocks > cap) blocks = cap;
  cond_to_nhwc_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const __half*>(x), static_cast<__half*>(out), nimg, ctot, c_first, c, h, w, ho, wo,
      cpad);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_relu_f16(mdk_ctx* ctx, void* x, int64_t n, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && x, "mdk_relu_f16: null argument");
  MDK_REQUIRE(n % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "mdk_relu_f16: n must be a multiple of 8 and x 16-byte aligned");
  if (n <= 0) return 0;
  const long long nvec = n / 8;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(ctx->num_sms) * 8;
  if (blocks > cap) blocks = cap;
  relu_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<uint4*>(x), nvec);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_softmax_rows_f16(mdk_ctx* ctx, void* x, int64_t rows, int32_t cols, int64_t ld, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && x, "mdk_softmax_rows_f16: null argument");
  MDK_REQUIRE(cols > 0 && cols % 8 == 0 && ld % 8 == 0 && ld >= cols && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "mdk_softmax_rows_f16: cols=%d / ld=%lld must be multiples of 8, x 16-byte aligned", cols, (long long)ld);
  MDK_REQUIRE(rows <= 2147483647LL, "mdk_softmax_rows_f16: too many rows");
  if (rows <= 0) return 0;
  softmax_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(static_cast<__half*>(x), ld, cols);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_im2col3x3_ex_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t h, int32_t w,
                                    int32_t c, int32_t stride, int32_t pad_lo, int32_t kpad, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && x && out, "mdk_im2col3x3_ex_f16: null argument");
  MDK_REQUIRE(c % 8 == 0 && kpad % 8 == 0 && kpad >= 9 * c, "mdk_im2col3x3_ex_f16: c=%d kpad=%d", c, kpad);
  MDK_REQUIRE(stride >= 1 && (pad_lo == 0 || pad_lo == 1) && h + pad_lo >= 2 && w + pad_lo >= 2,
              "mdk_im2col3x3_ex_f16: bad stride / pad / size");
  const int ho = (h + pad_lo - 2) / stride + 1, wo = (w + pad_lo - 2) / stride + 1;
  const long long total = static_cast<long long>(nimg) * ho * wo * (kpad / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  im2col3x3_ex_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), nimg, h, w, c / 8, stride, pad_lo, ho, wo, kpad / 8);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mdk_quick_gelu_f16(mdk_ctx* ctx, void* x, int64_t n, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && x, "mdk_quick_gelu_f16: null argument");
  MDK_REQUIRE(n % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "mdk_quick_gelu_f16: n must be a multiple of 8 and x 16-byte aligned");
  if (n <= 0) return 0;
  const long long nvec = n / 8;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(ctx->num_sms) * 8;
  if (blocks > cap) blocks = cap;
  quick_gelu_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<uint4*>(x), nvec);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t mdk_man_ws_bytes(int32_t nimg, int32_t c) {
  return static_cast<int64_t>(nimg) * mdk::MAN_MAX_CHUNKS * 2 * c * sizeof(float);
}

extern "C" int mdk_man_modulate_f16(mdk_ctx* ctx, const mdk_man_args* a, void* stream_) {
  using namespace mdk;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MDK_REQUIRE(ctx && a && a->x && a->gb && a->out && a->ws, "mdk_man_modulate_f16: null argument");
  MDK_REQUIRE(a->c > 0 && a->c % 8 == 0 && a->c / 8 <= 512, "mdk_man_modulate_f16: c=%d unsupported", a->c);
  MDK_REQUIRE(a->nimg > 0 && a->hw > 0 && a->nimg <= 65535, "mdk_man_modulate_f16: empty problem");
  MDK_REQUIRE(a->ldgb >= 2 * a->c && a->ldgb % 8 == 0, "mdk_man_modulate_f16: ldgb=%lld < 2c or not %%8",
              (long long)a->ldgb);
  MDK_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->gb) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
              "mdk_man_modulate_f16: pointers must be 16-byte aligned");
  ManParams p;
  p.x = static_cast<const __half*>(a->x);
  p.gb = static_cast<const __half*>(a->gb);
  p.ldgb = a->ldgb;
  p.nimg = a->nimg;
  p.hw = a->hw;
  p.C = a->c;
  p.eps = a->eps;
  p.out = static_cast<__half*>(a->out);
  p.ws = static_cast<float*>(a->ws);
  const int V = a->c / 8;
  int vy = 512 / V;
  if (vy < 1) vy = 1;
  int chunks = (ctx->num_sms * 4 + a->nimg - 1) / a->nimg;
  const int max_chunks = (a->hw + vy - 1) / vy;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > MAN_MAX_CHUNKS) chunks = MAN_MAX_CHUNKS;
  if (chunks < 1) chunks = 1;
  p.pix_per_cta = (a->hw + chunks - 1) / chunks;
  p.nchunks = (a->hw + p.pix_per_cta - 1) / p.pix_per_cta;
  const dim3 block(V, vy);
  const dim3 grid(p.nchunks, a->nimg);
  const size_t stats_smem = static_cast<size_t>(V) * vy * 16 * sizeof(float);
  MDK_REQUIRE(stats_smem <= 48 * 1024, "mdk_man_modulate_f16: c=%d too large", a->c);
  man_stats_kernel<<<grid, block, stats_smem, stream>>>(p);
  count_launch();
  man_apply_kernel<<<grid, block, 2 * static_cast<size_t>(a->c) * sizeof(float), stream>>>(p);
  count_launch();
  MDK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
