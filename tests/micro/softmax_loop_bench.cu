// Microbenchmark (diagnostic, not product): the instruction mix of the attention kernel's softmax inner
// loop — per 128-score row tile: 128 x (FFMA, MUFU.EX2), 64 x F2FP pack, 16 x STS.128 (+ optional 64 x
// FMNMX3 row max) — run back to back with NO barriers / TMEM / MMA around it, for 1..4 warps per SM
// sub-partition.  It answers: how many SM cycles does one 128 x 128 softmax tile cost when only the
// pipes are in the way?  (attn_tc_kernel<1,128,2,1> measures ~3400 cycles per tile pair per SMSP.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/softmax_loop_bench tests/micro/softmax_loop_bench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// the attention kernel's half2 polynomial exp2 (csrc/attn_tc.cu::ex2_poly_h2)
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
__device__ __forceinline__ uint32_t pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// MODE 0: FFMA + EX2 + pack + STS (the kernel's mix)   1: + row max (FMNMX)   2: no STS   3: EX2 only (sum)
// MODE 4: like 0 but every second pair on the FMA pipe (degree-3 polynomial, float)   5: every fourth pair (25 %)
// MODE 6: every second pair through the kernel's half2 polynomial (ex2_poly_h2)   7: every fourth pair, half2
template <int MODE>
__global__ void __launch_bounds__(384, 1) loop_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       long long* __restrict__ cycles, int tiles, float scale) {
  extern __shared__ uint4 sm[];
  const int row = threadIdx.x;
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = in[(blockIdx.x * blockDim.x + row) * 128 + i];
  float m = 3.0f, acc = 0.f;
  uint32_t last_bits = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    float mx[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = v[q * 8 + 2 * e], b = v[q * 8 + 2 * e + 1];
        if (MODE == 1) mx[e] = fmaxf(mx[e], fmaxf(a, b));
        const float x0 = fmaf(a, scale, -m), x1 = fmaf(b, scale, -m);
        float p0, p1;
        if ((MODE == 6 && (e & 1)) || (MODE == 7 && e == 3)) {
          pk[e] = ex2_poly_h2(x0, x1);
          continue;
        }
        if ((MODE == 4 && (e & 1)) || (MODE == 5 && e == 3)) {
          // 2^x = 2^n * poly(f): magic-number rounding on the FMA pipe, exponent added with integer ops
          const float j0 = x0 + 12582912.f, j1 = x1 + 12582912.f;
          const float f0 = x0 - (j0 - 12582912.f), f1 = x1 - (j1 - 12582912.f);
          float q0 = fmaf(0.05550411f, f0, 0.24022651f), q1 = fmaf(0.05550411f, f1, 0.24022651f);
          q0 = fmaf(q0, f0, 0.69314718f); q1 = fmaf(q1, f1, 0.69314718f);
          q0 = fmaf(q0, f0, 1.0f); q1 = fmaf(q1, f1, 1.0f);
          p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(j0) << 23));
          p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(j1) << 23));
        } else {
          p0 = ex2(x0);
          p1 = ex2(x1);
        }
        if (MODE == 3) acc += p0 + p1;
        pk[e] = pack(p0, p1);
      }
      last_bits = pk[3];
      if (MODE != 3) {
        if (MODE != 2)
          sm[row * 16 + (q ^ (row & 7))] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        else
          acc += __uint_as_float(pk[0] ^ pk[1] ^ pk[2] ^ pk[3]);
      }
    }
    if (MODE == 1) m = fminf(m, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * 1e-30f + m);
    // loop-carried through a value the compiler cannot fold (the first run of this benchmark updated m
    // with `m += 1e-7f`, which constant-folds to m: every mode but "row max" then had a loop-invariant body
    // that was hoisted / dead-store-eliminated — only the row-max rows of profiles/r01_softmax_loop_bench.log
    // are valid measurements)
    m = fmaf(__uint_as_float(last_bits), 1e-30f, m);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + row] = acc + m + __uint_as_float(sm[row * 16].x);
}

template <int MODE>
static void run(const char* name, const float* in, float* out, long long* cyc, int nsm) {
  const int tiles = 400;
  for (int warps = 4; warps <= 12; warps += 4) {
    const int threads = warps * 32;
    const size_t smem = (size_t)threads * 256;
    cudaFuncSetAttribute(loop_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    loop_kernel<MODE><<<nsm, threads, smem>>>(in, out, cyc, tiles, 0.01f);
    loop_kernel<MODE><<<nsm, threads, smem>>>(in, out, cyc, tiles, 0.01f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; ++i) avg += (double)h[i];
    avg /= nsm;
    const double per_tile_smsp = avg / tiles;   // cycles for one tile of EVERY warp of an SMSP (warps/4 of them)
    printf("%-28s warps/SMSP=%d  cycles per 128x128 tile-set per SMSP = %7.0f  -> per warp-tile %6.0f  (exp/clk/SM %.1f)\n",
           name, warps / 4, per_tile_smsp, per_tile_smsp / (warps / 4), 128.0 * threads / per_tile_smsp);
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, sizeof(float) * nsm * 384 * 128);
  cudaMalloc(&out, sizeof(float) * nsm * 384);
  cudaMalloc(&cyc, sizeof(long long) * 256);
  cudaMemset(in, 0, sizeof(float) * nsm * 384 * 128);
  printf("%s, %d SMs\n", prop.name, nsm);
  run<3>("EX2 only (+FFMA, sum)", in, out, cyc, nsm);
  run<2>("FFMA+EX2+F2FP (no STS)", in, out, cyc, nsm);
  run<0>("FFMA+EX2+F2FP+STS (kernel)", in, out, cyc, nsm);
  run<1>("kernel mix + row max", in, out, cyc, nsm);
  run<4>("kernel mix, 50% poly on FMA", in, out, cyc, nsm);
  run<5>("kernel mix, 25% poly on FMA", in, out, cyc, nsm);
  run<6>("kernel mix, 50% half2 poly", in, out, cyc, nsm);
  run<7>("kernel mix, 25% half2 poly", in, out, cyc, nsm);
  return 0;
}
