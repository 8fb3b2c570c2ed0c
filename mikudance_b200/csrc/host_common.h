// Host-side helpers shared by the C-ABI entry points: context, error string, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

struct mdk_ctx {
  int device;
  int num_sms;
  int max_smem_optin;
};

namespace mdk {

int set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;
inline void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

#define MDK_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return mdk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                    \
  } while (0)

#define MDK_REQUIRE(cond, ...)                       \
  do {                                               \
    if (!(cond)) return mdk::set_error(__VA_ARGS__); \
  } while (0)

// Encode a tiled fp16 tensor map (rank 2..5). dims/strides innermost first; strides in BYTES for
// dims 1..rank-1 (dim 0 is dense). 128-byte swizzle, zero fill out of bounds.
int encode_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);

}  // namespace mdk
