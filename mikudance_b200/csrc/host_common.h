// Host-side helpers shared by the C-ABI entry points: context, error string, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <utility>

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

// Programmatic dependent launch (PDL): the kernels of the denoising step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, execute `griddepcontrol.wait` before their first global
// memory access and `griddepcontrol.launch_dependents` right after it (ptx.cuh: pdl_wait / pdl_trigger).  The next
// kernel of the stream is then scheduled while this one still runs — its launch latency and prologue (barrier
// init, TMEM allocation, tensor-map prefetch, smem zeroing) overlap the tail of its predecessor — and blocks in
// hardware at its own wait until the predecessor has completed and flushed.  A step is ~750 launches of 20-400 us
// kernels (22 ms per step on 8 GPUs): the 2-3 us between dependent kernels is what PDL removes.  Captured by CUDA
// graphs as programmatic edges.  MDK_PDL=0 launches everything fully serialised.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Encode a tiled fp16 tensor map (rank 2..5). dims/strides innermost first; strides in BYTES for
// dims 1..rank-1 (dim 0 is dense). 128-byte swizzle, zero fill out of bounds.
int encode_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);

}  // namespace mdk
