#include <stdlib.h>

#include "host_common.h"

#include <mutex>
#include <string.h>

#include "../../include/mdk.h"

namespace mdk {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                        : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(
        "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] strides [%llu %llu "
        "%llu] box [%u %u %u %u] base %p",
        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 1 ? strides_bytes[1] : 0),
        (unsigned long long)(rank > 2 ? strides_bytes[2] : 0),
        (unsigned long long)(rank > 3 ? strides_bytes[3] : 0), box[0], rank > 1 ? box[1] : 0,
        rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
  }
  return 0;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MDK_PDL");
    on = e ? (atoi(e) != 0) : 1;
  }
  return on != 0;
}

}  // namespace mdk

extern "C" {

int mdk_create(int device, mdk_ctx** out) {
  if (!out) return mdk::set_error("mdk_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return mdk::set_error("mdk_create: no CUDA device (%s)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return mdk::set_error("mdk_create: bad device %d", device);
  cudaDeviceProp prop;
  MDK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return mdk::set_error("mdk_create: device %d is sm_%d%d; this library is sm_100a only", device,
                          prop.major, prop.minor);
  mdk_ctx* c = new mdk_ctx;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = c;
  return 0;
}

void mdk_destroy(mdk_ctx* ctx) { delete ctx; }

const char* mdk_last_error(void) { return mdk::g_err; }

int mdk_abi_version(void) { return MDK_ABI_VERSION; }

int64_t mdk_launch_count(void) { return (int64_t)mdk::g_launch_count.load(); }

}  // extern "C"
