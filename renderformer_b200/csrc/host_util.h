// Host-side helpers shared by the launchers: error codes, TMA descriptor encoding
// through the driver entry point (no link-time dependency on libcuda, so the
// library also loads on a CPU-only box for the symbol-export test).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rfb200.h"

namespace rfb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 16-bit element tensor map, 128-byte swizzle, zero OOB fill.
// dims/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
inline int make_tmap_16b(CUtensorMap* out, int dtype /*RFB_BF16|RFB_F16*/, const void* base,
                         int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return RFB_ERR_DRIVER;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return RFB_ERR_ALIGN;
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16 != 0) return RFB_ERR_ALIGN;
  CUresult r = fn(out, dtype == RFB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                        : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "rfb: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,..]\n",
            (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0));
    return RFB_ERR_TMAP;
  }
  return RFB_OK;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "rfb: launch of %s failed: %s\n", what, cudaGetErrorString(e));
    return RFB_ERR_LAUNCH;
  }
  return RFB_OK;
}

// Launch-side caches are kept PER DEVICE: one process may drive several GPUs (a model on cuda:1 while
// cuda:0 is current elsewhere), and cudaFuncSetAttribute / the SM count belong to the device.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
struct PerDeviceFlag {
  bool v[kMaxDevices] = {};
  bool& get() { return v[current_device()]; }
};

inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

}  // namespace rfb
