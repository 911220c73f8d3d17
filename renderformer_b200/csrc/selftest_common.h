// Shared helpers for the standalone (torch-free) GPU self-test binaries.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/rfb200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static inline uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16, x *= 0x7feb352dU, x ^= x >> 15, x *= 0x846ca68bU, x ^= x >> 16;
  return x;
}
// deterministic value in [-1, 1)
static inline float hval(uint32_t seed, uint64_t i) {
  uint32_t h = hash_u32((uint32_t)(i * 2654435761ULL) ^ hash_u32(seed + (uint32_t)(i >> 32)));
  return (float)(h >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

static inline uint16_t f2h16(float f, int dtype) {
  if (dtype == RFB_BF16) {
    __nv_bfloat16 b = __float2bfloat16(f);
    return *reinterpret_cast<uint16_t*>(&b);
  }
  __half h = __float2half(f);
  return *reinterpret_cast<uint16_t*>(&h);
}
static inline float h162f(uint16_t u, int dtype) {
  if (dtype == RFB_BF16) {
    uint32_t w = (uint32_t)u << 16;
    float f;
    memcpy(&f, &w, 4);
    return f;
  }
  __half h = *reinterpret_cast<__half*>(&u);
  return __half2float(h);
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  explicit DevBuf(size_t n_) : n(n_) { CK(cudaMalloc(&p, n * sizeof(T))); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  void up(const std::vector<T>& h) { CK(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice)); }
  std::vector<T> down() const {
    std::vector<T> h(n);
    CK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }
  void zero() { CK(cudaMemset(p, 0, n * sizeof(T))); }
  void fill_byte(int b) { CK(cudaMemset(p, b, n * sizeof(T))); }
};

// random 16-bit matrix; returns the float values actually representable (for the reference)
static inline std::vector<uint16_t> rand16(size_t n, uint32_t seed, float scale, int dtype,
                                           std::vector<float>* as_float = nullptr) {
  std::vector<uint16_t> h(n);
  if (as_float) as_float->resize(n);
  for (size_t i = 0; i < n; ++i) {
    h[i] = f2h16(hval(seed, i) * scale, dtype);
    if (as_float) (*as_float)[i] = h162f(h[i], dtype);
  }
  return h;
}
static inline std::vector<float> rand32(size_t n, uint32_t seed, float scale) {
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = hval(seed, i) * scale;
  return h;
}

struct CheckResult {
  double max_abs = 0, max_ref = 0;
  size_t bad = 0, n = 0;
};

static int g_fail = 0;

// compare got vs ref with |d| <= atol + rtol*|ref|; print a diagnosis on failure
static inline bool report(const char* name, const std::vector<float>& got,
                          const std::vector<float>& ref, double atol, double rtol, int ncols = 0) {
  CheckResult r;
  r.n = ref.size();
  size_t first_bad[8];
  int nb = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    double d = fabs((double)got[i] - (double)ref[i]);
    if (!(d <= atol + rtol * fabs((double)ref[i]))) {  // catches NaN too
      if (nb < 8) first_bad[nb++] = i;
      r.bad++;
    }
    if (d > r.max_abs || d != d) r.max_abs = d;
    if (fabs(ref[i]) > r.max_ref) r.max_ref = fabs(ref[i]);
  }
  bool ok = r.bad == 0;
  printf("[%s] %-46s n=%zu max_abs_err=%.3e max|ref|=%.3e bad=%zu\n", ok ? "PASS" : "FAIL", name,
         r.n, r.max_abs, r.max_ref, r.bad);
  if (!ok) {
    g_fail++;
    for (int i = 0; i < nb; ++i) {
      size_t k = first_bad[i];
      if (ncols)
        printf("    bad[%zu] (row %zu, col %zu): got %.6f ref %.6f\n", k, k / ncols, k % ncols,
               got[k], ref[k]);
      else
        printf("    bad[%zu]: got %.6f ref %.6f\n", k, got[k], ref[k]);
    }
    if (ncols) {  // coarse error map: which (row%128 / 8) x (col%256 / 32) classes are wrong
      int map[16][8] = {};
      for (size_t i = 0; i < ref.size(); ++i) {
        double d = fabs((double)got[i] - (double)ref[i]);
        if (!(d <= atol + rtol * fabs((double)ref[i])))
          map[((i / ncols) % 128) / 8][((i % ncols) % 256) / 32]++;
      }
      printf("    error map rows(row%%128/8) x cols(col%%256/32):\n");
      for (int a = 0; a < 16; ++a) {
        printf("     ");
        for (int b = 0; b < 8; ++b) printf(" %7d", map[a][b]);
        printf("\n");
      }
    }
  }
  fflush(stdout);
  return ok;
}

struct GpuTimer {
  cudaEvent_t a, b;
  GpuTimer() {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~GpuTimer() {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
  void start(cudaStream_t s = 0) { cudaEventRecord(a, s); }
  float stop(cudaStream_t s = 0) {
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  }
};
