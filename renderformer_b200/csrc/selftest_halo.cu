// Experiment (not part of the product path): can the nine tap-shifted A operands of a 3x3 convolution be
// read straight out of ONE halo tile in shared memory by offsetting the UMMA descriptor?
// Halo tile: 18 rows x 16 pixel slots x 64 channels (fp16, 128 B per pixel, SWIZZLE_128B, written by one
// 4-D TMA box).  Output tile: 16 rows x 8 pixels = 128 accumulator rows; 8-row group g = image row y, group
// stride (SBO) = 16 pixels * 128 B = 2048 B, start = halo + ((dy*16 + dx) * 128) B.
// usage: selftest_halo [base_offset_mode]   0: descriptor base_offset = 0, 1: base_offset = dx
#include <stdlib.h>

#include <vector>

#include "host_util.h"
#include "ptx.cuh"
#include "selftest_common.h"

using namespace rfb;

__device__ __forceinline__ uint64_t desc_sw128_ofs(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

constexpr uint32_t kHalo = 18 * 16 * 128;      // 36864 B
constexpr uint32_t kBTap = 128 * 64 * 2;       // 16384 B per tap
constexpr uint32_t kSmem = kHalo + 9 * kBTap + 1024 + 64;

__global__ void __launch_bounds__(128, 1)
    halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sH = smem;
  uint8_t* sB = smem + kHalo;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 9 * kBTap);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1), mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kHalo + 9 * kBTap);
    tma_load_4d(sH, &tmX, bar, 0, 0, 0, 0);
    for (int t = 0; t < 9; ++t) tma_load_2d(sB + t * kBTap, &tmW, bar, t * 64, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(0u, 128, 128);
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3, dx = t % 3;
      const uint32_t a0 = smem_u32(sH) + (dy * 16 + dx) * 128;
      const uint64_t ad = desc_sw128_ofs(a0, 2048, mode == 1 ? dx : 0);
      const uint64_t bd = umma_desc_sw128(smem_u32(sB + t * kBTap));
      for (int k = 0; k < 4; ++k) umma_f16(tm, ad + 2 * k, bd + 2 * k, idesc, (t | k) != 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) out[r * 128 + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 128);
  }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int H = 18, W = 16, C = 64, N = 128, K = 9 * C;
  std::vector<uint16_t> hX = rand16((size_t)H * W * C, 3, 1.0f, RFB_F16), hW = rand16((size_t)N * K, 4, 0.1f, RFB_F16);
  DevBuf<uint16_t> dX(hX.size()), dW(hW.size());
  dX.up(hX), dW.up(hW);
  DevBuf<float> dout(128 * 128);
  dout.fill_byte(0);
  CUtensorMap tmX, tmW;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, 1};
    uint64_t st[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, 16, 18, 1};
    if (make_tmap_16b(&tmX, RFB_F16, dX.p, 4, dims, st, box) != RFB_OK) return 2;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t st[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, 128};
    if (make_tmap_16b(&tmW, RFB_F16, dW.p, 2, dims, st, box) != RFB_OK) return 2;
  }
  CK(cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  halo_kernel<<<1, 128, kSmem>>>(tmX, tmW, dout.p, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[FAIL] halo mode %d: %s\n", mode, cudaGetErrorString(e));
    return 3;
  }
  std::vector<float> got = dout.down(), ref(128 * 128, 0.f);
  for (int y = 0; y < 16; ++y)
    for (int x = 0; x < 8; ++x)
      for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (int t = 0; t < 9; ++t)
          for (int c = 0; c < C; ++c)
            s += h162f(hX[((size_t)(y + t / 3) * W + (x + t % 3)) * C + c], RFB_F16) * h162f(hW[(size_t)n * K + t * C + c], RFB_F16);
        ref[(y * 8 + x) * 128 + n] = s;
      }
  char name[64];
  snprintf(name, sizeof(name), "halo-shifted UMMA descriptors, base_offset mode %d", mode);
  report(name, got, ref, 2e-3, 2e-3, 128);
  return g_fail ? 1 : 0;
}
