// Bandwidth-bound helpers of the DPT patch decoder (layers/dpt.py), NHWC fp16 activations:
// ConvTranspose(k = s) pixel-shuffle scatter, im2col for the stride-2 3x3 conv, and bilinear
// align_corners=True resampling.  All convolutions themselves run through rfb_gemm.
#include <stdlib.h>

#include <atomic>

#include "host_util.h"
#include "ptx.cuh"

namespace rfb {

extern std::atomic<long long> g_launch_count;

// in [B*h*w, s*s*C] with column = (i*s + j)*C + c  ->  out NHWC [B, h*s, w*s, C]
//   (ConvTranspose2d kernel == stride: layers/dpt.py:195-206; the GEMM adds the bias)
__global__ void pixel_shuffle_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int h,
                                     int w, int s, int C8) {
  const long long total = (long long)B * h * w * s * s * C8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = t % C8;
    long long r = t / C8;
    const int j = r % s;
    r /= s;
    const int i = r % s;
    r /= s;
    const int x = r % w;
    r /= w;
    const int y = r % h;
    const int b = r / h;
    const long long o = (((long long)b * h * s + (y * s + i)) * (w * s) + (x * s + j)) * C8 + c;
    out[o] = __ldg(in + t);
  }
}

// in NHWC [B,H,W,C] -> out [B*Ho*Wo, 9*C] (tap-major, channel-minor), kernel 3, stride 2, pad 1
//   (resize_layers[3]: layers/dpt.py:208-213)
__global__ void im2col_s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                                 int C8, int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo * 9 * C8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = t % C8;
    long long r = t / C8;
    const int tap = r % 9;
    r /= 9;
    const int xo = r % Wo;
    r /= Wo;
    const int yo = r % Ho;
    const int b = r / Ho;
    const int y = yo * 2 - 1 + tap / 3, x = xo * 2 - 1 + tap % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(in + (((long long)b * H + y) * W + x) * C8 + c);
    out[t] = v;
  }
}

__device__ __forceinline__ void h8_to_f(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}

// bilinear, align_corners=True: src = dst * (in-1)/(out-1)   (F.interpolate, layers/dpt.py:154-155)
__global__ void upsample_bilinear_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int Hi,
                                         int Wi, int Ho, int Wo, int C8) {
  const long long total = (long long)B * Ho * Wo * C8;
  const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = t % C8;
    long long r = t / C8;
    const int xo = r % Wo;
    r /= Wo;
    const int yo = r % Ho;
    const int b = r / Ho;
    const float fy = yo * sy, fx = xo * sx;
    int y0 = (int)fy, x0 = (int)fx;
    y0 = min(y0, Hi - 1), x0 = min(x0, Wi - 1);
    const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
    const float wy = fy - y0, wx = fx - x0;
    const uint4* base = in + (long long)b * Hi * Wi * C8 + c;
    float a[8], bb[8], cc[8], d[8];
    h8_to_f(__ldg(base + ((long long)y0 * Wi + x0) * C8), a);
    h8_to_f(__ldg(base + ((long long)y0 * Wi + x1) * C8), bb);
    h8_to_f(__ldg(base + ((long long)y1 * Wi + x0) * C8), cc);
    h8_to_f(__ldg(base + ((long long)y1 * Wi + x1) * C8), d);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float top = a[i] + (bb[i] - a[i]) * wx;
      const float bot = cc[i] + (d[i] - cc[i]) * wx;
      o[i] = top + (bot - top) * wy;
    }
    uint4 u;
    u.x = pack_f16(o[0], o[1]), u.y = pack_f16(o[2], o[3]), u.z = pack_f16(o[4], o[5]), u.w = pack_f16(o[6], o[7]);
    out[t] = u;
  }
}

// Tiled variant for ~2x magnification of 128-channel maps (every FeatureFusionBlock of the DPT head at the released
// resolutions).  The kernel above fetches four 16-byte taps per 16 bytes written, all through L1/L2: it runs at the
// L2 -> SM rate, four times the HBM traffic.  Here a block owns an 8 x 32 pixel output tile: the <= 6 x 18 input
// pixels underneath are staged in shared memory once (coalesced 16-byte loads), a thread walks one (column,
// 8-channel group) of the tile downwards keeping the horizontally interpolated input rows y0 / y0+1 in
// registers (two shared-memory reads per NEW input row instead of four per output), and every warp store is
// 512 contiguous bytes.  Arithmetic identical to the kernel above (same expressions, same rounding).
constexpr int kUpTH = 8, kUpTW = 32, kUpRows = 6, kUpCols = 18, kUpC8 = 16;

__global__ void __launch_bounds__(256)
    upsample_bilinear_tiled_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int Hi, int Wi, int Ho, int Wo) {
  __shared__ uint4 tile[kUpRows * kUpCols * kUpC8];  // 27 KB
  const int b = blockIdx.z;
  const int Y0 = blockIdx.y * kUpTH, X0 = blockIdx.x * kUpTW;
  const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const int iy0 = min((int)(Y0 * sy), Hi - 1), ix0 = min((int)(X0 * sx), Wi - 1);
  const int ylast = min(Y0 + kUpTH, Ho) - 1, xlast = min(X0 + kUpTW, Wo) - 1;
  const int rows = min(min((int)(ylast * sy), Hi - 1) + 1, Hi - 1) - iy0 + 1;  // <= kUpRows (host-checked scale)
  const int cols = min(min((int)(xlast * sx), Wi - 1) + 1, Wi - 1) - ix0 + 1;  // <= kUpCols
  const uint4* src = in + (long long)b * Hi * Wi * kUpC8;
  for (int i = threadIdx.x; i < rows * cols * kUpC8; i += 256) {
    const int c = i % kUpC8, px = (i / kUpC8) % cols, py = i / (kUpC8 * cols);
    tile[(py * kUpCols + px) * kUpC8 + c] = __ldg(src + ((long long)(iy0 + py) * Wi + ix0 + px) * kUpC8 + c);
  }
  __syncthreads();
  uint4* dst = out + (long long)b * Ho * Wo * kUpC8;
#pragma unroll 1
  for (int u = threadIdx.x; u < kUpTW * kUpC8; u += 256) {  // (column, channel group) strips of the tile
    const int c = u % kUpC8, xo = X0 + u / kUpC8;
    if (xo >= Wo) continue;
    const float fx = xo * sx;
    const int x0 = min((int)fx, Wi - 1), x1 = min(x0 + 1, Wi - 1);
    const float wx = fx - x0;
    auto hrow = [&](int y, float (&h)[8]) {  // horizontally interpolated input row y at this column
      float a[8], bb[8];
      h8_to_f(tile[((y - iy0) * kUpCols + (x0 - ix0)) * kUpC8 + c], a);
      h8_to_f(tile[((y - iy0) * kUpCols + (x1 - ix0)) * kUpC8 + c], bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = a[i] + (bb[i] - a[i]) * wx;
    };
    float top[8], bot[8];
    int ytop = -1, ybot = -1;
    for (int yo = Y0; yo <= ylast; ++yo) {
      const float fy = yo * sy;
      const int y0 = min((int)fy, Hi - 1), y1 = min(y0 + 1, Hi - 1);
      const float wy = fy - y0;
      if (y0 != ytop) {
        if (y0 == ybot) {
#pragma unroll
          for (int i = 0; i < 8; ++i) top[i] = bot[i];
        } else {
          hrow(y0, top);
        }
        ytop = y0;
      }
      if (y1 != ybot) {
        if (y1 == ytop) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bot[i] = top[i];
        } else {
          hrow(y1, bot);
        }
        ybot = y1;
      }
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = top[i] + (bot[i] - top[i]) * wy;
      uint4 v;
      v.x = pack_f16(o[0], o[1]), v.y = pack_f16(o[2], o[3]), v.z = pack_f16(o[4], o[5]), v.w = pack_f16(o[6], o[7]);
      dst[((long long)yo * Wo + xo) * kUpC8 + c] = v;
    }
  }
}

static inline int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 148LL * 32;
  return (int)(b < cap ? b : cap);
}

// ---------------------------------------------------------------------------------------------
// HDR -> 8-bit LDR (the step right after the path: infer.py:94-98, batch_infer.py:153-157).
// mode 0 ('none'): uint8(clip(x, 0, 1) * 255) -- the CLIs' default, reproduced bit-exactly (fp32
//                  multiply, truncation like numpy's astype).
// mode 1: Khronos PBR Neutral tone curve (published reference formula) followed by the sRGB OETF.
//         The reference reaches it through simple_ocio / OpenColorIO, which is not vendored: unpinned.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float srgb_oetf(float x) {
  x = fminf(fmaxf(x, 0.f), 1.f);
  return x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
}

__global__ void ldr_quantize_kernel(const float* __restrict__ hdr, uint8_t* __restrict__ out, long long n_px,
                                    int mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  float r = hdr[i * 3], g = hdr[i * 3 + 1], b = hdr[i * 3 + 2];
  if (mode == 1) {
    const float start = 0.8f - 0.04f, desat = 0.15f;
    const float x = fminf(r, fminf(g, b));
    const float off = x < 0.08f ? x - 6.25f * x * x : 0.04f;
    r -= off, g -= off, b -= off;
    const float peak = fmaxf(r, fmaxf(g, b));
    if (peak >= start) {
      const float d = 1.f - start;
      const float np_ = 1.f - d * d / (peak + d - start);
      const float sc = np_ / peak;
      r *= sc, g *= sc, b *= sc;
      const float t = 1.f - 1.f / (desat * (peak - np_) + 1.f);
      r = r + (np_ - r) * t, g = g + (np_ - g) * t, b = b + (np_ - b) * t;
    }
    r = srgb_oetf(r), g = srgb_oetf(g), b = srgb_oetf(b);
  }
  // NaN clips to 0 like np.clip(...).astype(uint8) does for the CLIs' finite outputs; fp32 product, truncation
  out[i * 3] = (uint8_t)(__fmul_rn(fminf(fmaxf(r, 0.f), 1.f), 255.f));
  out[i * 3 + 1] = (uint8_t)(__fmul_rn(fminf(fmaxf(g, 0.f), 1.f), 255.f));
  out[i * 3 + 2] = (uint8_t)(__fmul_rn(fminf(fmaxf(b, 0.f), 1.f), 255.f));
}

}  // namespace rfb

using namespace rfb;

extern "C" int rfb_pixel_shuffle(const void* in, void* out, int B, int h, int w, int s, int C,
                                 rfb_stream_t stream) {
  if (!in || !out || C % 8 || s < 1) return RFB_ERR_ARG;
  const long long total = (long long)B * h * w * s * s * (C / 8);
  pixel_shuffle_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, h, w,
                                                                          s, C / 8);
  g_launch_count++;
  return check_launch("pixel_shuffle_kernel");
}

extern "C" int rfb_im2col_s2(const void* in, void* out, int B, int H, int W, int C, rfb_stream_t stream) {
  if (!in || !out || C % 8) return RFB_ERR_ARG;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)B * Ho * Wo * 9 * (C / 8);
  im2col_s2_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, H, W,
                                                                      C / 8, Ho, Wo);
  g_launch_count++;
  return check_launch("im2col_s2_kernel");
}

extern "C" int rfb_upsample_bilinear(const void* in, void* out, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                     rfb_stream_t stream) {
  if (!in || !out || C % 8) return RFB_ERR_ARG;
  {
    // tiled kernel: 128 channels, magnification >= ~1.9 in both directions (the tile's input footprint must fit)
    static int tiled_on = -1;
    if (tiled_on < 0) {
      const char* e = getenv("RFB_UPSAMPLE_TILED");
      tiled_on = (e && e[0] == '0') ? 0 : 1;
    }
    const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
    const float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    const bool fits = (int)((kUpTH - 1) * sy) + 3 <= kUpRows && (int)((kUpTW - 1) * sx) + 3 <= kUpCols;
    if (tiled_on && C == 8 * kUpC8 && fits && B <= 65535 && Ho >= kUpTH && Wo >= kUpTW) {
      dim3 grid((Wo + kUpTW - 1) / kUpTW, (Ho + kUpTH - 1) / kUpTH, B);
      upsample_bilinear_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, Hi, Wi, Ho, Wo);
      g_launch_count++;
      return check_launch("upsample_bilinear_tiled_kernel");
    }
  }
  const long long total = (long long)B * Ho * Wo * (C / 8);
  upsample_bilinear_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B,
                                                                              Hi, Wi, Ho, Wo, C / 8);
  g_launch_count++;
  return check_launch("upsample_bilinear_kernel");
}

extern "C" int rfb_ldr_quantize(const float* hdr, uint8_t* out, long long n_pixels, int mode, rfb_stream_t stream) {
  if (!hdr || !out || n_pixels <= 0 || (mode != 0 && mode != 1)) return RFB_ERR_ARG;
  ldr_quantize_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hdr, out, n_pixels, mode);
  g_launch_count++;
  return check_launch("ldr_quantize_kernel");
}
