// Standalone GPU self-test + micro-benchmark of rfb_gemm (no torch).
// Checks every epilogue / A-operand mode against a naive CUDA-core reference kernel.
#include <stdlib.h>
#include <string.h>

#include "selftest_common.h"

__global__ void ref_linear(const uint16_t* A, long long lda, const uint16_t* W, long long ldw,
                           float* acc, int M, int N, int K, int bf16) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    uint16_t a = A[m * lda + k], w = W[n * ldw + k];
    float fa, fw;
    if (bf16) {
      fa = __uint_as_float((uint32_t)a << 16), fw = __uint_as_float((uint32_t)w << 16);
    } else {
      fa = __half2float(*reinterpret_cast<__half*>(&a));
      fw = __half2float(*reinterpret_cast<__half*>(&w));
    }
    s = fmaf(fa, fw, s);
  }
  acc[(long long)m * N + n] = s;
}

__global__ void ref_conv3(const uint16_t* X, const uint16_t* W, float* acc, int B, int H, int Wd,
                          int Cin, int Cout, int bf16) {
  int co = blockIdx.x * blockDim.x + threadIdx.x;
  long long pix = blockIdx.y;
  if (co >= Cout) return;
  int x = pix % Wd, y = (pix / Wd) % H, b = pix / ((long long)Wd * H);
  float s = 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= Wd) continue;
    const uint16_t* xp = X + (((long long)b * H + yy) * Wd + xx) * Cin;
    const uint16_t* wp = W + (long long)co * 9 * Cin + (long long)tap * Cin;
    for (int c = 0; c < Cin; ++c) {
      uint16_t a = xp[c], w = wp[c];
      float fa, fw;
      if (bf16) {
        fa = __uint_as_float((uint32_t)a << 16), fw = __uint_as_float((uint32_t)w << 16);
      } else {
        fa = __half2float(*reinterpret_cast<__half*>(&a));
        fw = __half2float(*reinterpret_cast<__half*>(&w));
      }
      s = fmaf(fa, fw, s);
    }
  }
  acc[pix * Cout + co] = s;
}

static float silu_h(float x) { return x / (1.0f + expf(-x)); }

struct Case {
  const char* name;
  int M, N, K, dtype, a_mode, B, H, Wd, Cin, epi, out_dtype;
  bool bias, res1, res2, act, rowmap;
  int res_dtype, bn;
};

static std::vector<float> to_float(const void* dev, size_t n, int dtype) {
  std::vector<float> out(n);
  if (dtype == RFB_F32) {
    CK(cudaMemcpy(out.data(), dev, n * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint16_t> h(n);
    CK(cudaMemcpy(h.data(), dev, n * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) out[i] = h162f(h[i], dtype);
  }
  return out;
}

static void run_case(const Case& c) {
  const int M = c.M, N = c.N, K = c.K;
  const long long lda = (c.a_mode == RFB_A_LINEAR) ? K : c.Cin;
  size_t a_elems = (c.a_mode == RFB_A_LINEAR) ? (size_t)M * K : (size_t)c.B * c.H * c.Wd * c.Cin;
  std::vector<uint16_t> hA = rand16(a_elems, 11, 1.0f, c.dtype);
  std::vector<uint16_t> hW = rand16((size_t)N * K, 22, 0.05f, c.dtype);
  DevBuf<uint16_t> dA(a_elems), dW((size_t)N * K);
  dA.up(hA), dW.up(hW);
  DevBuf<float> dacc((size_t)M * N);
  if (c.a_mode == RFB_A_LINEAR) {
    dim3 g((N + 127) / 128, M);
    ref_linear<<<g, 128>>>(dA.p, lda, dW.p, K, dacc.p, M, N, K, c.dtype == RFB_BF16);
  } else {
    dim3 g((N + 63) / 64, M);
    ref_conv3<<<g, 64>>>(dA.p, dW.p, dacc.p, c.B, c.H, c.Wd, c.Cin, N, c.dtype == RFB_BF16);
  }
  CK(cudaDeviceSynchronize());
  std::vector<float> acc = dacc.down();

  // epilogue inputs
  std::vector<float> hbias = rand32(N, 33, 0.5f);
  DevBuf<float> dbias(N);
  dbias.up(hbias);
  const int out_cols = (c.epi == RFB_EPI_SWIGLU) ? N / 2 : (c.epi == RFB_EPI_FINAL ? 3 : N);
  const long long ldo = (c.epi == RFB_EPI_STORE) ? ((out_cols + 7) & ~7) : out_cols;
  const int out_rows = M + (c.rowmap ? 64 : 0);  // row_map scatters into a larger buffer
  std::vector<int> hmap(M);
  for (int m = 0; m < M; ++m) hmap[m] = c.rowmap ? (int)(((long long)m * 7919) % M) + 64 * ((m % 2)) : m;
  if (c.rowmap) {  // make it a true permutation of [0,M) shifted by 13
    for (int m = 0; m < M; ++m) hmap[m] = (int)((m + 13) % M);
  }
  DevBuf<int> dmap(M);
  dmap.up(hmap);
  const size_t res_n = (size_t)out_rows * ldo;
  std::vector<float> hres1f, hres2f;
  std::vector<float> r1 = rand32(res_n, 44, 1.0f), r2 = rand32(res_n, 55, 1.0f);
  DevBuf<float> dres1f(res_n), dres2f(res_n);
  DevBuf<uint16_t> dres1h(res_n), dres2h(res_n);
  if (c.res_dtype == RFB_F32) {
    dres1f.up(r1), dres2f.up(r2);
  } else {
    std::vector<uint16_t> h1(res_n), h2(res_n);
    for (size_t i = 0; i < res_n; ++i) {
      h1[i] = f2h16(r1[i], c.res_dtype), r1[i] = h162f(h1[i], c.res_dtype);
      h2[i] = f2h16(r2[i], c.res_dtype), r2[i] = h162f(h2[i], c.res_dtype);
    }
    dres1h.up(h1), dres2h.up(h2);
  }
  std::vector<float> hw2 = rand32(96, 66, 0.3f), hb2 = rand32(3, 77, 0.2f);
  DevBuf<float> dw2(96), db2(3);
  dw2.up(hw2), db2.up(hb2);

  const size_t out_n = (size_t)out_rows * ldo;
  const int osz = (c.out_dtype == RFB_F32) ? 4 : 2;
  DevBuf<uint8_t> dout(out_n * osz), dact(out_n * osz);
  dout.fill_byte(0), dact.fill_byte(0);

  rfb_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.M = M, a.N = N, a.K = K, a.A = dA.p, a.lda = lda, a.W = dW.p, a.ldw = K, a.dtype = c.dtype;
  a.a_mode = c.a_mode, a.B = c.B, a.H = c.H, a.Wd = c.Wd, a.Cin = c.Cin, a.epi = c.epi;
  a.bias = (c.bias || c.epi == RFB_EPI_FINAL) ? dbias.p : nullptr;
  a.res_dtype = c.res_dtype, a.ldres = ldo;
  if (c.res1) a.res1 = (c.res_dtype == RFB_F32) ? (void*)dres1f.p : (void*)dres1h.p;
  if (c.res2) a.res2 = (c.res_dtype == RFB_F32) ? (void*)dres2f.p : (void*)dres2h.p;
  a.out = dout.p, a.out_dtype = c.out_dtype, a.ldo = ldo;
  a.out_act = c.act ? dact.p : nullptr;
  a.row_map = c.rowmap ? dmap.p : nullptr;
  a.w2 = dw2.p, a.b2 = db2.p, a.bn_override = c.bn;
  int rc = rfb_gemm(&a, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc != RFB_OK || e != cudaSuccess) {
    printf("[FAIL] %-46s rc=%d cuda=%s\n", c.name, rc, cudaGetErrorString(e));
    g_fail++;
    if (e != cudaSuccess) exit(3);  // context is gone
    return;
  }
  std::vector<float> got = to_float(dout.p, out_n, c.out_dtype);
  std::vector<float> got_act = to_float(dact.p, out_n, c.out_dtype);

  // expected
  std::vector<float> exp(out_n, 0.f), exp_act(out_n, 0.f);
  for (int m = 0; m < M; ++m) {
    const long long orow = hmap[m];
    if (c.epi == RFB_EPI_STORE) {
      for (int n = 0; n < N; ++n) {
        float v = acc[(size_t)m * N + n];
        if (c.bias) v += hbias[n];
        if (c.res1) v += r1[orow * ldo + n];
        if (c.res2) v += r2[orow * ldo + n];
        exp[orow * ldo + n] = v;
        exp_act[orow * ldo + n] = silu_h(v);
      }
    } else if (c.epi == RFB_EPI_SWIGLU) {
      for (int n = 0; n < N; n += 32)
        for (int j = 0; j < 16; ++j)
          exp[orow * ldo + n / 2 + j] =
              silu_h(acc[(size_t)m * N + n + j]) * acc[(size_t)m * N + n + 16 + j];
    } else {
      for (int ch = 0; ch < 3; ++ch) {
        float y = hb2[ch];
        for (int j = 0; j < 32; ++j) y += hw2[ch * 32 + j] * silu_h(acc[(size_t)m * N + j] + hbias[j]);
        y = y > 0 ? y : 1e-3f * expm1f(y);
        exp[orow * 3 + ch] = powf(10.f, y) - 1.f;
      }
    }
  }
  const double rt = (c.out_dtype == RFB_F32) ? 2e-3 : 1.2e-2;
  const double at = (c.out_dtype == RFB_F32) ? 2e-3 : 2e-2;
  report(c.name, got, exp, at, rt, (int)ldo);
  if (c.act) {
    std::string nm = std::string(c.name) + " [act]";
    report(nm.c_str(), got_act, exp_act, at, rt, (int)ldo);
  }
}

// fused-RMSNorm plumbing: in_sumsq partial sums / in_rscale (row or column), out_rscale,
// out_sumsq partial sums, out16 (+col_mul, aux map); the 16-bit side output is bf16, or fp16 with RFB_TEST_F16=1
static int g_o16 = RFB_BF16;
// kind: 0 = everything at once (generic epilogue), 1 = residual-stream shape (no col_mul -> EK_RESID when
// N % 256 == 0), 2 = 16-bit projection shape (no fp32 out, no residual -> EK_PROJ16)
static void run_fused(const char* name, int M, int N, int K, int scale_dim, bool swiglu, bool use_map, bool with_res,
                      bool rowmap = false, int kind = 0) {
  std::vector<uint16_t> hA = rand16((size_t)M * K, 5, 1.0f, RFB_BF16), hW = rand16((size_t)N * K, 6, 0.05f, RFB_BF16);
  DevBuf<uint16_t> dA((size_t)M * K), dW((size_t)N * K);
  dA.up(hA), dW.up(hW);
  DevBuf<float> dacc((size_t)M * N);
  dim3 g((N + 127) / 128, M);
  ref_linear<<<g, 128>>>(dA.p, K, dW.p, K, dacc.p, M, N, K, 1);
  CK(cudaDeviceSynchronize());
  std::vector<float> acc = dacc.down();
  const float eps = 1e-6f;
  const int norm_dim = 1024;
  const int in_parts = 8, in_ld = 8;
  // row mode: partial sums; column mode: ready-made factors
  std::vector<float> hparts((size_t)M * in_ld), hscale(scale_dim == 0 ? M : N);
  for (int m = 0; m < M; ++m) {
    float tot = 0.f;
    for (int j = 0; j < in_parts; ++j) {
      hparts[(size_t)m * in_ld + j] = (100.f + 900.f * (0.5f + 0.5f * hval(9, m * 8 + j))) / in_parts;
      tot += hparts[(size_t)m * in_ld + j];
    }
    if (scale_dim == 0) hscale[m] = 1.0f / sqrtf(tot / norm_dim + eps);
  }
  if (scale_dim == 1)
    for (int n = 0; n < N; ++n) hscale[n] = 0.5f + 0.5f * (0.5f + 0.5f * hval(19, n));
  std::vector<float> hcm = rand32(N, 10, 1.0f), hres = rand32((size_t)M * N, 12, 1.0f);
  std::vector<int> hmap(M), hrmap(M);
  for (int m = 0; m < M; ++m) hmap[m] = (m * 5 + 3) % M;   // permutations when gcd(., M) == 1
  for (int m = 0; m < M; ++m) hrmap[m] = rowmap ? (m * 3 + 1) % M : m;
  DevBuf<float> dparts((size_t)M * in_ld), dscale(hscale.size()), dcm(N), dres((size_t)M * N), drs(M);
  dparts.up(hparts), dscale.up(hscale), dcm.up(hcm), dres.up(hres);
  drs.fill_byte(0);
  DevBuf<int> dmap(M), drmap(M);
  dmap.up(hmap), drmap.up(hrmap);
  const bool want_sq = !swiglu && (N % 128 == 0);
  const int nparts = (N + 127) / 128;
  const int ocols = swiglu ? N / 2 : N;
  DevBuf<float> dout((size_t)M * ocols), dosq((size_t)M * nparts);
  DevBuf<uint16_t> do16((size_t)M * ocols), doutb((size_t)M * ocols);
  dout.fill_byte(0), do16.fill_byte(0), doutb.fill_byte(0);
  dosq.fill_byte(0x7f);  // every entry must be overwritten

  rfb_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.M = M, a.N = N, a.K = K, a.A = dA.p, a.lda = K, a.W = dW.p, a.ldw = K, a.dtype = RFB_BF16;
  a.norm_dim = norm_dim, a.norm_eps = eps;
  if (scale_dim == 0) {
    a.in_sumsq = dparts.p, a.in_sumsq_ld = in_ld, a.in_sumsq_parts = in_parts, a.out_rscale = swiglu ? nullptr : drs.p;
  } else {
    a.in_rscale = dscale.p, a.scale_dim = 1;
  }
  if (swiglu) {
    a.epi = RFB_EPI_SWIGLU, a.out = doutb.p, a.out_dtype = RFB_BF16, a.ldo = ocols;
  } else {
    a.epi = RFB_EPI_STORE, a.out = kind == 2 ? nullptr : dout.p, a.out_dtype = RFB_F32, a.ldo = N;
    if (with_res) a.res1 = dres.p, a.res_dtype = RFB_F32, a.ldres = N;
    if (want_sq) a.out_sumsq = dosq.p, a.out_sumsq_ld = nparts;
    a.out16 = do16.p, a.out16_dtype = g_o16, a.ld16 = N, a.col_mul = kind == 1 ? nullptr : dcm.p;
    if (kind == 1)
      for (auto& c : hcm) c = 1.0f;
    a.aux_row_map = use_map ? dmap.p : nullptr;
    a.row_map = rowmap ? drmap.p : nullptr;
  }
  int rc = rfb_gemm(&a, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc != RFB_OK || e != cudaSuccess) {
    printf("[FAIL] %-46s rc=%d cuda=%s\n", name, rc, cudaGetErrorString(e));
    g_fail++;
    if (e != cudaSuccess) exit(3);
    return;
  }
  auto sc = [&](int m, int n) { return scale_dim == 0 ? hscale[m] : hscale[n]; };
  if (swiglu) {
    std::vector<float> exp((size_t)M * ocols);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; n += 32)
        for (int j = 0; j < 16; ++j)
          exp[(size_t)m * ocols + n / 2 + j] =
              silu_h(acc[(size_t)m * N + n + j] * sc(m, 0)) * (acc[(size_t)m * N + n + 16 + j] * sc(m, 0));
    report(name, to_float(doutb.p, (size_t)M * ocols, RFB_BF16), exp, 2e-2, 1.2e-2, ocols);
    return;
  }
  std::vector<float> exp((size_t)M * N), exp16((size_t)M * N), expsq((size_t)M * nparts, 0.f);
  for (int m = 0; m < M; ++m) {
    const int orow = hrmap[m];
    const int ar = use_map ? hmap[orow] : orow;
    for (int n = 0; n < N; ++n) {
      float v = acc[(size_t)m * N + n] * sc(m, n);
      if (with_res) v += hres[(size_t)orow * N + n];
      exp[(size_t)orow * N + n] = v;
      exp16[(size_t)ar * N + n] = v * hcm[n];
      expsq[(size_t)ar * nparts + n / 128] += v * v;
    }
  }
  std::string nm(name);
  if (kind != 2) report((nm + " [out]").c_str(), dout.down(), exp, 2e-3, 2e-3, N);
  report((nm + " [out16]").c_str(), to_float(do16.p, (size_t)M * N, g_o16), exp16, 2e-2, 1.2e-2, N);
  if (want_sq) report((nm + " [sumsq]").c_str(), dosq.down(), expsq, 1e-2, 2e-3, nparts);
  if (scale_dim == 0) report((nm + " [rscale]").c_str(), drs.down(), hscale, 1e-5, 1e-4, 1);
}

// fused [q | k | v] projection: columns >= split leave transposed (V^T[b][dim][token]) times the row factor
static void run_vt(const char* name, int M, int N, int K, int split, int rows_per_batch, bool proj16) {
  std::vector<uint16_t> hA = rand16((size_t)M * K, 15, 1.0f, RFB_BF16), hW = rand16((size_t)N * K, 16, 0.05f, RFB_BF16);
  DevBuf<uint16_t> dA((size_t)M * K), dW((size_t)N * K);
  dA.up(hA), dW.up(hW);
  DevBuf<float> dacc((size_t)M * N);
  dim3 g((N + 127) / 128, M);
  ref_linear<<<g, 128>>>(dA.p, K, dW.p, K, dacc.p, M, N, K, 1);
  CK(cudaDeviceSynchronize());
  std::vector<float> acc = dacc.down();
  const float eps = 1e-6f;
  const int norm_dim = 512, parts = 4;
  std::vector<float> hparts((size_t)M * parts), hrs(M), hcm = rand32(split, 10, 1.0f);
  for (int m = 0; m < M; ++m) {
    float tot = 0.f;
    for (int j = 0; j < parts; ++j) tot += (hparts[(size_t)m * parts + j] = 50.f + 100.f * (0.5f + 0.5f * hval(29, m * 4 + j)));
    hrs[m] = 1.0f / sqrtf(tot / norm_dim + eps);
  }
  DevBuf<float> dparts(hparts.size()), dcm(split);
  dparts.up(hparts), dcm.up(hcm);
  const int nb = rows_per_batch > 0 ? M / rows_per_batch : 1;
  const int rpb = rows_per_batch > 0 ? rows_per_batch : M;
  const int nv = N - split;
  const long long vt_ld = (rpb + 7) & ~7;
  DevBuf<uint16_t> dvt((size_t)nb * nv * vt_ld), do16((size_t)M * split);
  DevBuf<float> dout((size_t)M * split), dosq((size_t)M * (split / 128));
  dvt.fill_byte(0), do16.fill_byte(0), dout.fill_byte(0), dosq.fill_byte(0x7f);
  rfb_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.M = M, a.N = N, a.K = K, a.A = dA.p, a.lda = K, a.W = dW.p, a.ldw = K, a.dtype = RFB_BF16, a.epi = RFB_EPI_STORE;
  a.in_sumsq = dparts.p, a.in_sumsq_ld = parts, a.in_sumsq_parts = parts, a.norm_dim = norm_dim, a.norm_eps = eps;
  a.out_dtype = RFB_F32, a.ldo = split;
  if (proj16) {
    a.out16 = do16.p, a.out16_dtype = g_o16, a.ld16 = split, a.col_mul = dcm.p;
    a.out_sumsq = dosq.p, a.out_sumsq_ld = split / 128;
  } else {
    a.out = dout.p;
  }
  a.vt_out = dvt.p, a.vt_dtype = RFB_BF16, a.vt_split = split, a.vt_rows_per_batch = rows_per_batch, a.vt_ld = vt_ld;
  a.vt_batch_stride = (long long)nv * vt_ld;
  int rc = rfb_gemm(&a, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc != RFB_OK || e != cudaSuccess) {
    printf("[FAIL] %-46s rc=%d cuda=%s\n", name, rc, cudaGetErrorString(e));
    g_fail++;
    if (e != cudaSuccess) exit(3);
    return;
  }
  std::vector<float> exp((size_t)M * split), exp16((size_t)M * split), expvt((size_t)nb * nv * vt_ld, 0.f);
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < split; ++n) {
      exp[(size_t)m * split + n] = acc[(size_t)m * N + n] * hrs[m];
      exp16[(size_t)m * split + n] = exp[(size_t)m * split + n] * hcm[n];
    }
    const int b = m / rpb, mr = m % rpb;
    for (int n = split; n < N; ++n) expvt[((size_t)b * nv + (n - split)) * vt_ld + mr] = acc[(size_t)m * N + n] * hrs[m];
  }
  std::string nm(name);
  if (proj16) report((nm + " [out16]").c_str(), to_float(do16.p, (size_t)M * split, g_o16), exp16, 2e-2, 1.2e-2, split);
  else report((nm + " [out]").c_str(), dout.down(), exp, 2e-3, 2e-3, split);
  report((nm + " [v^T]").c_str(), to_float(dvt.p, expvt.size(), RFB_BF16), expvt, 2e-2, 1.2e-2, (int)vt_ld);
}

// The residual-stream epilogue must not depend on the tile width: the row-sharded scene stage runs few-row
// launches on 128-wide tiles and has to reproduce the single-GPU schedule (256-wide tiles) bit for bit.
static void run_resid_widths(const char* name, int M, int N, int K, bool maps) {
  std::vector<uint16_t> hA = rand16((size_t)M * K, 25, 1.0f, RFB_BF16), hW = rand16((size_t)N * K, 26, 0.05f, RFB_BF16);
  DevBuf<uint16_t> dA((size_t)M * K), dW((size_t)N * K);
  dA.up(hA), dW.up(hW);
  std::vector<float> hres = rand32((size_t)M * N, 27, 1.0f), hparts((size_t)M * 8);
  for (size_t i = 0; i < hparts.size(); ++i) hparts[i] = 20.f + 100.f * (0.5f + 0.5f * hval(28, (int)i));
  std::vector<int> hmap(M), hrmap(M);
  for (int m = 0; m < M; ++m) hmap[m] = (m * 5 + 3) % M, hrmap[m] = (m * 3 + 1) % M;
  DevBuf<float> dparts(hparts.size());
  dparts.up(hparts);
  DevBuf<int> dmap(M), drmap(M);
  dmap.up(hmap), drmap.up(hrmap);
  std::vector<float> out[2], sq[2];
  std::vector<uint16_t> o16[2];
  for (int w = 0; w < 2; ++w) {
    DevBuf<float> dx((size_t)M * N), dsq((size_t)M * (N / 128));
    DevBuf<uint16_t> d16((size_t)M * N);
    dx.up(hres), dsq.fill_byte(0x7f), d16.fill_byte(0);
    rfb_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.M = M, a.N = N, a.K = K, a.A = dA.p, a.lda = K, a.W = dW.p, a.ldw = K, a.dtype = RFB_BF16, a.epi = RFB_EPI_STORE;
    a.in_sumsq = dparts.p, a.in_sumsq_ld = 8, a.in_sumsq_parts = 8, a.norm_dim = 1024, a.norm_eps = 1e-6f;
    a.out = dx.p, a.out_dtype = RFB_F32, a.ldo = N, a.res1 = dx.p, a.res_dtype = RFB_F32, a.ldres = N;
    a.out_sumsq = dsq.p, a.out_sumsq_ld = N / 128, a.out16 = d16.p, a.out16_dtype = g_o16, a.ld16 = N;
    if (maps) a.row_map = drmap.p, a.aux_row_map = dmap.p;
    a.bn_override = w == 0 ? 256 : 128;
    int rc = rfb_gemm(&a, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc != RFB_OK || e != cudaSuccess) {
      printf("[FAIL] %-46s bn=%d rc=%d cuda=%s\n", name, a.bn_override, rc, cudaGetErrorString(e));
      g_fail++;
      if (e != cudaSuccess) exit(3);
      return;
    }
    out[w] = dx.down(), sq[w] = dsq.down();
    o16[w].resize((size_t)M * N);
    CK(cudaMemcpy(o16[w].data(), d16.p, o16[w].size() * 2, cudaMemcpyDeviceToHost));
  }
  const bool same = !memcmp(out[0].data(), out[1].data(), out[0].size() * 4) &&
                    !memcmp(sq[0].data(), sq[1].data(), sq[0].size() * 4) &&
                    !memcmp(o16[0].data(), o16[1].data(), o16[0].size() * 2);
  size_t nsq = 0;
  for (size_t i = 0; i < sq[0].size(); ++i) nsq += memcmp(&sq[0][i], &sq[1][i], 4) != 0;
  printf("[%s] %-46s 256- vs 128-wide tiles: %s (%zu of %zu partial sums differ)\n", same ? " ok " : "FAIL", name,
         same ? "bit-identical" : "DIFFERENT", nsq, sq[0].size());
  if (!same) g_fail++;
}

// micro-benchmark of the fused residual / projection epilogues at the decoder's shapes
static void bench_fused(const char* name, int M, int N, int K, bool res, bool sumsq, bool o16, bool in_sq, bool maps,
                        bool f32out) {
  DevBuf<uint16_t> dA((size_t)M * K), dW((size_t)N * K), d16((size_t)M * N);
  CK(cudaMemset(dA.p, 0x3c, (size_t)M * K * 2));
  CK(cudaMemset(dW.p, 0x3c, (size_t)N * K * 2));
  const int pad = getenv("RFB_PAD") ? atoi(getenv("RFB_PAD")) : 0;   // experiment: padded row pitch
  const bool sep = getenv("RFB_SEP") != nullptr;                      // experiment: out != res
  const long long ldx = N + pad;
  DevBuf<float> dx((size_t)M * ldx), dx2((size_t)M * ldx), dsq((size_t)M * (N / 128)), dinsq((size_t)M * 8), dcm(N);
  dx.zero(), dx2.zero(), dinsq.zero(), dcm.zero();
  std::vector<int> hmap(M);
  for (int m = 0; m < M; ++m) hmap[m] = (int)(((long long)m * 8191 + 17) % M);
  DevBuf<int> dmap(M);
  dmap.up(hmap);
  rfb_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.M = M, a.N = N, a.K = K, a.A = dA.p, a.lda = K, a.W = dW.p, a.ldw = K, a.dtype = RFB_BF16;
  a.epi = RFB_EPI_STORE, a.out_dtype = RFB_F32, a.ldo = ldx;
  if (f32out) a.out = sep ? dx2.p : dx.p;
  if (res) a.res1 = dx.p, a.res_dtype = RFB_F32, a.ldres = ldx;
  if (sumsq) a.out_sumsq = dsq.p, a.out_sumsq_ld = N / 128;
  if (o16) a.out16 = d16.p, a.out16_dtype = RFB_BF16, a.ld16 = N, a.col_mul = res ? nullptr : dcm.p;
  if (in_sq) a.in_sumsq = dinsq.p, a.in_sumsq_ld = 8, a.in_sumsq_parts = 8, a.norm_dim = 1024, a.norm_eps = 1e-6f;
  if (maps) a.row_map = dmap.p;
  for (int i = 0; i < 3; ++i) {
    int rc = rfb_gemm(&a, 0);
    if (rc) {
      printf("bench %s rc=%d\n", name, rc);
      return;
    }
  }
  CK(cudaDeviceSynchronize());
  GpuTimer t;
  const int iters = 20;
  t.start();
  for (int i = 0; i < iters; ++i) rfb_gemm(&a, 0);
  float ms = t.stop() / iters;
  CK(cudaDeviceSynchronize());
  double tf = 2.0 * M * (double)N * K / (ms * 1e-3) / 1e12;
  printf("[BENCH] %-40s M=%d N=%d K=%d  %.3f ms  %.1f TFLOP/s\n", name, M, N, K, ms, tf);
  fflush(stdout);
}

static void bench(const char* name, int M, int N, int K, int epi, int out_dtype, int bn, int conv_hw = 0,
                  int Cin = 0, int batch = 1) {
  const int dtype = conv_hw ? RFB_F16 : RFB_BF16;
  size_t a_elems = conv_hw ? (size_t)batch * conv_hw * conv_hw * Cin : (size_t)M * K;
  DevBuf<uint16_t> dA(a_elems), dW((size_t)N * K);
  CK(cudaMemset(dA.p, 0x3c, a_elems * 2));  // small finite values
  CK(cudaMemset(dW.p, 0x3c, (size_t)N * K * 2));
  const int out_cols = (epi == RFB_EPI_SWIGLU) ? N / 2 : N;
  DevBuf<uint8_t> dout((size_t)M * out_cols * 4);
  DevBuf<float> dres((size_t)M * out_cols);
  dres.zero();
  rfb_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.M = M, a.N = N, a.K = K, a.A = dA.p, a.lda = conv_hw ? Cin : K, a.W = dW.p, a.ldw = K;
  a.dtype = dtype, a.epi = epi, a.out = dout.p, a.out_dtype = out_dtype, a.ldo = out_cols;
  a.bn_override = bn;
  if (conv_hw) a.a_mode = RFB_A_CONV3X3, a.B = batch, a.H = conv_hw, a.Wd = conv_hw, a.Cin = Cin;
  if (epi == RFB_EPI_STORE && out_dtype == RFB_F32 && !conv_hw)
    a.res1 = dres.p, a.res_dtype = RFB_F32, a.ldres = out_cols;  // residual-stream style
  for (int i = 0; i < 3; ++i) {
    int rc = rfb_gemm(&a, 0);
    if (rc) {
      printf("bench %s rc=%d\n", name, rc);
      return;
    }
  }
  CK(cudaDeviceSynchronize());
  GpuTimer t;
  const int iters = 20;
  t.start();
  for (int i = 0; i < iters; ++i) rfb_gemm(&a, 0);
  float ms = t.stop() / iters;
  CK(cudaDeviceSynchronize());
  double tf = 2.0 * M * (double)N * K / (ms * 1e-3) / 1e12;
  printf("[BENCH] %-40s M=%d N=%d K=%d bn=%d  %.3f ms  %.1f TFLOP/s\n", name, M, N, K, bn, ms, tf);
  fflush(stdout);
}

int main(int argc, char** argv) {
  if (const char* e = getenv("RFB_TEST_F16")) {
    if (e[0] == '1') g_o16 = RFB_F16, printf("selftest_gemm: fp16 side outputs (EK_RESID_H / EK_PROJ16_H)\n");
  }
  if (argc >= 8 && !strcmp(argv[1], "one")) {  // one M N K epi out_dtype bn
    // optional: conv_hw Cin batch
    bench("one", atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]),
          argc > 8 ? atoi(argv[8]) : 0, argc > 9 ? atoi(argv[9]) : 0, argc > 10 ? atoi(argv[10]) : 1);
    return 0;
  }
  if (argc >= 11 && !strcmp(argv[1], "fone")) {  // fone M N K res sumsq o16 insq maps f32out
    bench_fused("fone", atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]),
                atoi(argv[8]), atoi(argv[9]), atoi(argv[10]));
    return 0;
  }
  const bool do_bench = argc > 1 && !strcmp(argv[1], "bench");
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  printf("rfb version %d\n", rfb_version());
  if (!do_bench) {
    const int L = RFB_A_LINEAR, C3 = RFB_A_CONV3X3;
    std::vector<Case> cases = {
        // name, M,N,K, dtype, a_mode, B,H,W,Cin, epi, out_dtype, bias,res1,res2,act,rowmap, res_dtype, bn
        {"linear bf16 128x128x64 f32 (1 tile,1 kb)", 128, 128, 64, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F32, 0, 0, 0, 0, 0, RFB_F32, 128},
        {"linear bf16 128x128x256 f32 (k loop)", 128, 128, 256, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F32, 0, 0, 0, 0, 0, RFB_F32, 128},
        {"linear bf16 256x512x1024 f32 bn256", 256, 512, 1024, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F32, 0, 0, 0, 0, 0, RFB_F32, 256},
        {"linear bf16 300x96x192 ragged bn128", 300, 96, 192, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F32, 0, 0, 0, 0, 0, RFB_F32, 0},
        {"linear bf16 1040x1024 x1024 Nt-tail out", 1024, 1040, 1024, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_BF16, 0, 0, 0, 0, 0, RFB_F32, 0},
        {"linear bf16 4112x3072x1024 bf16 bn256", 4112, 3072, 1024, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_BF16, 0, 0, 0, 0, 0, RFB_F32, 256},
        {"linear bf16 bias+res f32 4112x1024x1024", 4112, 1024, 1024, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F32, 1, 1, 0, 0, 0, RFB_F32, 0},
        {"linear bf16 res+rowmap f32 2048x1024x512", 2048, 1024, 512, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F32, 0, 1, 0, 0, 1, RFB_F32, 0},
        {"linear bf16 swiglu 1000x2048x512", 1000, 2048, 512, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_SWIGLU, RFB_BF16, 0, 0, 0, 0, 0, RFB_F32, 0},
        {"linear f16 bias act f16 4096x256x1024", 4096, 256, 1024, RFB_F16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F16, 1, 0, 0, 1, 0, RFB_F16, 0},
        {"linear bf16 K=117->128 pad 512x256x128", 512, 256, 128, RFB_BF16, L, 0, 0, 0, 0, RFB_EPI_STORE, RFB_F32, 1, 0, 0, 0, 0, RFB_F32, 0},
        {"conv3x3 f16 2x32x32 128->128", 2 * 32 * 32, 128, 9 * 128, RFB_F16, C3, 2, 32, 32, 128, RFB_EPI_STORE, RFB_F16, 1, 0, 0, 0, 0, RFB_F16, 0},
        {"conv3x3 f16 1x16x16 256->128 res2 act", 256, 128, 9 * 256, RFB_F16, C3, 1, 16, 16, 256, RFB_EPI_STORE, RFB_F16, 1, 1, 1, 1, 0, RFB_F16, 0},
        {"conv3x3 f16 3x8x8 128->128 (W<16)", 3 * 64, 128, 9 * 128, RFB_F16, C3, 3, 8, 8, 128, RFB_EPI_STORE, RFB_F16, 0, 0, 0, 0, 0, RFB_F16, 0},
        {"conv3x3 f16 1x20x24 96->64 ragged", 480, 64, 9 * 96, RFB_F16, C3, 1, 20, 24, 96, RFB_EPI_STORE, RFB_F16, 1, 0, 0, 0, 0, RFB_F16, 0},
        {"conv3x3 f16 1x64x64 128->64 bn64", 4096, 64, 9 * 128, RFB_F16, C3, 1, 64, 64, 128, RFB_EPI_STORE, RFB_F16, 1, 0, 0, 0, 0, RFB_F16, 0},
        {"conv3x3 f16 final 1x64x64 64->32->3", 4096, 32, 9 * 64, RFB_F16, C3, 1, 64, 64, 64, RFB_EPI_FINAL, RFB_F32, 1, 0, 0, 0, 0, RFB_F16, 0},
        // large enough for the halo-tile kernel (>= 96 super-tiles of 16 x 16 pixels)
        {"conv3x3 halo 1x176x160 128->128 bias res2 act", 176 * 160, 128, 9 * 128, RFB_F16, C3, 1, 176, 160, 128, RFB_EPI_STORE, RFB_F16, 1, 1, 1, 1, 0, RFB_F16, 0},
        {"conv3x3 halo 1x170x150 128->128 ragged, act", 170 * 150, 128, 9 * 128, RFB_F16, C3, 1, 170, 150, 128, RFB_EPI_STORE, RFB_F16, 0, 0, 0, 1, 0, RFB_F16, 0},
        {"conv3x3 halo 2x100x140 64->64 bias", 2 * 100 * 140, 64, 9 * 64, RFB_F16, C3, 2, 100, 140, 64, RFB_EPI_STORE, RFB_F16, 1, 0, 0, 0, 0, RFB_F16, 0},
        {"conv3x3 halo 1x160x144 256->128 bias res1", 160 * 144, 128, 9 * 256, RFB_F16, C3, 1, 160, 144, 256, RFB_EPI_STORE, RFB_F16, 1, 1, 0, 0, 0, RFB_F16, 0},
        {"conv3x3 halo final 1x160x160 64->32->3", 160 * 160, 32, 9 * 64, RFB_F16, C3, 1, 160, 160, 64, RFB_EPI_FINAL, RFB_F32, 1, 0, 0, 0, 0, RFB_F16, 0},
    };
    int n = quick ? 3 : (int)cases.size();
    for (int i = 0; i < n; ++i) run_case(cases[i]);
    if (!quick) {
      run_fused("fused rowscale+sumsq+out16 1031x1024x512", 1031, 1024, 512, 0, false, false, true);
      run_fused("fused rowscale auxmap 517x512x256", 517, 512, 256, 0, false, true, false);
      run_fused("fused rowscale rowmap+res 1031x640x256", 1031, 640, 256, 0, false, false, true, true);
      run_fused("fused colscale 1024x1096x512", 1024, 1096, 512, 1, false, false, false);
      run_fused("EK_RESID 1031x1024x512 res", 1031, 1024, 512, 0, false, false, true, false, 1);
      run_fused("EK_RESID 1031x512x256 res rowmap auxmap", 1031, 512, 256, 0, false, true, true, true, 1);
      run_fused("EK_PROJ16 1031x2048x256", 1031, 2048, 256, 0, false, false, false, false, 2);
      run_fused("EK_RESID narrow tile 130x1024x512 res", 130, 1024, 512, 0, false, false, true, false, 1);
      // more tiles than SMs: the cross-tile residual prefetch of the persistent epilogue
      run_fused("EK_RESID 5000x1024x256 res (160 tiles)", 5000, 1024, 256, 0, false, false, true, false, 1);
      run_fused("EK_RESID 4999x1024x128 res rowmap auxmap", 4999, 1024, 128, 0, false, true, true, true, 1);
      run_resid_widths("EK_RESID 5000x1024x256 (160 / 320 tiles)", 5000, 1024, 256, false);
      run_resid_widths("EK_RESID 4999x1024x128 rowmap auxmap", 4999, 1024, 128, true);
      run_resid_widths("EK_RESID 520x1024x1024", 520, 1024, 1024, false);
      run_resid_widths("EK_RESID 1031x512x256 rowmap auxmap", 1031, 512, 256, true);
      run_vt("fused qkv (generic) 517x768x256 split 512", 517, 768, 256, 512, 0, false);
      run_vt("fused qkv (generic) 2x264 rows, batched v^T", 528, 768, 256, 512, 264, false);
      run_vt("fused qkv (PROJ16) 1031x1536x256 split 1024", 1031, 1536, 256, 1024, 0, true);
      run_fused("fused swiglu rowscale 777x2048x512", 777, 2048, 512, 0, true, false, false);
    }
    printf("selftest_gemm: %d failure(s)\n", g_fail);
    if (g_fail) return 1;
  }
  if (do_bench || argc == 1) {
    for (int bn : {256, 128}) {
      bench("enc qkv        ", 4112, 3072, 1024, RFB_EPI_STORE, RFB_BF16, bn);
      bench("enc out_proj+res", 4112, 1024, 1024, RFB_EPI_STORE, RFB_F32, bn);
      bench("enc w13 swiglu ", 4112, 8192, 1024, RFB_EPI_SWIGLU, RFB_BF16, bn);
      bench("enc w2+res     ", 4112, 1024, 4096, RFB_EPI_STORE, RFB_F32, bn);
      bench("dec 8v w13     ", 32768, 8192, 1024, RFB_EPI_SWIGLU, RFB_BF16, bn);
      bench("dec 8v w2+res  ", 32768, 1024, 4096, RFB_EPI_STORE, RFB_F32, bn);
      bench("dec 8v q_proj  ", 32768, 1024, 1024, RFB_EPI_STORE, RFB_F32, bn);
      bench("square 8192    ", 8192, 8192, 8192, RFB_EPI_STORE, RFB_BF16, bn);
    }
    bench_fused("shard/8 wo: res+sumsq+out16 ", 520, 1024, 1024, true, true, true, false, false, true);
    bench_fused("shard/8 w2: res+sumsq+out16 ", 520, 1024, 4096, true, true, true, false, false, true);
    for (int bn : {256, 128, 64, 0}) {
      bench("shard/8 tokens K=13312", 504, 1024, 13312, RFB_EPI_STORE, RFB_BF16, bn);
      bench("shard/8 w13 swiglu    ", 520, 8192, 1024, RFB_EPI_SWIGLU, RFB_BF16, bn);
      bench("shard/8 qk f32        ", 520, 2048, 1024, RFB_EPI_STORE, RFB_BF16, bn);
    }
    bench_fused("dec wout: res+sumsq+out16 ", 16384, 1024, 1024, true, true, true, false, false, true);
    bench_fused("dec s.wo: res+sumsq+out16+map", 16384, 1024, 1024, true, true, true, false, true, true);
    bench_fused("dec w2: res+sumsq+out16 K4096", 16384, 1024, 4096, true, true, true, false, false, true);
    bench_fused("dec wq: insq+out16+sumsq    ", 16384, 1024, 1024, false, true, true, true, false, false);
    bench_fused("dec s.wqk: insq+out16+sumsq ", 16384, 2048, 1024, false, true, true, true, false, false);
    bench_fused("plain res only              ", 16384, 1024, 1024, true, false, false, false, false, true);
    bench_fused("plain f32 store only        ", 16384, 1024, 1024, false, false, false, false, false, true);
    bench("conv 4x256^2 128->128", 262144, 128, 1152, RFB_EPI_STORE, RFB_F16, 128, 256, 128, 4);
    bench("conv 4x512^2 128->64 ", 1048576, 64, 1152, RFB_EPI_STORE, RFB_F16, 64, 512, 128, 4);
    bench("conv 256^2 128->128", 65536, 128, 1152, RFB_EPI_STORE, RFB_F16, 128, 256, 128);
    bench("conv 512^2 128->64 ", 262144, 64, 1152, RFB_EPI_STORE, RFB_F16, 64, 512, 128);
  }
  return g_fail ? 1 : 0;
}
