// Dense flash attention, fourth generation: attention3's data flow (one 128-query tile per CTA, Q and P as
// TMEM A operands, S double-buffered, 3-stage K / V^T rings) with the softmax split over TWO threads per
// query row.
//
//   warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-3 idle (warpgroup alignment),
//   warps 4..11 = softmax: warp w works on TMEM lane quarter (w & 3) and on key columns
//   [64 * h, 64 * h + 64) of every S tile, h = (w - 4) >> 2.
//
// Why: ncu on attention3 (profiles/r01e_attn_cross_summary.txt) shows the softmax warps at 0.36
// instructions per cycle and scheduler with `wait` / `long_scoreboard` as the top stalls and every pipe
// below 40 % -- one warp per scheduler cannot hide its own dependency latencies, and the 2000 cycles it needs
// per 128 x 128 tile are twice the tile's tensor time.  Two warps per scheduler, each with half the
// elements of a row, halve the per-tile latency and overlap each other's stalls.
//
// The two threads of a row share ONE running maximum: each publishes the maximum of its 64 columns in
// shared memory, a 64-thread named barrier (one per lane quarter: both warps of a quarter sit on the same
// scheduler) makes the pair's values visible, and both continue with max(own, partner).  Row sums stay
// private (same maximum schedule => they add up at the end); the rare O rescale and the epilogue are split
// by columns.
//
// Used for mode 0 when the grid of one-tile CTAs fills the SMs best (decoder cross-attention).
// Reference call sites: include/rfb200.h.
#include <atomic>

#include "host_util.h"
#include "ptx.cuh"

namespace rfb {

extern std::atomic<long long> g_launch_count;

struct Attn4Params {
  int Nq, Nk, n_kv_tiles;
  int k_batched, v_batched;
  const uint16_t* Q;
  long long ldq, q_batch_stride;
  const uint32_t* mask_bits;
  long long mask_stride_words;
  void* O;
  long long ldo, o_batch_stride;
  float scale_log2;
  const float* q_sumsq;
  int sumsq_ld, sumsq_parts;
  float inv_norm_dim, norm_eps;
};

constexpr uint32_t kT4 = 128 * 128 * 2;  // 128 x 128 16-bit tile
constexpr uint32_t kH4 = 128 * 64 * 2;   // one 64-column swizzle half
constexpr int kStg4 = 3;
constexpr int kAttn4Threads = 128 + 256;
constexpr uint32_t kAttn4Smem = kT4 * 2 * kStg4 + 1024 + 256 + 2 * 2 * 128 * 4;  // + row-max exchange (2 buffers)
constexpr int kPoly4 = 4;  // every 4th pair of exponentials on the FMA pipe

__device__ __forceinline__ void exp2_poly2_4(uint64_t t2, float& p0, float& p1) {
  const float ta = fmaxf(lo2f(t2), -126.0f), tb = fmaxf(hi2f(t2), -126.0f);
  const uint64_t t = pack2f(ta, tb);
  const uint64_t r = fadd2(t, pack2f(12582912.0f, 12582912.0f));
  const uint64_t n = fadd2(r, pack2f(-12582912.0f, -12582912.0f));
  const uint64_t f = ffma2(n, pack2f(-1.0f, -1.0f), t);
  uint64_t p = ffma2(f, pack2f(0.0551716685f, 0.0551716685f), pack2f(0.2426111251f, 0.2426111251f));
  p = ffma2(p, f, pack2f(0.6932609677f, 0.6932609677f));
  p = ffma2(p, f, pack2f(0.9999280572f, 0.9999280572f));
  p0 = __uint_as_float(__float_as_uint(lo2f(p)) + (__float_as_uint(lo2f(r)) << 23));
  p1 = __uint_as_float(__float_as_uint(hi2f(p)) + (__float_as_uint(hi2f(r)) << 23));
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool F16>
__global__ void __launch_bounds__(kAttn4Threads, 1)
    attn4_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const Attn4Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;                  // kStg4 tiles
  uint8_t* sV = sK + kStg4 * kT4;      // kStg4 tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStg4 * kT4);
  uint64_t* k_full = bars;              // kStg4
  uint64_t* k_empty = k_full + kStg4;   // kStg4
  uint64_t* v_full = k_empty + kStg4;   // kStg4
  uint64_t* v_empty = v_full + kStg4;   // kStg4
  uint64_t* s_full = v_empty + kStg4;   // 2
  uint64_t* p_full = s_full + 2;        // 2 (8 softmax warps each)
  uint64_t* pv_done = p_full + 2;       // 1
  uint64_t* q_ready = pv_done + 1;      // 1 (8 softmax warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 buffers][2 halves][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStg4; ++i) {
      mbar_init(&k_full[i], 1), mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1), mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&s_full[i], 1), mbar_init(&p_full[i], 8);
    mbar_init(pv_done, 1);
    mbar_init(q_ready, 8);
    fence_mbar_init();
    tma_prefetch_desc(&tmK), tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + 256, tQ = tmem_base + 384;
  const int n_tiles = p.n_kv_tiles;
  const int kb = p.k_batched ? b : 0;
  const int vb = p.v_batched ? b : 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kStg4;
        const uint32_t ph = (j / kStg4) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kT4);
        tma_load_3d(sK + s * kT4, &tmK, &k_full[s], h * 128, j * 128, kb);
        tma_load_3d(sK + s * kT4 + kH4, &tmK, &k_full[s], h * 128 + 64, j * 128, kb);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kT4);
        tma_load_3d(sV + s * kT4, &tmV, &v_full[s], j * 128, h * 128, vb);
        tma_load_3d(sV + s * kT4 + kH4, &tmV, &v_full[s], j * 128 + 64, h * 128, vb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(F16 ? 0u : 1u, 128, 128);
      auto issue_qk = [&](int j) {  // S[j&1] = Q K_j^T, A operand = Q in TMEM (packed 16-bit pairs)
        const int s = j % kStg4;
        mbar_wait(&k_full[s], (j / kStg4) & 1);
        tc_fence_after();
        const uint64_t bd = umma_desc_sw128(smem_u32(sK + s * kT4));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kH4 >> 4) + (k & 3) * 2;
          umma_f16_ts(tmem_base + (j & 1) * 128, tQ + k * 8, bd + off, idesc, k != 0);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[j & 1]);
      };
      auto issue_pv = [&](int j) {  // O += P_j V_j, A operand = P in S[j&1] columns [0,64)
        const int s = j % kStg4;
        mbar_wait(&v_full[s], (j / kStg4) & 1);
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t bd = umma_desc_sw128(smem_u32(sV + s * kT4));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kH4 >> 4) + (k & 3) * 2;
          umma_f16_ts(tO, tmem_base + (j & 1) * 128 + k * 8, bd + off, idesc, (j | k) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      };
      mbar_wait(q_ready, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        issue_pv(j);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ softmax warps ------------------------------
    const int q = warp & 3;              // TMEM lane quarter
    const int hc = (warp - 4) >> 2;      // which 64 key columns of a tile (and which 64 O columns)
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int qrow = q0 + r;
    const bool row_ok = qrow < p.Nq;

    // Q row -> TMEM (64 packed columns; this thread writes its half of them); rows past Nq are zero
    {
      const uint16_t* qsrc = p.Q + static_cast<long long>(b) * p.q_batch_stride + static_cast<long long>(qrow) * p.ldq +
                             h * 128 + hc * 64;
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok) u = __ldg(reinterpret_cast<const uint4*>(qsrc + i * 8));
        w[i * 4] = u.x, w[i * 4 + 1] = u.y, w[i * 4 + 2] = u.z, w[i * 4 + 3] = u.w;
      }
      tmem_st32(tQ + lane_addr + hc * 32, w);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }

    float sl2 = p.scale_log2;
    if (p.q_sumsq && row_ok) {  // fused q RMSNorm: per-row 1/rms folded into the scale
      const float* sp = p.q_sumsq + (static_cast<long long>(b) * p.Nq + qrow) * p.sumsq_ld;
      float ss = 0.f;
      for (int j = 0; j < p.sumsq_parts; ++j) ss += sp[j];
      sl2 *= rsqrtf(ss * p.inv_norm_dim + p.norm_eps);
    }
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t tOl = tO + lane_addr + hc * 64;

    for (int j = 0; j < n_tiles; ++j) {
      uint32_t mw[2];
      if (p.mask_bits) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p.mask_bits + static_cast<long long>(b) * p.mask_stride_words +
                                                             j * 4 + hc * 2));
        mw[0] = u.x, mw[1] = u.y;
      } else {
        const int rem = p.Nk - j * 128 - hc * 64;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int lo = w * 32;
          mw[w] = rem <= lo ? 0u : (rem - lo >= 32 ? 0xffffffffu : ((1u << (rem - lo)) - 1u));
        }
      }
      const bool all_valid = (mw[0] & mw[1]) == 0xffffffffu;
      const uint32_t tS = tmem_base + (j & 1) * 128 + lane_addr;

      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[2][32];
      tmem_ld32(tS + hc * 64, v[0]);
      tmem_ld32(tS + hc * 64 + 32, v[1]);
      tmem_wait_ld();

      if (!all_valid) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t bits = mw[c];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!((bits >> i) & 1u)) v[c][i] = 0xff800000u;  // -inf
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[c][i]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // one running maximum per row: exchange the half-row maxima with the partner warp (same lane quarter, same
      // scheduler) through shared memory; buffers alternate with j so the next tile's writes cannot overtake
      float* xb = xchg + (j & 1) * 256;
      xb[hc * 128 + r] = mx;
      named_bar_sync(1 + q, 64);
      mx = fmaxf(mx, xb[(hc ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_f((m_run - m_use) * sl2);
      const float neg_ms = -m_use * sl2;

      // alpha is identical in both warps of the quarter (same row maxima), so the vote agrees between them
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        mbar_wait(pv_done, (j - 1) & 1);  // P_{j-1} V_{j-1} retired: O is stable
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {  // this thread's 64 of the 128 O columns
          uint32_t o[32];
          tmem_ld32(tOl + c * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tOl + c * 32, o);
        }
      }

      // P = exp2(s*sl2 - m*sl2) -> 16-bit pairs -> columns [32*hc, 32*hc + 32) of S[j&1]
      const uint64_t sl2_2 = pack2f(sl2, sl2), nms_2 = pack2f(neg_ms, neg_ms);
      uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t t2 = ffma2(pack2f(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1])), sl2_2, nms_2);
          float p0, p1;
          if (kPoly4 > 0 && (i % (kPoly4 > 0 ? kPoly4 : 1)) == kPoly4 - 1) {
            exp2_poly2_4(t2, p0, p1);
          } else {
            p0 = ex2_f(lo2f(t2)), p1 = ex2_f(hi2f(t2));
          }
          rs2[i & 3] = fadd2(rs2[i & 3], pack2f(p0, p1));
          pk[i] = pack16<F16>(p0, p1);
        }
        tmem_st16(tS + hc * 32 + c * 16, pk);
      }
      const uint64_t rsum = fadd2(fadd2(rs2[0], rs2[1]), fadd2(rs2[2], rs2[3]));
      l_run = l_run * alpha + (lo2f(rsum) + hi2f(rsum));
      m_run = m_new;

      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }

    // epilogue: the row sum is the sum of the two halves' sums; O / l -> 16 bit -> global (64 columns each)
    float* xb = xchg + (n_tiles & 1) * 256;
    xb[hc * 128 + r] = l_run;
    named_bar_sync(1 + q, 64);
    const float l_tot = l_run + xb[(hc ^ 1) * 128 + r];
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = (l_tot > 0.f) ? 1.0f / l_tot : 0.f;
    uint16_t* orow = static_cast<uint16_t*>(p.O) + static_cast<long long>(b) * p.o_batch_stride +
                     static_cast<long long>(qrow) * p.ldo + h * 128 + hc * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tOl + c * 32, o);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack16<F16>(__uint_as_float(o[i * 8 + 0]) * inv_l, __uint_as_float(o[i * 8 + 1]) * inv_l);
          u.y = pack16<F16>(__uint_as_float(o[i * 8 + 2]) * inv_l, __uint_as_float(o[i * 8 + 3]) * inv_l);
          u.z = pack16<F16>(__uint_as_float(o[i * 8 + 4]) * inv_l, __uint_as_float(o[i * 8 + 5]) * inv_l);
          u.w = pack16<F16>(__uint_as_float(o[i * 8 + 6]) * inv_l, __uint_as_float(o[i * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// called by rfb_attention (attention.cu) for mode 0
int launch_attention4(const CUtensorMap& tmK, const CUtensorMap& tmV, const rfb_attn_args* a, int k_batched,
                      int v_batched, cudaStream_t stream) {
  if (a->ldq % 8 || (reinterpret_cast<uintptr_t>(a->Q) & 15) || (a->B > 1 && a->q_batch_stride % 8)) return RFB_ERR_ALIGN;
  Attn4Params p{};
  p.Nq = a->Nq, p.Nk = a->Nk;
  p.n_kv_tiles = (a->Nk + 127) / 128;
  p.k_batched = k_batched, p.v_batched = v_batched;
  p.Q = static_cast<const uint16_t*>(a->Q), p.ldq = a->ldq;
  p.q_batch_stride = a->B > 1 ? a->q_batch_stride : 0;
  p.mask_bits = a->key_mask_bits;
  p.mask_stride_words = a->mask_batch_stride_words;
  p.O = a->O, p.ldo = a->ldo, p.o_batch_stride = a->o_batch_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.q_sumsq = a->q_sumsq, p.sumsq_ld = a->sumsq_ld > 0 ? a->sumsq_ld : 1;
  p.sumsq_parts = a->sumsq_parts > 0 ? a->sumsq_parts : 1;
  p.inv_norm_dim = a->norm_dim > 0 ? 1.0f / (float)a->norm_dim : 0.f, p.norm_eps = a->norm_eps;
  const bool f16 = a->dtype == RFB_F16;
  auto kern = f16 ? attn4_tc_kernel<true> : attn4_tc_kernel<false>;
  static PerDeviceFlag attr_flags[2];
  bool& attr_set = attr_flags[f16].get();
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn4Smem) != cudaSuccess)
      return RFB_ERR_LAUNCH;
    attr_set = true;
  }
  dim3 grid((a->Nq + 127) / 128, a->H, a->B);
  kern<<<grid, kAttn4Threads, kAttn4Smem, stream>>>(tmK, tmV, p);
  g_launch_count++;
  return check_launch("attn4_tc_kernel");
}

}  // namespace rfb
