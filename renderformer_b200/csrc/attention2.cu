// Dense flash attention, second generation: two 128-query tiles per CTA sharing one K/V stream,
// P kept in tensor memory.
//
//   warpgroup 0: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-3 idle (they exist so
//   that setmaxnreg can hand the group's registers to the softmax warpgroups);
//   warpgroup 1 (warps 4-7) = softmax of query tile A, warpgroup 2 (warps 8-11) = softmax of query
//   tile B, one thread per query row.
//
// The softmax is the co-critical resource next to the tensor pipe (128 x 128 exponentials per
// 2 x 4.2 MFLOP tile pair: the 16/clk/SM MUFU rate equals the tensor time), so its inner loop uses
// packed fp32x2 arithmetic (FFMA2 / FADD2) and evaluates a quarter of the exponentials with a
// degree-3 polynomial on the FMA pipe (Cody-Waite split, exponent inserted with an integer add)
// instead of MUFU.EX2.
//
// TMEM map (512 columns): S_A [0,128)  S_B [128,256)  O_A [256,384)  O_B [384,512).
// After the softmax has pulled its whole S row into registers it writes P (bf16 pairs) back into
// the first 64 columns of the same S region; the P.V product then takes its A operand straight
// from TMEM (tcgen05.mma with a TMEM A descriptor), so P never touches shared memory and the
// only smem traffic of the second GEMM is V.  The MMA issue order
//     PV_A(j), QK_A(j+1), PV_B(j), QK_B(j+1)
// lets the tensor pipe work on one query tile while the other tile's softmax runs.  Because
// tcgen05 operations of one thread complete in order, "S_t(j) is ready" already implies
// "PV_t(j-1) is done", so the (rare) O rescale needs no extra barrier.
//
// Used for mode 0 (encoder self-attention, decoder cross-attention, key-padding mask); the
// block-diagonal swin mode stays on attn_tc_kernel.  Reference call sites: include/rfb200.h.
#include <atomic>

#include "host_util.h"
#include "ptx.cuh"

namespace rfb {

extern std::atomic<long long> g_launch_count;

struct Attn2Params {
  int Nq, Nk, n_kv_tiles;
  int k_batched, v_batched;
  const uint32_t* mask_bits;
  long long mask_stride_words;
  void* O;
  long long ldo, o_batch_stride;
  float scale_log2;
  const float* q_sumsq;
  int sumsq_ld, sumsq_parts;
  float inv_norm_dim, norm_eps;
  // key splitting: blockIdx.z = b + B * split; split s visits key tiles [s*split_tiles, min(.., n_kv_tiles)) and
  // leaves its un-normalised result in part_o / part_ml (attention.cu: attn_combine_kernel finishes the job)
  int B, split_tiles;
  float* part_o;   // [split][B][H][Nq][128] fp32
  float* part_ml;  // [split][B][H][Nq][2]   (running max in log2 units, running sum)
};

constexpr uint32_t kT2 = 128 * 128 * 2;  // 128 x 128 bf16 tile
constexpr uint32_t kH2 = 128 * 64 * 2;   // one 64-column swizzle half
constexpr int kStg = 2;
constexpr int kAttn2Threads = 3 * 128;
constexpr int kRegsWg0 = 56, kRegsSoftmax = 224;  // setmaxnreg targets (128*56 + 256*224 = 384*168)
// every kPolyEvery-th pair of exponentials goes to the FMA-pipe polynomial (0 = all on MUFU)
constexpr int kPolyEvery = 4;

// 2^t for t in [-126, 0] on the FMA pipe, two lanes at a time: t = n + f, n = round(t), |f| <= 0.5,
// 2^f ~ degree-3 minimax polynomial (rel. error 7.5e-5, far below bf16 resolution of P), and 2^n is
// applied by adding n to the exponent field (r = t + 1.5*2^23 keeps n in its low mantissa bits).
__device__ __forceinline__ void exp2_poly2(uint64_t t2, float& p0, float& p1) {
  const float ta = fmaxf(lo2f(t2), -126.0f), tb = fmaxf(hi2f(t2), -126.0f);
  const uint64_t t = pack2f(ta, tb);
  const uint64_t magic = pack2f(12582912.0f, 12582912.0f);
  const uint64_t r = fadd2(t, magic);
  const uint64_t n = fadd2(r, pack2f(-12582912.0f, -12582912.0f));
  const uint64_t f = ffma2(n, pack2f(-1.0f, -1.0f), t);
  uint64_t p = ffma2(f, pack2f(0.0551716685f, 0.0551716685f), pack2f(0.2426111251f, 0.2426111251f));
  p = ffma2(p, f, pack2f(0.6932609677f, 0.6932609677f));
  p = ffma2(p, f, pack2f(0.9999280572f, 0.9999280572f));
  p0 = __uint_as_float(__float_as_uint(lo2f(p)) + (__float_as_uint(lo2f(r)) << 23));
  p1 = __uint_as_float(__float_as_uint(hi2f(p)) + (__float_as_uint(hi2f(r)) << 23));
}
constexpr uint32_t kAttn2Smem = kT2 * (2 + 2 * kStg) + 1024 + 256;

template <bool F16>
__global__ void __launch_bounds__(kAttn2Threads, 1)
    attn2_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                 // 2 tiles
  uint8_t* sK = sQ + 2 * kT2;         // kStg tiles
  uint8_t* sV = sK + kStg * kT2;      // kStg tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStg * kT2);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = q_full + 1;      // kStg
  uint64_t* k_empty = k_full + kStg;  // kStg
  uint64_t* v_full = k_empty + kStg;  // kStg
  uint64_t* v_empty = v_full + kStg;  // kStg
  uint64_t* s_full = v_empty + kStg;  // 2 (per query tile)
  uint64_t* p_full = s_full + 2;      // 2
  uint64_t* o_done = p_full + 2;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int split = p.split_tiles > 0 ? blockIdx.z / p.B : 0;
  const int b = blockIdx.z - split * p.B;
  const int q0 = blockIdx.x * 256;
  const int j_begin = split * p.split_tiles;  // first key tile of this CTA

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kStg; ++i) {
      mbar_init(&k_full[i], 1), mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1), mbar_init(&v_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) mbar_init(&s_full[t], 1), mbar_init(&p_full[t], 4), mbar_init(&o_done[t], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ), tma_prefetch_desc(&tmK), tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = p.split_tiles > 0 ? min(p.split_tiles, p.n_kv_tiles - j_begin) : p.n_kv_tiles;
  // Short tail: when the LAST key tile of the sequence holds <= 32 keys (16 register tokens + a multiple of 128
  // triangles is the common case) its products run 32 keys wide -- S = Q K^T with N = 32, P V with two K = 16 steps --
  // and the softmax touches one 32-column chunk instead of four.  The skipped columns are masked keys, whose
  // probabilities are exactly 0: the result is bit-identical to the full-width tile.
  const int tail_local = (p.Nk - (p.n_kv_tiles - 1) * 128 <= 32) ? p.n_kv_tiles - 1 - j_begin : -1;  // local index or never
  const int kb = p.k_batched ? b : 0;
  const int vb = p.v_batched ? b : 0;

  if (warp == 0) {
    setmaxnreg_dec<kRegsWg0>();
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * kT2);
      for (int t = 0; t < 2; ++t) {
        tma_load_3d(sQ + t * kT2, &tmQ, q_full, h * 128, q0 + t * 128, b);
        tma_load_3d(sQ + t * kT2 + kH2, &tmQ, q_full, h * 128 + 64, q0 + t * 128, b);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kStg;
        const uint32_t ph = (j / kStg) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kT2);
        const int key0 = (j_begin + j) * 128;
        tma_load_3d(sK + s * kT2, &tmK, &k_full[s], h * 128, key0, kb);
        tma_load_3d(sK + s * kT2 + kH2, &tmK, &k_full[s], h * 128 + 64, key0, kb);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kT2);
        tma_load_3d(sV + s * kT2, &tmV, &v_full[s], key0, h * 128, vb);
        tma_load_3d(sV + s * kT2 + kH2, &tmV, &v_full[s], key0 + 64, h * 128, vb);
      }
    }
  } else if (warp == 1) {
    setmaxnreg_dec<kRegsWg0>();
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(F16 ? 0u : 1u, 128, 128);
      const uint32_t idesc_tail = umma_idesc_f16(F16 ? 0u : 1u, 128, 32);
      auto issue_qk = [&](int t, int j) {  // S_t = Q_t K_j^T
        const int s = j % kStg;
        const uint64_t ad = umma_desc_sw128(smem_u32(sQ + t * kT2));
        const uint64_t bd = umma_desc_sw128(smem_u32(sK + s * kT2));
        const uint32_t id = (j == tail_local) ? idesc_tail : idesc;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kH2 >> 4) + (k & 3) * 2;
          umma_f16(tmem_base + t * 128, ad + off, bd + off, id, k != 0);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {  // O_t += P_t V_j   (A operand from TMEM)
        const int s = j % kStg;
        mbar_wait(&p_full[t], j & 1);
        tc_fence_after();
        const uint64_t bd = umma_desc_sw128(smem_u32(sV + s * kT2));
        const int ksteps = (j == tail_local) ? 2 : 8;  // short tail: 32 keys = two K = 16 steps
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kH2 >> 4) + (k & 3) * 2;
          if (k < ksteps)
            umma_f16_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + k * 8, bd + off, idesc, (j | k) != 0);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kStg;
        const bool more = j + 1 < n_tiles;
        const int s1 = (j + 1) % kStg;
        mbar_wait(&v_full[s], (j / kStg) & 1);
        issue_pv(0, j);
        if (more) {
          mbar_wait(&k_full[s1], ((j + 1) / kStg) & 1);
          tc_fence_after();
          issue_qk(0, j + 1);
        }
        issue_pv(1, j);
        umma_commit(&v_empty[s]);
        if (more) {
          issue_qk(1, j + 1);
          umma_commit(&k_empty[s1]);
        }
      }
      umma_commit(&o_done[0]);
      umma_commit(&o_done[1]);
    }
  } else if (warp < 4) {
    setmaxnreg_dec<kRegsWg0>();  // idle warps of warpgroup 0
  } else {
    // ------------------------------ softmax warps ------------------------------
    setmaxnreg_inc<kRegsSoftmax>();
    const int t = (warp - 4) >> 2;  // query tile of this warp group
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_addr;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_addr;
    float sl2 = p.scale_log2;
    if (p.q_sumsq && q0 + t * 128 + r < p.Nq) {  // fused q RMSNorm: per-row 1/rms folded into the scale
      const float* sp = p.q_sumsq + (static_cast<long long>(b) * p.Nq + q0 + t * 128 + r) * p.sumsq_ld;
      float ss = 0.f;
      for (int j = 0; j < p.sumsq_parts; ++j) ss += sp[j];
      sl2 *= rsqrtf(ss * p.inv_norm_dim + p.norm_eps);
    }
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      uint32_t mw[4];
      if (p.mask_bits) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.mask_bits + static_cast<long long>(b) * p.mask_stride_words + (j_begin + j) * 4));
        mw[0] = u.x, mw[1] = u.y, mw[2] = u.z, mw[3] = u.w;
      } else {
        const int rem = p.Nk - (j_begin + j) * 128;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int lo = w * 32;
          mw[w] = rem <= lo ? 0u : (rem - lo >= 32 ? 0xffffffffu : ((1u << (rem - lo)) - 1u));
        }
      }
      const bool all_valid = (mw[0] & mw[1] & mw[2] & mw[3]) == 0xffffffffu;

      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const int nc = (j == tail_local) ? 1 : 4;  // 32-column chunks of S that exist (short tail: one)
      uint32_t v[4][32];
      tmem_ld32(tS, v[0]);
      if (nc == 4) {
        tmem_ld32(tS + 32, v[1]);
        tmem_ld32(tS + 64, v[2]);
        tmem_ld32(tS + 96, v[3]);
      }
      tmem_wait_ld();

      if (!all_valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nc) {
            const uint32_t bits = mw[c];
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!((bits >> i) & 1u)) v[c][i] = 0xff800000u;  // -inf
          }
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nc) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[c][i]));
        }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_f((m_run - m_use) * sl2);
      const float neg_ms = -m_use * sl2;

      // S_t(j) ready implies PV_t(j-1) retired (in-order tcgen05 pipe): O_t may be rescaled now
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + c * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tO + c * 32, o);
        }
      }

      // P = exp2(s*sl2 - m*sl2) -> bf16 pairs -> columns [0,64) of this tile's S region
      const uint64_t sl2_2 = pack2f(sl2, sl2), nms_2 = pack2f(neg_ms, neg_ms);
      uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nc) continue;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t t2 = ffma2(pack2f(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1])), sl2_2, nms_2);
          float p0, p1;
          if (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) {
            exp2_poly2(t2, p0, p1);
          } else {
            p0 = ex2_f(lo2f(t2)), p1 = ex2_f(hi2f(t2));
          }
          rs2[i & 3] = fadd2(rs2[i & 3], pack2f(p0, p1));
          pk[i] = pack16<F16>(p0, p1);
        }
        tmem_st16(tS + c * 16, pk);
      }
      const uint64_t rsum = fadd2(fadd2(rs2[0], rs2[1]), fadd2(rs2[2], rs2[3]));
      float rs4[4] = {lo2f(rsum), hi2f(rsum), 0.f, 0.f};
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      m_run = m_new;

      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }

    // epilogue: O / l -> bf16 -> global
    mbar_wait(&o_done[t], 0);
    tc_fence_after();
    const int qrow = q0 + t * 128 + r;
    const bool row_ok = qrow < p.Nq;
    if (p.part_o) {  // key split: un-normalised O (relative to m_run), (m_run in log2 units, l_run)
      const long long prow = ((static_cast<long long>(split) * p.B + b) * gridDim.y + h) * p.Nq + qrow;
      if (row_ok) {
        float2* ml = reinterpret_cast<float2*>(p.part_ml) + prow;
        *ml = make_float2(m_run * sl2, l_run);
      }
      float* po = p.part_o + prow * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(po + c * 32 + i * 4) = make_uint4(o[i * 4], o[i * 4 + 1], o[i * 4 + 2], o[i * 4 + 3]);
        }
      }
    } else {
    const float inv_l = (l_run > 0.f) ? 1.0f / l_run : 0.f;
    uint16_t* orow = static_cast<uint16_t*>(p.O) + static_cast<long long>(b) * p.o_batch_stride +
                     static_cast<long long>(qrow) * p.ldo + h * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// called by rfb_attention (attention.cu) for mode 0
int launch_attention2(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                      const rfb_attn_args* a, int k_batched, int v_batched, int split_tiles, int n_splits,
                      float* part_o, float* part_ml, cudaStream_t stream) {
  Attn2Params p{};
  p.B = a->B, p.split_tiles = n_splits > 1 ? split_tiles : 0;
  p.part_o = n_splits > 1 ? part_o : nullptr, p.part_ml = n_splits > 1 ? part_ml : nullptr;
  p.Nq = a->Nq, p.Nk = a->Nk;
  p.n_kv_tiles = (a->Nk + 127) / 128;
  p.k_batched = k_batched, p.v_batched = v_batched;
  p.mask_bits = a->key_mask_bits;
  p.mask_stride_words = a->mask_batch_stride_words;
  p.O = a->O, p.ldo = a->ldo, p.o_batch_stride = a->o_batch_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.q_sumsq = a->q_sumsq, p.sumsq_ld = a->sumsq_ld > 0 ? a->sumsq_ld : 1;
  p.sumsq_parts = a->sumsq_parts > 0 ? a->sumsq_parts : 1;
  p.inv_norm_dim = a->norm_dim > 0 ? 1.0f / (float)a->norm_dim : 0.f, p.norm_eps = a->norm_eps;
  const bool f16 = a->dtype == RFB_F16;
  auto kern = f16 ? attn2_tc_kernel<true> : attn2_tc_kernel<false>;
  static PerDeviceFlag attr_flags[2];
  bool& attr_set = attr_flags[f16].get();
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2Smem) !=
        cudaSuccess)
      return RFB_ERR_LAUNCH;
    // setmaxnreg moves registers inside the CTA's launch allocation only: the softmax warpgroups'
    // request must be covered by what warpgroup 0 gives back, or the kernel would wait forever
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return RFB_ERR_LAUNCH;
    if (128 * kRegsWg0 + 256 * kRegsSoftmax > kAttn2Threads * fa.numRegs) {
      fprintf(stderr, "rfb: attn2_tc_kernel compiled with %d registers/thread; setmaxnreg split %d/%d does not fit\n",
              fa.numRegs, kRegsWg0, kRegsSoftmax);
      return RFB_ERR_LAUNCH;
    }
    attr_set = true;
  }
  dim3 grid((a->Nq + 255) / 256, a->H, a->B * (n_splits > 1 ? n_splits : 1));
  kern<<<grid, kAttn2Threads, kAttn2Smem, stream>>>(tmQ, tmK, tmV, p);
  g_launch_count++;
  return check_launch("attn2_tc_kernel");
}

}  // namespace rfb
