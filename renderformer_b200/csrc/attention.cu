// Flash-style attention on tcgen05 / TMEM for head_dim = 128 (sm_100a).
//
//   O = softmax(scale * Q K^T + mask) V        per (batch, head, 128-query tile)
//
// One CTA per query tile.  Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = softmax (one thread per query row).  S = Q K^T is double-buffered in TMEM
// (2 x 128 fp32 columns), O accumulates in TMEM (128 columns); P is written as bf16 into a
// 128B-swizzled smem tile and fed back as the A operand of P V.  V is consumed transposed
// ([head_dim, keys], keys contiguous) so every operand is K-major.
//
// mode 0: every query tile visits all key tiles; keys are masked by a packed bit mask
//         (the key-padding mask of layers/attention.py:145-161, True = attend).
// mode 1: block-diagonal -- query tile i only sees key tile i; a tile holds two 64-token swin
//         windows and tokens attend iff they sit in the same window (tile half) and carry the
//         same shifted-window region id (layers/attention.py:238-271,327-358).
#include <stdlib.h>

#include <atomic>

#include "host_util.h"
#include "ptx.cuh"

namespace rfb {

extern std::atomic<long long> g_launch_count;

int launch_attention_swin(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                          const rfb_attn_args* a, int k_batched, int v_batched, cudaStream_t stream);
int launch_attention3(const CUtensorMap& tmK, const CUtensorMap& tmV, const rfb_attn_args* a, int k_batched,
                      int v_batched, cudaStream_t stream);
int launch_attention2(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                      const rfb_attn_args* a, int k_batched, int v_batched, cudaStream_t stream);

struct AttnKParams {
  int Nq, Nk, H;
  int n_kv_tiles;
  int mode;
  int k_batched, v_batched;  // 0: batch coordinate pinned to 0 (shared across the grid's batch dim)
  const uint32_t* mask_bits;
  long long mask_stride_words;
  const uint8_t* group_id;
  int group_period;
  void* O;
  long long ldo, o_batch_stride;
  float scale_log2;
  const float* q_sumsq;
  const float* k_sumsq;
  int sumsq_ld, sumsq_parts;
  float inv_norm_dim, norm_eps;
};

// 1/rms of one row from its partial sums of squares
__device__ __forceinline__ float rms_factor(const float* sumsq, long long row, int ld, int parts, float inv_dim,
                                            float eps) {
  const float* sp = sumsq + row * ld;
  float ss = 0.f;
  for (int j = 0; j < parts; ++j) ss += sp[j];
  return rsqrtf(ss * inv_dim + eps);
}

constexpr uint32_t kTileBytes = 128 * 128 * 2;  // one 128 x 128 bf16 operand tile (two SW128 halves)
constexpr uint32_t kHalfBytes = 128 * 64 * 2;
constexpr int kKVStages = 2;
constexpr uint32_t kAttnSmem = kTileBytes * (1 + 2 * kKVStages + 1) + 1024 + 1024;

__global__ void __launch_bounds__(192, 1)
    attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kKVStages * kTileBytes;
  uint8_t* sP = sV + kKVStages * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kTileBytes);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = q_full + 1;            // kKVStages
  uint64_t* k_empty = k_full + kKVStages;   // kKVStages
  uint64_t* v_full = k_empty + kKVStages;   // kKVStages
  uint64_t* v_empty = v_full + kKVStages;   // kKVStages
  uint64_t* s_full = v_empty + kKVStages;   // 2
  uint64_t* s_empty = s_full + 2;           // 2
  uint64_t* p_full = s_empty + 2;           // 1
  uint64_t* pv_done = p_full + 1;           // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  uint8_t* s_gid = reinterpret_cast<uint8_t*>(tmem_slot + 2);  // 128 group ids (mode 1)
  float* s_rk = reinterpret_cast<float*>(s_gid + 128);          // 128 per-key 1/rms factors (mode 1)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * 128;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&k_full[i], 1), mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1), mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&s_full[i], 1), mbar_init(&s_empty[i], 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ), tma_prefetch_desc(&tmK), tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (p.mode == 1 && threadIdx.x >= 64) {
    const int i = threadIdx.x - 64;
    s_gid[i] = p.group_id[(q0 + i) % p.group_period];
    float rk = 1.0f;
    if (p.k_sumsq && q0 + i < p.Nk)
      rk = rms_factor(p.k_sumsq, q0 + i, p.sumsq_ld, p.sumsq_parts, p.inv_norm_dim, p.norm_eps);
    s_rk[i] = rk;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;         // + 128 * (j & 1)
  const uint32_t tO = tmem_base + 256;

  const int n_tiles = (p.mode == 1) ? 1 : p.n_kv_tiles;
  const int kb = p.k_batched ? b : 0;
  const int vb = p.v_batched ? b : 0;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tmQ, q_full, h * 128, q0, b);
      tma_load_3d(sQ + kHalfBytes, &tmQ, q_full, h * 128 + 64, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        const int key0 = ((p.mode == 1) ? qt : j) * 128;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kTileBytes);
        tma_load_3d(sK + s * kTileBytes, &tmK, &k_full[s], h * 128, key0, kb);
        tma_load_3d(sK + s * kTileBytes + kHalfBytes, &tmK, &k_full[s], h * 128 + 64, key0, kb);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kTileBytes);
        tma_load_3d(sV + s * kTileBytes, &tmV, &v_full[s], key0, h * 128, vb);
        tma_load_3d(sV + s * kTileBytes + kHalfBytes, &tmV, &v_full[s], key0 + 64, h * 128, vb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(1u, 128, 128);
      auto issue_qk = [&](int j) {
        const int s = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        const int sb = j & 1;
        mbar_wait(&k_full[s], ph);
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t ad = umma_desc_sw128(smem_u32(sQ));
        const uint64_t bd = umma_desc_sw128(smem_u32(sK + s * kTileBytes));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kHalfBytes >> 4) + (k & 3) * 2;
          umma_f16(tS + sb * 128, ad + off, bd + off, idesc, k != 0);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[sb]);
      };
      auto issue_pv = [&](int j) {
        const int s = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        mbar_wait(&v_full[s], ph);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t ad = umma_desc_sw128(smem_u32(sP));
        const uint64_t bd = umma_desc_sw128(smem_u32(sV + s * kTileBytes));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kHalfBytes >> 4) + (k & 3) * 2;
          umma_f16(tO, ad + off, bd + off, idesc, (j | k) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        issue_pv(j);
      }
    }
  } else {
    // ------------------------------ softmax warps ------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float sl2 = p.scale_log2;
    if (p.q_sumsq && q0 + r < p.Nq)
      sl2 *= rms_factor(p.q_sumsq, static_cast<long long>(b) * p.Nq + q0 + r, p.sumsq_ld, p.sumsq_parts,
                        p.inv_norm_dim, p.norm_eps);
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (p.mode == 1) {
      const uint8_t g = s_gid[r];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        uint32_t bits = 0;
        for (int i = 0; i < 32; ++i)  // same 64-token window (tile half) and same region id
          bits |= ((s_gid[w * 32 + i] == g && ((w * 32 + i) >> 6) == (r >> 6)) ? 1u : 0u) << i;
        mw[w] = bits;
      }
      const int rem = p.Nk - q0;
      if (rem < 128) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int lo = w * 32;
          const uint32_t keep = rem <= lo ? 0u : (rem - lo >= 32 ? 0xffffffffu : ((1u << (rem - lo)) - 1u));
          mw[w] &= keep;
        }
      }
    }
    uint8_t* prow = sP + r * 128;
    const uint32_t rx = r & 7;

    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      if (p.mode == 0) {
        if (p.mask_bits) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(
              p.mask_bits + static_cast<long long>(b) * p.mask_stride_words + j * 4));
          mw[0] = u.x, mw[1] = u.y, mw[2] = u.z, mw[3] = u.w;
        } else {
          const int rem = p.Nk - j * 128;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int lo = w * 32;
            mw[w] = rem <= lo ? 0u : (rem - lo >= 32 ? 0xffffffffu : ((1u << (rem - lo)) - 1u));
          }
        }
      }
      const bool all_valid = (mw[0] & mw[1] & mw[2] & mw[3]) == 0xffffffffu;

      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tS + sb * 128 + lane_addr;

      // the whole 128-column row of S goes to registers in one shot; the TMEM buffer is
      // released to the MMA warp immediately
      uint32_t v[4][32];
      tmem_ld32(ts, v[0]);
      tmem_ld32(ts + 32, v[1]);
      tmem_ld32(ts + 64, v[2]);
      tmem_ld32(ts + 96, v[3]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);

      if (p.mode == 1 && p.k_sumsq) {  // fused k RMSNorm: per-key 1/rms
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c][i] = __float_as_uint(__uint_as_float(v[c][i]) * s_rk[c * 32 + i]);
      }
      if (!all_valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t bits = mw[c];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!((bits >> i) & 1u)) v[c][i] = 0xff800000u;  // -inf
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[c][i]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_f((m_run - m_use) * sl2);
      const float neg_ms = -m_use * sl2;

      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // O stable, P tile free
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + lane_addr + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + lane_addr + c * 32, o);
          }
          tmem_wait_st();
        }
      }

      // P = exp2(s*sl2 - m*sl2) -> bf16 -> swizzled smem
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint8_t* dst = prow + (c >> 1) * kHalfBytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float pv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pv[e] = ex2_f(fmaf(__uint_as_float(v[c][i * 8 + e]), sl2, neg_ms));
            rs4[e & 3] += pv[e];
          }
          uint4 u;
          u.x = pack_bf16(pv[0], pv[1]);
          u.y = pack_bf16(pv[2], pv[3]);
          u.z = pack_bf16(pv[4], pv[5]);
          u.w = pack_bf16(pv[6], pv[7]);
          const uint32_t chunk = ((c & 1) * 4 + i) ^ rx;
          *reinterpret_cast<uint4*>(dst + chunk * 16) = u;
        }
      }
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      m_run = m_new;

      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // epilogue: O / l -> bf16 -> global
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = (l_run > 0.f) ? 1.0f / l_run : 0.f;
    const bool row_ok = (q0 + r) < p.Nq;
    uint16_t* orow = static_cast<uint16_t*>(p.O) + static_cast<long long>(b) * p.o_batch_stride +
                     static_cast<long long>(q0 + r) * p.ldo + h * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tO + lane_addr + c * 32, v);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[i * 8 + 0]) * inv_l, __uint_as_float(v[i * 8 + 1]) * inv_l);
          u.y = pack_bf16(__uint_as_float(v[i * 8 + 2]) * inv_l, __uint_as_float(v[i * 8 + 3]) * inv_l);
          u.z = pack_bf16(__uint_as_float(v[i * 8 + 4]) * inv_l, __uint_as_float(v[i * 8 + 5]) * inv_l);
          u.w = pack_bf16(__uint_as_float(v[i * 8 + 6]) * inv_l, __uint_as_float(v[i * 8 + 7]) * inv_l);
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

}  // namespace rfb

using namespace rfb;

extern "C" int rfb_attention(const rfb_attn_args* a, rfb_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !a->Q || !a->K || !a->Vt || !a->O || a->B <= 0 || a->H <= 0 || a->Nq <= 0 || a->Nk <= 0)
    return RFB_ERR_ARG;
  if (a->mode == 1 && (!a->group_id || a->group_period <= 0 || a->Nq != a->Nk)) return RFB_ERR_ARG;
  if (a->ldo % 8) return RFB_ERR_ALIGN;

  const uint64_t hd_cols = (uint64_t)a->H * 128;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {hd_cols, (uint64_t)a->Nq, (uint64_t)a->B};
    uint64_t st[2] = {(uint64_t)a->ldq * 2, (uint64_t)(a->B > 1 ? a->q_batch_stride : (long long)a->Nq * a->ldq) * 2};
    if ((rc = make_tmap_16b(&tmQ, RFB_BF16, a->Q, 3, dims, st, box)) != RFB_OK) return rc;
  }
  const int k_batched = (a->B > 1 && a->k_batch_stride != 0);
  const int v_batched = (a->B > 1 && a->vt_batch_stride != 0);
  {
    uint64_t dims[3] = {hd_cols, (uint64_t)a->Nk, (uint64_t)(k_batched ? a->B : 1)};
    uint64_t st[2] = {(uint64_t)a->ldk * 2, (uint64_t)(k_batched ? a->k_batch_stride : (long long)a->Nk * a->ldk) * 2};
    if ((rc = make_tmap_16b(&tmK, RFB_BF16, a->K, 3, dims, st, box)) != RFB_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->Nk, hd_cols, (uint64_t)(v_batched ? a->B : 1)};
    uint64_t st[2] = {(uint64_t)a->ldvt * 2, (uint64_t)(v_batched ? a->vt_batch_stride : (long long)hd_cols * a->ldvt) * 2};
    if ((rc = make_tmap_16b(&tmV, RFB_BF16, a->Vt, 3, dims, st, box)) != RFB_OK) return rc;
  }

  if (a->mode == 0) {
    static int forced = -1;  // RFB_ATTN_GEN = 1 | 2 | 3 forces one dense kernel generation (A/B runs)
    if (forced < 0) {
      const char* e = getenv("RFB_ATTN_GEN");
      forced = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 0;
    }
    int gen = forced;
    if (gen == 0) {
      // Both kernels sustain the same rate per tile; what differs is how evenly their grids fill
      // the SMs (one CTA per SM): two-tile CTAs (gen 2) vs one-tile CTAs (gen 3).
      const int sms = num_sms();
      const long long n2 = (long long)((a->Nq + 255) / 256) * a->H * a->B;
      const long long n3 = (long long)((a->Nq + 127) / 128) * a->H * a->B;
      const double e2 = (double)n2 / (double)(((n2 + sms - 1) / sms) * sms);
      const double e3 = (double)n3 / (double)(((n3 + sms - 1) / sms) * sms);
      gen = (e3 > e2 + 0.02) ? 3 : 2;
    }
    if (gen == 3) return launch_attention3(tmK, tmV, a, k_batched, v_batched, stream);
    if (gen == 2) return launch_attention2(tmQ, tmK, tmV, a, k_batched, v_batched, stream);
  } else {
    static int swin_v1 = -1;  // RFB_SWIN_V1=1: per-(tile, head) CTAs of attn_tc_kernel (A/B reference)
    if (swin_v1 < 0) {
      const char* e = getenv("RFB_SWIN_V1");
      swin_v1 = (e && e[0] == '1') ? 1 : 0;
    }
    if (!swin_v1) return launch_attention_swin(tmQ, tmK, tmV, a, k_batched, v_batched, stream);
  }

  AttnKParams p{};
  p.Nq = a->Nq, p.Nk = a->Nk, p.H = a->H;
  p.n_kv_tiles = (a->Nk + 127) / 128;
  p.mode = a->mode;
  p.k_batched = k_batched, p.v_batched = v_batched;
  p.mask_bits = a->key_mask_bits;
  p.mask_stride_words = a->mask_batch_stride_words;
  p.group_id = a->group_id, p.group_period = a->group_period;
  p.O = a->O, p.ldo = a->ldo, p.o_batch_stride = a->o_batch_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.q_sumsq = a->q_sumsq, p.k_sumsq = a->k_sumsq, p.sumsq_ld = a->sumsq_ld > 0 ? a->sumsq_ld : 1;
  p.sumsq_parts = a->sumsq_parts > 0 ? a->sumsq_parts : 1;
  p.inv_norm_dim = a->norm_dim > 0 ? 1.0f / (float)a->norm_dim : 0.f, p.norm_eps = a->norm_eps;

  static PerDeviceFlag attr_flags;
  bool& attr_set = attr_flags.get();
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem) !=
        cudaSuccess)
      return RFB_ERR_LAUNCH;
    attr_set = true;
  }
  dim3 grid((a->Nq + 127) / 128, a->H, a->B);
  attn_tc_kernel<<<grid, 192, kAttnSmem, stream>>>(tmQ, tmK, tmV, p);
  g_launch_count++;
  return check_launch("attn_tc_kernel");
}
