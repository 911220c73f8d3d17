// rfb_attention: entry point of the flash-style attention kernels (tcgen05 / TMEM, head_dim = 128, sm_100a).
//
//   O = softmax(scale * Q K^T + mask) V        per (batch, head, 128-query tile)
//
// The kernels live in attention2.cu (two query tiles per CTA), attention3.cu (one query tile per CTA, Q in
// TMEM) and attention_swin.cu (block-diagonal shifted-window mode); this file builds the TMA descriptors
// and picks the kernel.  Operands are bf16 or fp16 (`dtype`), V is consumed transposed ([head_dim, keys],
// keys contiguous) so every operand is K-major.
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

}  // namespace rfb

using namespace rfb;

extern "C" int rfb_attention(const rfb_attn_args* a, rfb_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !a->Q || !a->K || !a->Vt || !a->O || a->B <= 0 || a->H <= 0 || a->Nq <= 0 || a->Nk <= 0)
    return RFB_ERR_ARG;
  if (a->mode == 1 && (!a->group_id || a->group_period <= 0 || a->Nq != a->Nk)) return RFB_ERR_ARG;
  if (a->ldo % 8) return RFB_ERR_ALIGN;
  if (a->dtype != 0 && a->dtype != RFB_BF16 && a->dtype != RFB_F16) return RFB_ERR_ARG;
  const int dt = a->dtype == RFB_F16 ? RFB_F16 : RFB_BF16;  // 0 (unset) = bf16

  const uint64_t hd_cols = (uint64_t)a->H * 128;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {hd_cols, (uint64_t)a->Nq, (uint64_t)a->B};
    uint64_t st[2] = {(uint64_t)a->ldq * 2, (uint64_t)(a->B > 1 ? a->q_batch_stride : (long long)a->Nq * a->ldq) * 2};
    if ((rc = make_tmap_16b(&tmQ, dt, a->Q, 3, dims, st, box)) != RFB_OK) return rc;
  }
  const int k_batched = (a->B > 1 && a->k_batch_stride != 0);
  const int v_batched = (a->B > 1 && a->vt_batch_stride != 0);
  {
    uint64_t dims[3] = {hd_cols, (uint64_t)a->Nk, (uint64_t)(k_batched ? a->B : 1)};
    uint64_t st[2] = {(uint64_t)a->ldk * 2, (uint64_t)(k_batched ? a->k_batch_stride : (long long)a->Nk * a->ldk) * 2};
    if ((rc = make_tmap_16b(&tmK, dt, a->K, 3, dims, st, box)) != RFB_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->Nk, hd_cols, (uint64_t)(v_batched ? a->B : 1)};
    uint64_t st[2] = {(uint64_t)a->ldvt * 2, (uint64_t)(v_batched ? a->vt_batch_stride : (long long)hd_cols * a->ldvt) * 2};
    if ((rc = make_tmap_16b(&tmV, dt, a->Vt, 3, dims, st, box)) != RFB_OK) return rc;
  }

  if (a->mode == 0) {
    static int forced = -1;  // RFB_ATTN_GEN = 2 | 3 forces one dense kernel generation (A/B runs)
    if (forced < 0) {
      const char* e = getenv("RFB_ATTN_GEN");
      forced = (e && e[0] >= '2' && e[0] <= '3') ? e[0] - '0' : 0;
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
    return launch_attention2(tmQ, tmK, tmV, a, k_batched, v_batched, stream);
  }
  return launch_attention_swin(tmQ, tmK, tmV, a, k_batched, v_batched, stream);
}
