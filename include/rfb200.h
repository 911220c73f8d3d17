/*
 * rfb200.h -- C ABI of the B200-native RenderFormer inference kernels.
 *
 * The reference (agwi-lab/renderformer) has no FFI layer: its hot path is Python
 * calling torch / flash-attn / cuDNN.  Each entry point below replaces one family of
 * those call sites (cited as reference file:line, relative to the reference root).
 * All pointers are device pointers unless stated; all functions are asynchronous on
 * `stream`, never allocate, and return RFB_OK (0) or a negative error code.
 * No torch types cross this boundary.
 */
#ifndef RFB200_H_
#define RFB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rfb_stream_t; /* cudaStream_t */

enum {
  RFB_OK = 0,
  RFB_ERR_ARG = -1,    /* bad argument / unsupported shape */
  RFB_ERR_ALIGN = -2,  /* pointer or stride not 16-byte aligned */
  RFB_ERR_DRIVER = -3, /* CUDA driver entry point unavailable (no GPU driver) */
  RFB_ERR_TMAP = -4,   /* cuTensorMapEncodeTiled rejected the descriptor */
  RFB_ERR_LAUNCH = -5  /* kernel launch failed */
};

/* element types */
enum { RFB_F32 = 0, RFB_BF16 = 1, RFB_F16 = 2 };

int rfb_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long rfb_launch_count(void);

/* ---------------------------------------------------------------------------------
 * rfb_gemm: C[M,N] = A[M,K] * W[N,K]^T  (16-bit operands, fp32 accumulate in TMEM;
 * TMA-fed tcgen05.mma, persistent, fused epilogue).
 *
 * Replaces every nn.Linear on the path -- layers/attention.py:51-57 (SwiGLU w1,w2,w3),
 * :95-100,120-125,202 (in_proj/q_proj/k_proj/v_proj/out_proj), :300-301,343,361 (swin),
 * models/renderformer.py:49,63,139-147 (token encoders), models/view_transformer.py:45,108
 * -- and, with a_mode = RFB_A_CONV3X3, the DPT convolutions layers/dpt.py:44-52,68-71,
 * 184-240 as implicit GEMM (1x1 convs and ConvTranspose k=s are plain linear GEMMs).
 * ------------------------------------------------------------------------------- */
enum { RFB_A_LINEAR = 0, RFB_A_CONV3X3 = 1 };
enum {
  RFB_EPI_STORE = 0,  /* v = acc (+bias) (+res1) (+res2); out = v; out_act = silu(v) */
  RFB_EPI_SWIGLU = 1, /* W rows interleaved [16 gate | 16 up] per 32: out[:, n/2] = silu(g)*u */
  RFB_EPI_FINAL = 2   /* N==32: y = w2 * silu(acc+bias) + b2 (3 ch); out = 10^elu(y) - 1, fp32 */
};

typedef struct {
  int M, N, K;   /* conv: M = B*H*W output pixels, K = 9*Cin */
  const void* A; /* [M,K] row-major (lda elements) | conv: NHWC [B,H,W,Cin] */
  long long lda;
  const void* W; /* [N,K] row-major (ldw elements); conv: K index = tap*Cin + cin, tap = ky*3+kx */
  long long ldw;
  int dtype; /* operand type: RFB_BF16 or RFB_F16 */
  int a_mode;
  int B, H, Wd, Cin; /* conv geometry (stride 1, pad 1) */
  int epi;
  const float* bias; /* [N] or NULL */
  const void* res1;  /* [M, ldres] addend or NULL */
  const void* res2;
  int res_dtype;
  long long ldres;
  void* out; /* [M, ldo] or NULL */
  int out_dtype;
  long long ldo;
  void* out_act;      /* optional silu(v) copy, same dtype/ld as out */
  const int* row_map; /* optional: output/residual row = row_map[m] (swin window order -> token order) */
  const float* w2;    /* RFB_EPI_FINAL: [3,32] */
  const float* b2;    /* RFB_EPI_FINAL: [3] */
  int bn_override;    /* 0 = auto; else force the N tile (32/64/128/256) */
  int max_ctas;       /* 0 = one per SM */
} rfb_gemm_args;

int rfb_gemm(const rfb_gemm_args* a, rfb_stream_t stream);

/* ---------------------------------------------------------------------------------
 * rfb_attention: O = softmax(scale * Q K^T + mask) V, head_dim 128, bf16 in/out, fp32
 * softmax and accumulation (tcgen05 flash-style kernel, S and O live in TMEM).
 *
 * Replaces F.scaled_dot_product_attention / flash_attn_* at layers/attention.py:143-198
 * (encoder self-attention and decoder cross-attention, key-padding mask fused) and the
 * per-window SDPA of SwinSelfAttention at layers/attention.py:349-358 (mode 1: the
 * window partition and the shifted-window region mask become a group-id equality test).
 *
 * Layouts: Q [B][Nq][ldq], K [B][Nk][ldk] with head h in columns [h*128, h*128+128);
 * Vt [B][H*128][ldvt] = V transposed (keys contiguous).  A batch stride of 0 on K / Vt /
 * mask shares that tensor across the batch (hoisted per-scene K/V).
 * ------------------------------------------------------------------------------- */
typedef struct {
  int B, H, Nq, Nk;
  const void* Q;
  long long ldq, q_batch_stride;
  const void* K;
  long long ldk, k_batch_stride;
  const void* Vt;
  long long ldvt, vt_batch_stride;
  void* O;
  long long ldo, o_batch_stride;
  const uint32_t* key_mask_bits; /* mode 0: [B][4*ceil(Nk/128)] packed bits, 1 = attend; NULL = all */
  long long mask_batch_stride_words;
  int mode;                /* 0 = dense over all keys, 1 = block-diagonal 128-token tiles */
  const uint8_t* group_id; /* mode 1: [group_period]; i,j attend iff ids equal */
  int group_period;
  float scale;
} rfb_attn_args;

int rfb_attention(const rfb_attn_args* a, rfb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RFB200_H_ */
