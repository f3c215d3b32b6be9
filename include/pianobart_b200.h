/* pianobart_b200 - C ABI of the sm_100a kernels behind the PianoBART hot path.
 *
 * The reference (RS2002/PianoBart) is pure Python and has no FFI layer: its "operator
 * interface" for this path is the Python module API (PianoBart.forward PianoBart.py:56-80,
 * PianoBartLM.forward model.py:20-66, Pretrainer.gen_mask / compute_loss pretrain.py:112-118,
 * 211-546).  The host-side mirror of that API lives in pianobart_b200/*.py; every device
 * operation it performs is one of the entry points below, bound with ctypes
 * (pianobart_b200/_lib.py).  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends
 *     in _host; no torch types; the caller owns every buffer (nothing is allocated here).
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered and asynchronous.
 *   - return value 0 = success; non-zero = failure, message via pb_last_error().
 *   - dtype: 0 = fp32 buffers (parity mode), 1 = bf16 buffers (production mode).
 *   - there is NO CPU fallback: without a CUDA device these functions fail.
 */
#ifndef PIANOBART_B200_H
#define PIANOBART_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_DTYPE_F32 0
#define PB_DTYPE_BF16 1

/* ------------------------------------------------------------------ misc */
const char* pb_last_error(void);
int pb_version(void);
/* number of kernels launched by this library since the last reset (bench: gpu_launches) */
long long pb_launch_count(void);
void pb_reset_launch_count(void);

/* ------------------------------------------------------------------ GEMM
 * C[b,h][m,n] = epi(alpha * sum_k A[b,h][m,k] * B[b,h][n,k])
 * Replaces every nn.Linear / torch.bmm on the path: PianoBart.py:68,71 (in_linear),
 * HF modeling_bart.py BartAttention q/k/v/out_proj + eager_attention_forward matmuls,
 * BartEncoderLayer/BartDecoderLayer fc1/fc2, model.py:119-126 (8 heads as one N=1280 GEMM),
 * and autograd's backward products for them.
 *
 * a_mn_major = 0: A stored [M][K] (k contiguous), lda = row stride in elements
 * a_mn_major = 1: A stored [K][M] (m contiguous), lda = row stride in elements   (same for B / N)
 */
#define PB_GEMM_OUT_F32 1    /* C is fp32 (default: same dtype as the inputs)                */
#define PB_GEMM_GELU 2       /* exact erf GELU after bias (HF ACT2FN["gelu"])                */
#define PB_GEMM_ATOMIC_ACC 4 /* C += result with fp32 atomics (grad accumulation, split-K)   */
#define PB_GEMM_RES_F32 8    /* residual is fp32 (default: same dtype as C)                  */
#define PB_GEMM_AUX_PREACT 16 /* also store the pre-activation (after bias) to aux (dtype of C) */
#define PB_GEMM_MUL_DGELU 32  /* multiply the result by gelu'(aux[m,n]) (fc1 backward)          */

typedef struct pb_gemm_desc {
  const void* a;
  const void* b;
  void* c;
  const float* bias;    /* [N] fp32 or NULL */
  const void* residual; /* added after activation, or NULL */
  int M, N, K;
  int a_mn_major, b_mn_major;
  long long lda, ldb, ldc, ldr;
  int batch_h, batch_b; /* two batch dimensions (0/1 = none) */
  long long a_stride_h, a_stride_b, b_stride_h, b_stride_b;
  long long c_stride_h, c_stride_b, r_stride_h, r_stride_b;
  float alpha;
  int flags;
  int split_k; /* >1 requires OUT_F32|ATOMIC_ACC */
  int causal;  /* 0 none; 1 skip output tiles with n > m (scores); 2 limit k to <= m (P.V)      */
  int block_n; /* 0 = auto, else 128 or 256 (tcgen05 path only)                                 */
  void* aux;   /* see PB_GEMM_AUX_PREACT / PB_GEMM_MUL_DGELU; row stride ldaux, not batched      */
  long long ldaux;
  int r_row_mod; /* >0: residual row index = m % r_row_mod (learned position table broadcast over batch) */
} pb_gemm_desc;

/* bf16 operands, tcgen05/TMEM/TMA kernel */
int pb_gemm_bf16(const pb_gemm_desc* d, void* stream);
/* fp32 operands, SIMT kernel - the fp32 parity mode of the same graph */
int pb_gemm_f32(const pb_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif
