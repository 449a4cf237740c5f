/* microdit_b200 -- C ABI of the B200 (sm_100a) MicroDiT training hot path.
 *
 * Every entry point takes plain device pointers, integer sizes and a CUDA stream handle
 * (cudaStream_t passed as void*), returns 0 on success or a negative MD_ERR_* code, never allocates
 * device memory and keeps no global state besides cached function attributes; all calls are
 * asynchronous on the given stream and re-entrant per stream.  md_last_error() returns the message of
 * the last failure on the calling thread.
 *
 * The reference (SonyResearch/micro_diffusion) has no FFI: its hot path is Python calling torch ops.
 * Each function below names the reference code it replaces (file:line under the reference root).
 * The Python binding a maintainer would add is the ctypes stub in INTEGRATION.md.
 */
#ifndef MICRODIT_B200_H_
#define MICRODIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(MD_BUILDING_LIB)
#define MD_API __attribute__((visibility("default")))
#else
#define MD_API
#endif

#define MD_OK 0
#define MD_ERR_INVALID (-1)     /* bad argument (null pointer, misaligned operand, bad size) */
#define MD_ERR_CUDA (-2)        /* CUDA runtime / driver error */
#define MD_ERR_UNSUPPORTED (-3) /* device is not sm_100a, or shape outside the kernel's envelope */

MD_API const char* md_last_error(void);
MD_API int md_abi_version(void);

/* ------------------------------------------------------------------------------------------ GEMM */
/* layouts */
#define MD_GEMM_NT 0 /* C[M,N] = A[M,K] . B[N,K]^T   (A, B row-major, K contiguous)                 */
#define MD_GEMM_TN 1 /* C[M,N] = A[K,M]^T . B[K,N]   (A, B row-major, reduction index K strided)    */
/* epilogues */
#define MD_EPI_STORE_BF16 0 /* C(bf16) = alpha*acc (+bias)                                           */
#define MD_EPI_STORE_F32 1  /* C(f32)  = alpha*acc (+bias)                                           */
#define MD_EPI_RESID_F32 2  /* C(f32)  = res + gate[row/rows_per_gate] * (alpha*acc+bias);           */
                            /*           C2(bf16, optional) = alpha*acc+bias                         */
#define MD_EPI_ATOMIC_F32 3 /* C(f32) += alpha*acc   (red.global.add; the only mode allowing splits) */
#define MD_EPI_GELU_DUAL 4  /* C(bf16) = pre = alpha*acc+bias ; C2(bf16) = gelu_erf(pre)             */

typedef struct md_gemm_args {
  const void* A; /* bf16 */
  const void* B; /* bf16 */
  void* C;
  void* C2;
  const void* bias; /* f32 [batch][N] or NULL */
  const void* res;  /* f32, same indexing as C (may alias C) */
  const void* gate; /* f32 [M / rows_per_gate][ldgate] or NULL (=1) */
  int64_t M, N, K;
  int64_t lda, ldb, ldc; /* row pitches in elements */
  int64_t batch;         /* >= 1; batch strides in elements */
  int64_t strideA, strideB, strideC, strideBias;
  int64_t ldgate, rows_per_gate;
  int32_t layout;   /* MD_GEMM_* */
  int32_t epilogue; /* MD_EPI_*  */
  int32_t splits;   /* split of the reduction dimension (>=1) */
  float alpha;      /* 0 is treated as 1 */
} md_gemm_args;

/* Dense / batched bf16 GEMM with fp32 accumulation on tcgen05 tensor cores.
 * Replaces: nn.Linear under autocast -- qkv/proj utils.py:172-173, cross-attn q/kv/proj utils.py:109-111,
 * SwiGLU dit.py:84-89, adaLN dit.py:227-230, stem/mixer maps dit.py:377-388, final linear utils.py:226-230 --
 * the expert einsums dit.py:135-137 (batch = experts), and the autograd dgrad/wgrad of all of them. */
MD_API int md_gemm_bf16(const md_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MICRODIT_B200_H_ */
