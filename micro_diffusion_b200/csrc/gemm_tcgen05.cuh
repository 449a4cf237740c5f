#pragma once
#include "common.cuh"

namespace md {
enum {
  EPI_STORE_BF16 = MD_EPI_STORE_BF16,
  EPI_STORE_F32 = MD_EPI_STORE_F32,
  EPI_RESID_F32 = MD_EPI_RESID_F32,
  EPI_ATOMIC_F32 = MD_EPI_ATOMIC_F32,
  EPI_ACT_DUAL = MD_EPI_ACT_DUAL,
  EPI_ACT_GRAD = MD_EPI_ACT_GRAD,
  EPI_SWIGLU = MD_EPI_SWIGLU,
  EPI_SWIGLU_GRAD = MD_EPI_SWIGLU_GRAD,
  EPI_COUNT
};
// Kernel-side argument block (everything the device needs besides the two tensor maps).
struct GemmDev {
  void* C;
  void* C2;
  const float* bias;
  const float* res;
  const float* gate;
  const void* aux;
  float* split_ws;  // deterministic split-K: partial tiles [split][batch][M][N] instead of atomics
  long long ldc, strideC, strideBias, ldgate, ldc2, strideC2;
  int M, N, K, batch, splits, rows_per_gate, epi, res_mod, act;
  int row_interleave;  // > 0: output row p of the atomic epilogue goes to the parameter row of the 32-row-interleaved stack
  int debug;      // diagnostics only (MD_GEMM_DEBUG): 1 = epilogue skips its stores, 2 = also skips the TMEM loads
  int tma_store;  // bf16 store epilogue goes through smem staging + cp.async.bulk.tensor (tmC valid)
  float alpha;
};
}  // namespace md
