// EXPERIMENTAL (round 2): attention forward on the 5th-generation tensor cores for head_dim 64 and Tk <= 256 --
// every sequence length of the res-256 configs (64, 77 -> 80, 256).  Written and compiled in round 1, NOT yet run on
// hardware and NOT dispatched by md_attn_fwd: the only entry point is md_attn_fwd_tc, which CudaOps calls solely under
// MD_ATTN_TC=1 (tests/test_attn_tc_gpu.py is skipped without it).  See DESIGN.md section 8 for the plan.
//
// Same contract as md_attn_fwd (F.scaled_dot_product_attention at reference utils.py:188-193 / 127-132; lse in the
// log2 domain).  One CTA = 128 queries of one (sample, head):
//   warp 0  : TMEM alloc; lane 0 issues the TMA loads (Q, K: K-major boxes; V: 64-key boxes used as an MN-major B
//             operand) and the two UMMA chains  S[128 x Tk] = Q . K^T  and  O[128 x 64] = P . V
//   warps 1-4: one query row per thread.  Whole-row softmax straight from TMEM (two passes of tcgen05.ld.32x32b.x32:
//             max, then exp2 / row sum) -- no online rescaling because all of S fits the 256 TMEM columns; P goes to
//             shared memory as bf16 in 128B-swizzled K-major atoms (the A operand of the second chain); the same
//             warps read O back, divide by the row sum and store.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace md {
namespace attn_tc0 {

constexpr int kQ = 128;        // queries per CTA (UMMA M)
constexpr int kHd = 64;        // head_dim == one 128-byte swizzle row of bf16
constexpr int kMaxKeys = 256;  // S must fit 256 fp32 TMEM columns
constexpr int kThreads = 160;  // 1 control warp + 4 softmax warps
constexpr int kTmemCols = 512; // S: columns [0, 256), O: [256, 320)
constexpr int kOCol = 256;

constexpr int kBytesQ = kQ * kHd * 2;                 // 16 KB
constexpr int kBytesKBox = 128 * kHd * 2;             // 16 KB (128 keys)
constexpr int kBytesVBox = 64 * kHd * 2;              // 8 KB  (64 keys)
constexpr int kBytesPAtom = kQ * 64 * 2;              // 16 KB (128 queries x 64 keys)
constexpr int kOffK = kBytesQ;
constexpr int kOffV = kOffK + 2 * kBytesKBox;
constexpr int kOffP = kOffV + 4 * kBytesVBox;
constexpr int kOffBar = kOffP + 4 * kBytesPAtom;
constexpr int kSmemBytes = kOffBar + 128 + 1024;      // barriers + alignment slack

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ o, long long ldo,
                   float* __restrict__ lse, int H, int Tq, int Tk, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_pad = (Tk + 15) & ~15;            // UMMA N of the first chain / K of the second (multiple of 16)
  const int k_boxes = (n_pad + 127) / 128;      // 128-key K boxes
  const int v_boxes = (n_pad + 63) / 64;        // 64-key V boxes

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);  // one arrive per softmax warp
      mbar_init(bar_o, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- loads: Q (128 x 64) and K (k_boxes x 128 x 64) on one barrier, V (v_boxes x 64 x 64) on another
      mbar_expect_tx(bar_qk, kBytesQ + k_boxes * kBytesKBox);
      tma_load_3d(&tmQ, bar_qk, sQ, h * kHd, q0, b);
      for (int j = 0; j < k_boxes; ++j) tma_load_3d(&tmK, bar_qk, sK + j * kBytesKBox, h * kHd, j * 128, b);
      mbar_expect_tx(bar_v, v_boxes * kBytesVBox);
      for (int j = 0; j < v_boxes; ++j) tma_load_3d(&tmV, bar_v, sV + j * kBytesVBox, h * kHd, j * 64, b);

      // ---- S = Q . K^T : both operands K-major (head_dim contiguous), 4 UMMA-K steps of 16
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      {
        const uint32_t idesc = umma_idesc_bf16(kQ, n_pad, false, false);
        const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
#pragma unroll
        for (int ks = 0; ks < kHd / 16; ++ks)
          umma_bf16(tmem_base, umma_smem_desc(aq + ks * 32, 16, 1024), umma_smem_desc(ak + ks * 32, 16, 1024), idesc,
                    ks > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      // ---- O = P . V : A = P (K-major atoms of 64 keys), B = V (MN-major: head_dim contiguous, keys strided)
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after();
      {
        const uint32_t idesc = umma_idesc_bf16(kQ, kHd, false, true);
        const uint32_t ap = smem_u32(sP), av = smem_u32(sV);
        const int ksteps = n_pad / 16;
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t da = umma_smem_desc(ap + (kk >> 2) * kBytesPAtom + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_smem_desc(av + (kk >> 2) * kBytesVBox + (kk & 3) * (16 * 128), 64 * 128, 1024);
          umma_bf16(tmem_base + kOCol, da, db, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
    }
  } else {
    // ================================ softmax / epilogue warps ================================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch (hardware rule)
    const int row = quarter * 32 + lane;          // query row inside the tile == TMEM lane
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int chunks = (n_pad + 31) / 32;

    mbar_wait(bar_s, 0);
    tc_fence_after();
    float m2 = -INFINITY;  // running max of s * scale_log2
    for (int c = 0; c < chunks; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < Tk) m2 = fmaxf(m2, __uint_as_float(r[j]) * scale_log2);
    }
    float l = 0.f;
    for (int c = 0; c < chunks; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c * 32, r);
      tmem_ld_wait();
      float p[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        p[j] = (c * 32 + j < Tk) ? exp2f(fmaf(__uint_as_float(r[j]), scale_log2, -m2)) : 0.f;
        l += p[j];
      }
      // 32 keys = four 16-byte chunks of this row in the K-major, 128B-swizzled atom of 64 keys
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int key8 = c * 4 + g;               // index of the 8-key group along the row
        __nv_bfloat162 v0 = __floats2bfloat162_rn(p[8 * g + 0], p[8 * g + 1]);
        __nv_bfloat162 v1 = __floats2bfloat162_rn(p[8 * g + 2], p[8 * g + 3]);
        __nv_bfloat162 v2 = __floats2bfloat162_rn(p[8 * g + 4], p[8 * g + 5]);
        __nv_bfloat162 v3 = __floats2bfloat162_rn(p[8 * g + 6], p[8 * g + 7]);
        uint4 w;
        w.x = *reinterpret_cast<uint32_t*>(&v0);
        w.y = *reinterpret_cast<uint32_t*>(&v1);
        w.z = *reinterpret_cast<uint32_t*>(&v2);
        w.w = *reinterpret_cast<uint32_t*>(&v3);
        uint8_t* dst = sP + (key8 >> 3) * kBytesPAtom + row * 128 + (((key8 & 7) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = w;
      }
    }
    fence_proxy_async_smem();  // the UMMA reads P through the async proxy
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int qrow = q0 + row;
    const float inv = 1.f / l;
    __nv_bfloat16* dst = o + (static_cast<long long>(b) * Tq + qrow) * ldo + h * kHd;
#pragma unroll
    for (int c = 0; c < kHd / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(trow + kOCol + c * 32, r);
      tmem_ld_wait();
      if (qrow < Tq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __nv_bfloat162 v0 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 0]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
          __nv_bfloat162 v1 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
          __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
          __nv_bfloat162 v3 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
          uint4 w;
          w.x = *reinterpret_cast<uint32_t*>(&v0);
          w.y = *reinterpret_cast<uint32_t*>(&v1);
          w.z = *reinterpret_cast<uint32_t*>(&v2);
          w.w = *reinterpret_cast<uint32_t*>(&v3);
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = w;
        }
      }
    }
    if (qrow < Tq) lse[(static_cast<long long>(b) * H + h) * Tq + qrow] = m2 + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// bf16 [batch][rows][cols] view with a 64-column box (same encoding as the GEMM operand maps; kept local to this
// experimental translation unit so the validated GEMM file stays untouched -- fold into one helper when this lands).
static int make_map(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long batch, long long ld,
                    int box_rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return md_set_error(MD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(rows * ld) * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return md_set_error(MD_ERR_CUDA, "md_attn_fwd_tc: cuTensorMapEncodeTiled failed");
  return 0;
}

}  // namespace attn_tc0
}  // namespace md

extern "C" int md_attn_fwd_tc_v0(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                              int64_t ldo, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t hd,
                              void* stream) {
  using namespace md;
  using namespace md::attn_tc0;
  if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0) return B == 0 ? 0 : md_set_error(MD_ERR_INVALID, "md_attn_fwd_tc: bad sizes");
  if (hd != kHd || Tk > kMaxKeys || H > 65535 || B > 65535)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_attn_fwd_tc: needs head_dim 64 and Tk <= 256");
  if (!q || !k || !v || !o || !lse) return md_set_error(MD_ERR_INVALID, "md_attn_fwd_tc: null pointer");
  const uintptr_t align = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                          reinterpret_cast<uintptr_t>(o);
  if ((align & 15) != 0 || ((ldq | ldk | ldv | ldo) % 8) != 0)
    return md_set_error(MD_ERR_INVALID, "md_attn_fwd_tc: operands must be 16-byte aligned with pitches % 8 == 0");
  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_map(&tmQ, q, H * hd, Tq, B, ldq, kQ)) return rc;
  if (int rc = make_map(&tmK, k, H * hd, Tk, B, ldk, 128)) return rc;
  if (int rc = make_map(&tmV, v, H * hd, Tk, B, ldv, 64)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid(static_cast<unsigned>((Tq + kQ - 1) / kQ), static_cast<unsigned>(H), static_cast<unsigned>(B));
  const float sl2 = 1.4426950408889634f / sqrtf(static_cast<float>(hd));
  attn_fwd_tc_kernel<<<grid, kThreads, kSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      tmQ, tmK, tmV, reinterpret_cast<__nv_bfloat16*>(o), ldo, lse, static_cast<int>(H), static_cast<int>(Tq),
      static_cast<int>(Tk), sl2);
  return check_launch("md_attn_fwd_tc");
}

namespace md {
// ------------------------------------------------------------------------------------------------ backward
// One CTA per (sample, head), Tk <= 256 keys resident, 128-query blocks streamed.  TMEM: S-half [0,128), dP-half
// [128,256) (dQ reuses [0,64) once both halves are consumed), dK [256,384) and dV [384,512) as two 128-key tiles of 64
// columns each, accumulated across query blocks.  P and dS are written once per (block, half) as bf16 [query][key] in
// 128B-swizzled atoms of 64 keys: that buffer is the K-major A operand of dQ = dS . K and -- the same bytes read with
// the MN-major descriptor -- the A operand of dV = P^T . dO and dK = dS^T . Q.  delta = rowsum(dO * O) is taken from
// global memory by the thread that owns the row.  dQ / dK carry the 1/sqrt(hd) factor in their epilogues.
namespace attn_tc0 {

constexpr int kBwdOffdO = kBytesQ;
constexpr int kBwdOffK = kBwdOffdO + kBytesQ;
constexpr int kBwdOffV = kBwdOffK + 2 * kBytesKBox;
constexpr int kBwdOffP = kBwdOffV + 2 * kBytesKBox;
constexpr int kBwdOffdS = kBwdOffP + 4 * kBytesPAtom;
constexpr int kBwdOffBar = kBwdOffdS + 4 * kBytesPAtom;
constexpr int kBwdSmemBytes = kBwdOffBar + 128 + 1024;
static_assert(kBwdSmemBytes <= 232448, "backward staging exceeds the 227 KB shared-memory limit");

__device__ __forceinline__ void store_row64(__nv_bfloat16* dst, uint32_t taddr, float mul, bool ok) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c * 32, r);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __nv_bfloat162 v0 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 0]) * mul, __uint_as_float(r[8 * g + 1]) * mul);
        __nv_bfloat162 v1 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 2]) * mul, __uint_as_float(r[8 * g + 3]) * mul);
        __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 4]) * mul, __uint_as_float(r[8 * g + 5]) * mul);
        __nv_bfloat162 v3 = __floats2bfloat162_rn(__uint_as_float(r[8 * g + 6]) * mul, __uint_as_float(r[8 * g + 7]) * mul);
        uint4 w;
        w.x = *reinterpret_cast<uint32_t*>(&v0);
        w.y = *reinterpret_cast<uint32_t*>(&v1);
        w.z = *reinterpret_cast<uint32_t*>(&v2);
        w.w = *reinterpret_cast<uint32_t*>(&v3);
        *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = w;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __nv_bfloat16* __restrict__ dout, long long lddo, const __nv_bfloat16* __restrict__ o,
                   long long ldo, const float* __restrict__ lse, __nv_bfloat16* __restrict__ dq, long long lddq,
                   __nv_bfloat16* __restrict__ dk, long long lddk, __nv_bfloat16* __restrict__ dv, long long lddv, int H,
                   int Tq, int Tk, float scale, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + kBwdOffdO;
  uint8_t* sK = smem + kBwdOffK;
  uint8_t* sV = smem + kBwdOffV;
  uint8_t* sP = smem + kBwdOffP;
  uint8_t* sdS = smem + kBwdOffdS;
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kBwdOffBar);
  uint64_t* bar_q = bar_kv + 1;    // Q / dO block landed              (phase = block parity)
  uint64_t* bar_sdp = bar_kv + 2;  // S-half and dP-half in TMEM        (phase = (block * halves + half) parity)
  uint64_t* bar_pds = bar_kv + 3;  // P-half / dS-half written (4 warps)
  uint64_t* bar_dq = bar_kv + 4;   // dQ block in TMEM, all earlier MMAs retired
  uint64_t* bar_dqr = bar_kv + 5;  // dQ block read back (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x;
  const int b = blockIdx.y;
  const int n_pad = (Tk + 15) & ~15;
  const int halves = (n_pad + 127) / 128;
  const int q_blocks = (Tq + kQ - 1) / kQ;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmdO);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_init(bar_kv, 1);
      mbar_init(bar_q, 1);
      mbar_init(bar_sdp, 1);
      mbar_init(bar_pds, 4);
      mbar_init(bar_dq, 1);
      mbar_init(bar_dqr, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColdP = 128, kColdQ = 0, kColdK = 256, kColdV = 384;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP),
                     adS = smem_u32(sdS);
      mbar_expect_tx(bar_kv, 2 * halves * kBytesKBox);
      for (int j = 0; j < halves; ++j) {
        tma_load_3d(&tmK, bar_kv, sK + j * kBytesKBox, h * kHd, j * 128, b);
        tma_load_3d(&tmV, bar_kv, sV + j * kBytesKBox, h * kHd, j * 128, b);
      }
      const uint32_t idesc_kk = 0;  // placeholder to keep the three descriptor kinds next to each other
      (void)idesc_kk;
      const uint32_t idesc_mn = umma_idesc_bf16(kQ, kHd, true, true);    // dV, dK: A and B MN-major
      const uint32_t idesc_dq = umma_idesc_bf16(kQ, kHd, false, true);   // dQ: A K-major, B MN-major
      uint32_t it = 0;  // (block, half) counter for the single-use phases of bar_sdp / bar_pds
      for (int qb = 0; qb < q_blocks; ++qb) {
        if (qb > 0) {
          mbar_wait(bar_dq, (qb - 1) & 1);   // every MMA that read sQ / sdO / sP / sdS has retired
          mbar_wait(bar_dqr, (qb - 1) & 1);  // the dQ columns have been read back: TMEM [0, 256) is free again
        }
        mbar_expect_tx(bar_q, 2 * kBytesQ);
        tma_load_3d(&tmQ, bar_q, sQ, h * kHd, qb * kQ, b);
        tma_load_3d(&tmdO, bar_q, sdO, h * kHd, qb * kQ, b);
        mbar_wait(bar_q, qb & 1);
        if (qb == 0) mbar_wait(bar_kv, 0);
        tc_fence_after();
        for (int hf = 0; hf < halves; ++hf, ++it) {
          const int n_half = min(128, n_pad - hf * 128);
          const uint32_t idesc_s = umma_idesc_bf16(kQ, n_half, false, false);
          // S-half = Q . K_half^T and dP-half = dO . V_half^T (all K-major)
#pragma unroll
          for (int ks = 0; ks < kHd / 16; ++ks)
            umma_bf16(tmem_base + kColS, umma_smem_desc(aQ + ks * 32, 16, 1024),
                      umma_smem_desc(aK + hf * kBytesKBox + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < kHd / 16; ++ks)
            umma_bf16(tmem_base + kColdP, umma_smem_desc(adO + ks * 32, 16, 1024),
                      umma_smem_desc(aV + hf * kBytesKBox + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
          umma_commit(bar_sdp);
          mbar_wait(bar_pds, it & 1);  // P-half / dS-half are in shared memory, the S / dP columns have been read
          tc_fence_after();
          // dV_half += P_half^T . dO   and   dK_half += dS_half^T . Q   (reduction over the block's 128 queries)
          for (int kk = 0; kk < kQ / 16; ++kk) {
            const uint64_t db_do = umma_smem_desc(adO + kk * (16 * 128), 64 * 128, 1024);
            const uint64_t db_q = umma_smem_desc(aQ + kk * (16 * 128), 64 * 128, 1024);
            const uint64_t da_p = umma_smem_desc(aP + hf * 2 * kBytesPAtom + kk * (16 * 128), kBytesPAtom, 1024);
            const uint64_t da_ds = umma_smem_desc(adS + hf * 2 * kBytesPAtom + kk * (16 * 128), kBytesPAtom, 1024);
            const uint32_t acc = (qb > 0 || kk > 0) ? 1u : 0u;
            umma_bf16(tmem_base + kColdV + hf * kHd, da_p, db_do, idesc_mn, acc);
            umma_bf16(tmem_base + kColdK + hf * kHd, da_ds, db_q, idesc_mn, acc);
          }
        }
        // dQ block = dS . K over all resident keys (A K-major atoms, B = K tile read MN-major)
        for (int kk = 0; kk < n_pad / 16; ++kk)
          umma_bf16(tmem_base + kColdQ, umma_smem_desc(adS + (kk >> 2) * kBytesPAtom + (kk & 3) * 32, 16, 1024),
                    umma_smem_desc(aK + kk * (16 * 128), 64 * 128, 1024), idesc_dq, kk > 0 ? 1u : 0u);
        umma_commit(bar_dq);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t it = 0;
    for (int qb = 0; qb < q_blocks; ++qb) {
      const int qrow = qb * kQ + row;
      const bool q_ok = qrow < Tq;
      // lse (log2 domain) and delta = sum_d dO * O of this thread's query row
      float lrow = INFINITY, delta = 0.f;  // +inf -> P = 0 for padded queries
      if (q_ok) {
        lrow = lse[(static_cast<long long>(b) * H + h) * Tq + qrow];
        const uint4* po = reinterpret_cast<const uint4*>(o + (static_cast<long long>(b) * Tq + qrow) * ldo + h * kHd);
        const uint4* pd = reinterpret_cast<const uint4*>(dout + (static_cast<long long>(b) * Tq + qrow) * lddo + h * kHd);
#pragma unroll
        for (int j = 0; j < kHd / 8; ++j) {
          const uint4 a = po[j], c = pd[j];
          const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fa = __bfloat1622float2(ah[e]), fc = __bfloat1622float2(ch[e]);
            delta = fmaf(fa.x, fc.x, delta);
            delta = fmaf(fa.y, fc.y, delta);
          }
        }
      }
      for (int hf = 0; hf < halves; ++hf, ++it) {
        mbar_wait(bar_sdp, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t rs[32], rp[32];
          tmem_ld_32x32(trow + 0 + c * 32, rs);
          tmem_ld_32x32(trow + 128 + c * 32, rp);
          tmem_ld_wait();
          float p[32], ds[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = hf * 128 + c * 32 + j;
            p[j] = key < Tk ? exp2f(fmaf(__uint_as_float(rs[j]), scale_log2, -lrow)) : 0.f;
            ds[j] = key < Tk ? p[j] * (__uint_as_float(rp[j]) - delta) : 0.f;  // stale TMEM beyond the MMA's N may be NaN
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int key8 = hf * 16 + c * 4 + g;
            const int off = (key8 >> 3) * kBytesPAtom + row * 128 + (((key8 & 7) ^ (row & 7)) << 4);
            __nv_bfloat162 a0 = __floats2bfloat162_rn(p[8 * g + 0], p[8 * g + 1]);
            __nv_bfloat162 a1 = __floats2bfloat162_rn(p[8 * g + 2], p[8 * g + 3]);
            __nv_bfloat162 a2 = __floats2bfloat162_rn(p[8 * g + 4], p[8 * g + 5]);
            __nv_bfloat162 a3 = __floats2bfloat162_rn(p[8 * g + 6], p[8 * g + 7]);
            uint4 w;
            w.x = *reinterpret_cast<uint32_t*>(&a0);
            w.y = *reinterpret_cast<uint32_t*>(&a1);
            w.z = *reinterpret_cast<uint32_t*>(&a2);
            w.w = *reinterpret_cast<uint32_t*>(&a3);
            *reinterpret_cast<uint4*>(sP + off) = w;
            a0 = __floats2bfloat162_rn(ds[8 * g + 0], ds[8 * g + 1]);
            a1 = __floats2bfloat162_rn(ds[8 * g + 2], ds[8 * g + 3]);
            a2 = __floats2bfloat162_rn(ds[8 * g + 4], ds[8 * g + 5]);
            a3 = __floats2bfloat162_rn(ds[8 * g + 6], ds[8 * g + 7]);
            w.x = *reinterpret_cast<uint32_t*>(&a0);
            w.y = *reinterpret_cast<uint32_t*>(&a1);
            w.z = *reinterpret_cast<uint32_t*>(&a2);
            w.w = *reinterpret_cast<uint32_t*>(&a3);
            *reinterpret_cast<uint4*>(sdS + off) = w;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pds);
      }
      // dQ block
      mbar_wait(bar_dq, qb & 1);
      tc_fence_after();
      store_row64(dq + (static_cast<long long>(b) * Tq + qrow) * lddq + h * kHd, trow + 0, scale, q_ok);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dqr);
    }
    // dK / dV: the last bar_dq phase also covered their MMAs.  TMEM lane = key inside the 128-key tile.
    for (int hf = 0; hf < halves; ++hf) {
      const int key = hf * 128 + row;
      const bool k_ok = key < Tk;
      store_row64(dk + (static_cast<long long>(b) * Tk + key) * lddk + h * kHd, trow + 256 + hf * kHd, scale, k_ok);
      store_row64(dv + (static_cast<long long>(b) * Tk + key) * lddv + h * kHd, trow + 384 + hf * kHd, 1.0f, k_ok);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace attn_tc0
}  // namespace md

extern "C" int md_attn_bwd_tc_v0(const void* dout, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                              const void* v, int64_t ldv, const void* o, int64_t ldo, const float* lse, void* dq,
                              int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t B, int64_t H,
                              int64_t Tq, int64_t Tk, int64_t hd, void* stream) {
  using namespace md;
  using namespace md::attn_tc0;
  if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0) return B == 0 ? 0 : md_set_error(MD_ERR_INVALID, "md_attn_bwd_tc: bad sizes");
  if (hd != kHd || Tk > kMaxKeys || H > 65535 || B > 65535)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_attn_bwd_tc: needs head_dim 64 and Tk <= 256");
  if (!dout || !q || !k || !v || !o || !lse || !dq || !dk || !dv)
    return md_set_error(MD_ERR_INVALID, "md_attn_bwd_tc: null pointer");
  const uintptr_t align = reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                          reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dq) |
                          reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv);
  if ((align & 15) != 0 || ((lddo | ldq | ldk | ldv | ldo | lddq | lddk | lddv) % 8) != 0)
    return md_set_error(MD_ERR_INVALID, "md_attn_bwd_tc: operands must be 16-byte aligned with pitches % 8 == 0");
  CUtensorMap tmQ, tmdO, tmK, tmV;
  if (int rc = make_map(&tmQ, q, H * hd, Tq, B, ldq, kQ)) return rc;
  if (int rc = make_map(&tmdO, dout, H * hd, Tq, B, lddo, kQ)) return rc;
  if (int rc = make_map(&tmK, k, H * hd, Tk, B, ldk, 128)) return rc;
  if (int rc = make_map(&tmV, v, H * hd, Tk, B, ldv, 128)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes);
    if (e != cudaSuccess) return md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid(static_cast<unsigned>(H), static_cast<unsigned>(B));
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  attn_bwd_tc_kernel<<<grid, kThreads, kBwdSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      tmQ, tmdO, tmK, tmV, reinterpret_cast<const __nv_bfloat16*>(dout), lddo, reinterpret_cast<const __nv_bfloat16*>(o), ldo,
      lse, reinterpret_cast<__nv_bfloat16*>(dq), lddq, reinterpret_cast<__nv_bfloat16*>(dk), lddk,
      reinterpret_cast<__nv_bfloat16*>(dv), lddv, static_cast<int>(H), static_cast<int>(Tq), static_cast<int>(Tk), scale,
      1.4426950408889634f * scale);
  return check_launch("md_attn_bwd_tc");
}
