// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue with fused element-wise tails.
//
// Replaces, on the MicroDiT training path, every nn.Linear / einsum contraction the reference sends
// to cuBLAS: qkv/proj (reference micro_diffusion/models/utils.py:172-173), cross-attention q/kv/proj
// (utils.py:109-111), SwiGLU w1/w2/w3 (dit.py:84-89), the expert einsums (dit.py:135-137), the
// adaLN / stem / mixer-map / final linears, and all their dgrad / wgrad counterparts.
//
// Two operand layouts (template kMN):
//   kMN=false  "NT":  C[M,N] = sum_k A[M,k] * B[N,k]     A,B row-major with k contiguous (K-major)
//   kMN=true   "TN":  C[P,Q] = sum_r A[r,P] * B[r,Q]     A,B row-major with the reduction index r
//                     strided (MN-major UMMA operands) -- the weight-gradient contraction.
// Roles: warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator, warps 4.. = epilogue (warp w reads TMEM
// lanes 32*(w%4)..+31).  kEpiW = 4 epilogue warps for the store-only tails (8 measured equal or slightly slower once
// the epilogue stopped spilling its accumulator chunk to local memory); kEpiW = 8 (a second warp per lane quarter,
// half of the columns each) for the math-heavy tails -- GELU + dual store, activation gradient -- where one warp per
// scheduler cannot hide the MUFU / FMA latency of 256 activations per row and tile.
//
// kCtas = 2 is the Blackwell CTA-pair mode: two CTAs of a cluster (adjacent SMs) run ONE tcgen05.mma.cta_group::2
// with M = 256 (128 rows of D in each CTA's TMEM); every CTA stages its own 128 rows of A and only HALF of the B
// tile, so shared-memory traffic per flop drops by a third and the pair reads B once from L2.  Measured on B200
// the 1-CTA 128x256 mainloop is shared-memory-bandwidth bound (TMA writes + UMMA reads ~192 B/clk vs 128 B/clk).
// Protocol: both CTAs issue their TMA loads with .cta_group::2 so the bytes are credited to the LEADER's (even
// CTA's) full barrier; only the leader issues MMAs; tcgen05.commit multicasts the stage-free / accumulator-ready
// arrivals to both CTAs; both epilogues arrive on the leader's accumulator-free barrier.
#include "gemm_tcgen05.cuh"

#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "ptx.cuh"
#include "tensormap.cuh"

namespace md {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;

template <int BLOCK_N, int kCtas, int kEpiW>
struct GemmCfg {
  static_assert(kEpiW == 4 || kEpiW == 8, "epilogue warps: 4 or 8");
  static constexpr int kColSplit = kEpiW / 4;
  static constexpr int kThreads = 128 + 32 * kEpiW;  // 4 control warps + the epilogue warps
  static constexpr int kStageBytesA = kBlockM * kBlockK * 2;
  static constexpr int kStageBytesB = (BLOCK_N / kCtas) * kBlockK * 2;  // a CTA of a pair stages half of B
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (kStageBytes > 40 * 1024) ? 4 : (kStageBytes > 30 * 1024 ? 6 : 7);
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = kAccStages * BLOCK_N;  // 256 or 512 (power of two)
  // epilogue staging for the TMA store: per warp 2 buffers x (32 rows x 32 bf16 = 2 KB, 64B-swizzled)
  static constexpr int kStoreBufBytes = 32 * 32 * 2;
  static constexpr int kStoreBytes = kEpiW * 2 * kStoreBufBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Cheap activations for the epilogue (it shares 4 issue ports with nothing else but is on the critical path of
// short-K tiles).  erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below bf16 resolution).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);  // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float u = 0.7978845608028654f * fmaf(0.044715f * x, x * x, x);
  const float e = __expf(2.0f * u);                    // tanh(u) = 1 - 2/(e^{2u}+1)
  const float th = 1.0f - __fdividef(2.0f, e + 1.0f);
  return 0.5f * x * (1.0f + th);
}

// d gelu / dx for the activation-gradient tail; shares exp(-x^2/2) between the erf and the density term.
__device__ __forceinline__ float gelu_erf_grad_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float ez = __expf(-z * z);                     // exp(-x^2 / 2)
  const float e = 1.0f - poly * t * ez;                // erf(|x| / sqrt2)
  const float cdf = 0.5f * (1.0f + copysignf(e, x));
  return fmaf(x * 0.3989422804014327f, ez, cdf);
}
__device__ __forceinline__ float gelu_tanh_grad_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * fmaf(k1 * x, x * x, x);
  const float e = __expf(2.0f * u);
  const float th = 1.0f - __fdividef(2.0f, e + 1.0f);
  return 0.5f * (1.0f + th) + 0.5f * x * (1.0f - th * th) * k0 * fmaf(3.0f * k1 * x, x, 1.0f);
}

// The activation kind is a compile-time parameter of the chunk loops: with a runtime `act ? tanh : erf` per element the
// compiler evaluated BOTH activations and selected (the first fused epilogues spent twice the issue slots they needed).
template <bool kTanh> __device__ __forceinline__ float gelu_fast(float x) {
  if constexpr (kTanh) return gelu_tanh_fast(x);
  else return gelu_erf_fast(x);
}
template <bool kTanh> __device__ __forceinline__ float gelu_grad_fast(float x) {
  if constexpr (kTanh) return gelu_tanh_grad_fast(x);
  else return gelu_erf_grad_fast(x);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b);
__device__ __forceinline__ uint4 pack_bf16x8(float a0, float a1, float a2, float a3, float a4, float a5, float a6,
                                             float a7);
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint4 pack_bf16x8(float a0, float a1, float a2, float a3, float a4, float a5, float a6,
                                             float a7) {
  return make_uint4(pack_bf16(a0, a1), pack_bf16(a2, a3), pack_bf16(a4, a5), pack_bf16(a6, a7));
}

// Tile order inside one (batch, split) slice: bands of kBand n-blocks, n fastest inside a band, then m, then the
// next band.  Co-resident CTAs share A rows (as before), and a band's B tiles (<= 4 MB) stay in L2 while the sweep
// over m reuses them -- without this a wide-N GEMM (the stacked K/V projection, N = 57k) re-streams B from HBM for
// every row block.
constexpr int kBand = 8;
__device__ __forceinline__ void decode_tile(int r, int m_blocks, int n_blocks, int& mb, int& nb) {
  const int per_band = kBand * m_blocks;
  const int band = r / per_band;
  const int bw = min(kBand, n_blocks - band * kBand);
  const int rr = r - band * per_band;
  mb = rr / bw;
  nb = band * kBand + rr % bw;
}

template <int BLOCK_N, bool kMN, int kCtas, int kEpiW>
__global__ void __launch_bounds__(128 + 32 * kEpiW, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmDev p) {
  using Cfg = GemmCfg<BLOCK_N, kCtas, kEpiW>;
  constexpr int kColSplit = Cfg::kColSplit;
  // the 8-warp instantiations serve the activation tails only (host dispatch): compiling the residual / atomic / fp32
  // tails out keeps their prefetch registers from pushing the 384-thread kernel over its 168-register budget
  constexpr bool kMath = kEpiW == 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* store_stage = smem + Cfg::kStages * Cfg::kStageBytes;  // 1024-aligned (stage sizes are multiples of 1 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + Cfg::kStoreBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + Cfg::kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + Cfg::kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store == 1) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < Cfg::kAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiW * kCtas);  // one arrive per epilogue warp (of both CTAs in pair mode)
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (kCtas == 2) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCtas == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int cta_rank = (kCtas == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const bool is_leader = cta_rank == 0;

  // Tile enumeration: n fastest so that CTAs resident together share the same A rows through L2.
  // In pair mode a "tile" is 256 x BLOCK_N: m_blocks counts 256-row units and this CTA owns rows
  // (mb * kCtas + cta_rank) * 128 .. +127 of it (an odd tail is an all-out-of-bounds 128-row block: TMA zero-fills,
  // the epilogue skips it, the protocol stays symmetric).
  const int m_blocks = (p.M + kBlockM * kCtas - 1) / (kBlockM * kCtas);
  const int n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (p.K + kBlockK - 1) / kBlockK;
  const int kb_per_split = (kb_total + p.splits - 1) / p.splits;
  const long long tiles = 1LL * p.batch * p.splits * m_blocks * n_blocks;
  const long long tile0 = blockIdx.x / kCtas, tile_step = gridDim.x / kCtas;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = tile0; tile < tiles; tile += tile_step) {
        long long t = tile;
        const long long mn = 1LL * n_blocks * m_blocks;
        int mb, nb;
        decode_tile(static_cast<int>(t % mn), m_blocks, n_blocks, mb, nb);
        t /= mn;
        const int sp = t % p.splits; t /= p.splits;
        const int bz = static_cast<int>(t);
        const int kb0 = sp * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kStageBytesA;
          const int m0 = (mb * kCtas + cta_rank) * kBlockM;                 // this CTA's 128 rows of A / D
          const int n0 = nb * BLOCK_N + cta_rank * (BLOCK_N / kCtas);       // this CTA's share of the B tile
          if (is_leader) mbar_expect_tx(&full_bar[s], Cfg::kStageBytes * kCtas);
          auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1) {
            if constexpr (kCtas == 2) tma_load_3d_2sm(tm, &full_bar[s], dst, c0, c1, bz);
            else tma_load_3d(tm, &full_bar[s], dst, c0, c1, bz);
          };
          if constexpr (!kMN) {
            load(&tmA, sa, kb * kBlockK, m0);
            load(&tmB, sb, kb * kBlockK, n0);
          } else {
            // boxes of 64 (MN, contiguous) x 64 (reduction rows); one box per 64 MN elements
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j) load(&tmA, sa + j * (kBlockK * 128), m0 + j * 64, kb * kBlockK);
#pragma unroll
            for (int j = 0; j < BLOCK_N / kCtas / 64; ++j)
              load(&tmB, sb + j * (kBlockK * 128), n0 + j * 64, kb * kBlockK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0 && is_leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * kCtas, BLOCK_N, kMN, kMN);
      // K-major: 8-row groups 1024 B apart (SBO), LBO unused (1).  MN-major: 64-element MN groups one
      // whole box apart (LBO = 64 rows * 128 B), 8-row reduction groups 1024 B apart (SBO).
      constexpr uint32_t lbo = kMN ? (kBlockK * 128) : 16;
      constexpr uint32_t sbo = 1024;
      constexpr uint32_t kstep = kMN ? (kUmmaK * 128) : (kUmmaK * 2);  // bytes per UMMA_K advance
      uint32_t it = 0;
      uint32_t acc_it = 0;
      for (long long tile = tile0; tile < tiles; tile += tile_step, ++acc_it) {
        long long t = tile / (1LL * n_blocks * m_blocks);
        const int sp = t % p.splits;
        const int kb0 = sp * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        const int as = acc_it % Cfg::kAccStages;
        const uint32_t aph = (acc_it / Cfg::kAccStages) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kStageBytesA;
          // one descriptor per operand and stage; the UMMA_K steps advance its 14-bit (address >> 4) field by plain adds
          // (rebuilding both descriptors for every MMA cost the issuing thread ~100 cycles per UTCHMMA: harmless under a
          // 128-cycle 256x256x16 pair MMA, but the limit for the 64-cycle 128-wide tiles)
          const uint64_t da0 = umma_smem_desc(sa, lbo, sbo);
          const uint64_t db0 = umma_smem_desc(sb, lbo, sbo);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = da0 + static_cast<uint64_t>(k * (kstep >> 4));
            const uint64_t db = db0 + static_cast<uint64_t>(k * (kstep >> 4));
            if constexpr (kCtas == 2) umma_bf16_2sm(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (kCtas == 2) umma_commit_2sm(&empty_bar[s], 3);
          else umma_commit(&empty_bar[s]);
        }
        // accumulator complete (signalled to both epilogues of a pair)
        if constexpr (kCtas == 2) umma_commit_2sm(&tfull_bar[as], 3);
        else umma_commit(&tfull_bar[as]);
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    // TMEM lane quarter q = warp % 4 (hardware rule); with 8 epilogue warps, column half = (warp - 4) / 4.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kChunksPerWarp = BLOCK_N / kColSplit / 32;  // 32-column chunks per warp
    uint8_t* my_stage = store_stage + (warp - 4) * (2 * Cfg::kStoreBufBytes);
    uint32_t store_it = 0;  // chunks this warp has handed to the TMA (selects the staging buffer)
    uint32_t acc_it = 0;
    for (long long tile = tile0; tile < tiles; tile += tile_step, ++acc_it) {
      long long t = tile;
      const long long mn = 1LL * n_blocks * m_blocks;
      int mb, nb;
      decode_tile(static_cast<int>(t % mn), m_blocks, n_blocks, mb, nb);
      t /= mn;
      const int sp = t % p.splits; t /= p.splits;
      const int bz = static_cast<int>(t);
      const int kb0 = sp * kb_per_split;
      const bool has_k = kb0 < kb_total;  // empty split (possible when splits does not divide)
      const int as = acc_it % Cfg::kAccStages;
      const uint32_t aph = (acc_it / Cfg::kAccStages) & 1;

      const int row = (mb * kCtas + cta_rank) * kBlockM + q * 32 + lane;
      const bool row_ok = row < p.M;
      // weight gradient of a 32-row-interleaved stack (the fused-SwiGLU w1 | w2 layout): tile row -> parameter row
      int row_out = row;
      if (p.row_interleave > 0) {
        const int blk = row >> 6, in = row & 63;
        row_out = (in < 32 ? 0 : p.row_interleave) + 32 * blk + (in & 31);
      }
      const long long crow = 1LL * bz * p.strideC + 1LL * row_out * p.ldc;
      const long long rrow = 1LL * bz * p.strideC + 1LL * (p.res_mod > 0 ? row % p.res_mod : row) * p.ldc;
      const float* gate_row = nullptr;
      if (p.gate != nullptr && row_ok) gate_row = p.gate + 1LL * (row / p.rows_per_gate) * p.ldgate;
      const int colbase = nb * BLOCK_N + half * (BLOCK_N / kColSplit);
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + half * (BLOCK_N / kColSplit);
      const bool want_res = !kMath && (p.epi == EPI_RESID_F32) && row_ok && has_k;

      // residual prefetch for chunk 0 is independent of the accumulator: issue it before waiting on the MMAs
      float4 resn[8];
      auto load_res = [&](int c) {
        const int col0 = colbase + c * 32;
        const float* res = p.res + rrow + col0;
        const bool vec = want_res && (col0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(res) & 15) == 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (vec) {
            resn[j] = *reinterpret_cast<const float4*>(res + 4 * j);
          } else {
            float tmp[4] = {0.f, 0.f, 0.f, 0.f};
            if (want_res)
              for (int e = 0; e < 4; ++e)
                if (col0 + 4 * j + e < p.N) tmp[e] = res[4 * j + e];
            resn[j] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
          }
        }
      };
      if constexpr (!kMath) { if (p.epi == EPI_RESID_F32) load_res(0); }
      // saved pre-activation tile of the activation-gradient tail (bf16, indexed like C): same software pipeline
      uint4 auxn[4] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
      auto load_aux = [&](int c) {
        const int col0 = colbase + c * 32;
        if (p.tma_store == 2) {
          // coalesced: lane -> (row (lane >> 2) + 8 i, 16-byte piece lane & 3) of the warp's 32 x 32 box, i.e. full 64-byte
          // row segments per four lanes; the box is transposed to one row per lane through shared memory when it is used
          const int row0 = (mb * kCtas + cta_rank) * kBlockM + q * 32;
          const int pc = lane & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = row0 + (lane >> 2) + 8 * i;
            const bool ok = r < p.M && col0 + pc * 8 < p.N;  // N % 8 == 0 in this mode: whole pieces
            auxn[i] = ok ? *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) +
                                                            1LL * bz * p.strideC + 1LL * r * p.ldc + col0 + pc * 8)
                         : make_uint4(0u, 0u, 0u, 0u);
          }
          return;
        }
        const __nv_bfloat16* ax = reinterpret_cast<const __nv_bfloat16*>(p.aux) + crow + col0;
        if (p.tma_store == 3) {  // one row per lane, two full 32-byte sectors per chunk (N % 16 == 0: whole pieces)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (row_ok && col0 + 16 * k < p.N) ld_global_256(ax + 16 * k, auxn[2 * k], auxn[2 * k + 1]);
            else auxn[2 * k] = auxn[2 * k + 1] = make_uint4(0u, 0u, 0u, 0u);
          }
          return;
        }
        const bool live = row_ok && has_k && col0 < p.N;
        const bool vec = live && (col0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(ax) & 15) == 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (vec) {
            auxn[j] = *reinterpret_cast<const uint4*>(ax + 8 * j);
          } else {
            uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
            if (live) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                if (col0 + 8 * j + e < p.N) {
                  const uint32_t b = static_cast<uint32_t>(__bfloat16_as_ushort(ax[8 * j + e])) << (16 * (e & 1));
                  if ((e >> 1) == 0) w0 |= b;
                  else if ((e >> 1) == 1) w1 |= b;
                  else if ((e >> 1) == 2) w2 |= b;
                  else w3 |= b;
                }
              }
            }
            auxn[j] = make_uint4(w0, w1, w2, w3);
          }
        }
      };
      if constexpr (kMath) { if (p.epi == EPI_ACT_GRAD) load_aux(0); }

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      bool swiglu_done = false;
      if constexpr (kMath) {
        if (p.epi == EPI_SWIGLU) {
          // SwiGLU (dit.py:88-89) in the epilogue of the stacked w1 | w2 GEMM.  The weight stack is interleaved in blocks
          // of 32 rows (w1 block j, w2 block j, ...), so chunk 2c holds u1 and chunk 2c + 1 the matching u2 columns of the
          // same lane: C = u (bf16, the interleaved layout backward reads again), C2 = silu(u1) * u2 (bf16, natural).
          swiglu_done = true;
#pragma unroll 1
          for (int cp = 0; cp < kChunksPerWarp / 2; ++cp) {
            uint32_t r1[32], r2[32];
            tmem_ld_32x32(tbase + (2 * cp) * 32, r1);
            tmem_ld_32x32(tbase + (2 * cp + 1) * 32, r2);
            tmem_ld_wait();
            const int col0 = colbase + 2 * cp * 32;
            if (row_ok && col0 < p.N) {
              __nv_bfloat16* du = reinterpret_cast<__nv_bfloat16*>(p.C) + crow + col0;
              __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(p.C2) + 1LL * bz * p.strideC2 + 1LL * row * p.ldc2 + (col0 >> 1);
              float a[32], b[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                a[j] = bf16_round(__uint_as_float(r1[j]) * p.alpha);
                b[j] = bf16_round(__uint_as_float(r2[j]) * p.alpha);
              }
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                st_global_256(du + 16 * k, pack_bf16x8(a[16 * k], a[16 * k + 1], a[16 * k + 2], a[16 * k + 3], a[16 * k + 4], a[16 * k + 5], a[16 * k + 6], a[16 * k + 7]),
                              pack_bf16x8(a[16 * k + 8], a[16 * k + 9], a[16 * k + 10], a[16 * k + 11], a[16 * k + 12], a[16 * k + 13], a[16 * k + 14], a[16 * k + 15]));
                st_global_256(du + 32 + 16 * k, pack_bf16x8(b[16 * k], b[16 * k + 1], b[16 * k + 2], b[16 * k + 3], b[16 * k + 4], b[16 * k + 5], b[16 * k + 6], b[16 * k + 7]),
                              pack_bf16x8(b[16 * k + 8], b[16 * k + 9], b[16 * k + 10], b[16 * k + 11], b[16 * k + 12], b[16 * k + 13], b[16 * k + 14], b[16 * k + 15]));
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) a[j] = a[j] * __fdividef(1.0f, 1.0f + __expf(-a[j])) * b[j];
#pragma unroll
              for (int k = 0; k < 2; ++k)
                st_global_256(dh + 16 * k, pack_bf16x8(a[16 * k], a[16 * k + 1], a[16 * k + 2], a[16 * k + 3], a[16 * k + 4], a[16 * k + 5], a[16 * k + 6], a[16 * k + 7]),
                              pack_bf16x8(a[16 * k + 8], a[16 * k + 9], a[16 * k + 10], a[16 * k + 11], a[16 * k + 12], a[16 * k + 13], a[16 * k + 14], a[16 * k + 15]));
            }
            __syncwarp();
          }
        } else if (p.epi == EPI_SWIGLU_GRAD) {
          // backward of the above inside the w3 dgrad GEMM: acc = d h (never stored); with the saved u1 | u2 of the same 32
          // columns (aux, interleaved layout, 128 contiguous bytes per lane and chunk) the chunk leaves as
          // d u1 = d h * u2 * silu'(u1), d u2 = d h * silu(u1) -- again 128 contiguous bytes of the interleaved d u.
          swiglu_done = true;
          uint4 ua[8];
          auto load_u = [&](int c) {
            const int col0 = colbase + c * 32;
            const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.aux) + crow + 2 * col0;
            if (row_ok && col0 < p.N) {
#pragma unroll
              for (int k = 0; k < 4; ++k) ld_global_256(src + 16 * k, ua[2 * k], ua[2 * k + 1]);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) ua[k] = make_uint4(0u, 0u, 0u, 0u);
            }
          };
          load_u(0);
#pragma unroll 1
          for (int c = 0; c < kChunksPerWarp; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tbase + c * 32, r);
            tmem_ld_wait();
            const int col0 = colbase + c * 32;
            const bool live = row_ok && col0 < p.N;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + crow + 2 * col0;
            uint4 o1[4], o2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 8 columns per k: u1 words ua[k], u2 words ua[4 + k]
              uint32_t w1[4] = {ua[k].x, ua[k].y, ua[k].z, ua[k].w};
              uint32_t w2[4] = {ua[4 + k].x, ua[4 + k].y, ua[4 + k].z, ua[4 + k].w};
              float d1[8], d2[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float u1 = __uint_as_float((e & 1) ? (w1[e >> 1] & 0xffff0000u) : (w1[e >> 1] << 16));
                const float u2 = __uint_as_float((e & 1) ? (w2[e >> 1] & 0xffff0000u) : (w2[e >> 1] << 16));
                const float d = __uint_as_float(r[8 * k + e]) * p.alpha;
                const float sg = __fdividef(1.0f, 1.0f + __expf(-u1));
                d1[e] = d * u2 * (sg * fmaf(u1, 1.0f - sg, 1.0f));
                d2[e] = d * u1 * sg;
              }
              o1[k] = pack_bf16x8(d1[0], d1[1], d1[2], d1[3], d1[4], d1[5], d1[6], d1[7]);
              o2[k] = pack_bf16x8(d2[0], d2[1], d2[2], d2[3], d2[4], d2[5], d2[6], d2[7]);
            }
            if (c + 1 < kChunksPerWarp) load_u(c + 1);   // flies under the next chunk's TMEM wait and the other warp's math
            if (live) {
              st_global_256(dst, o1[0], o1[1]);
              st_global_256(dst + 16, o1[2], o1[3]);
              st_global_256(dst + 32, o2[0], o2[1]);
              st_global_256(dst + 48, o2[2], o2[3]);
            }
            __syncwarp();
          }
        }
      }
      uint32_t rnext[32];
      if (p.debug != 2 && !swiglu_done) tmem_ld_32x32(tbase, rnext);

#pragma unroll 1
      for (int c = 0; c < ((p.debug == 2 || swiglu_done) ? 0 : kChunksPerWarp); ++c) {
        tmem_ld_wait();
        float v[32];
        float4 resv[8];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rnext[j]) * p.alpha;
#pragma unroll
        for (int j = 0; j < 8; ++j) resv[j] = resn[j];
        uint4 auxv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) auxv[j] = auxn[j];
        if (c + 1 < kChunksPerWarp) {  // software pipeline: next chunk's TMEM + residual loads fly during this chunk
          tmem_ld_32x32(tbase + (c + 1) * 32, rnext);
          if constexpr (!kMath) { if (p.epi == EPI_RESID_F32) load_res(c + 1); }
          if constexpr (kMath) { if (p.epi == EPI_ACT_GRAD) load_aux(c + 1); }
        }
        const int col0 = colbase + c * 32;
        if (kMath && p.epi == EPI_ACT_GRAD) {  // C = acc * act'(pre): the dgrad GEMM hands the pre-activation gradient on directly
          if (p.tma_store == 2) {
            uint8_t* ab = my_stage + Cfg::kStoreBufBytes;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = (lane >> 2) + 8 * i;
              *reinterpret_cast<uint4*>(ab + r * 64 + (((lane & 3) ^ ((r >> 1) & 3)) << 4)) = auxv[i];
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              auxv[j] = *reinterpret_cast<const uint4*>(ab + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
            __syncwarp();
          }
          auto apply_grad = [&](auto tanh_tag) {
            constexpr bool kTanh = decltype(tanh_tag)::value;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t w = e == 0 ? auxv[j].x : e == 1 ? auxv[j].y : e == 2 ? auxv[j].z : auxv[j].w;
                const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
                v[8 * j + 2 * e] *= gelu_grad_fast<kTanh>(x0);
                v[8 * j + 2 * e + 1] *= gelu_grad_fast<kTanh>(x1);
              }
            }
          };
          if (p.act) apply_grad(std::true_type{});
          else apply_grad(std::false_type{});
        }
        if (p.debug == 1) {
          if (v[0] == 123.456f && v[31] == -654.321f) reinterpret_cast<float*>(p.C)[0] = v[5];  // keep the loads alive
        } else if (p.tma_store == 3) {
          // bf16 store straight from the accumulator registers: every lane owns one output row and writes its 64 bytes of the
          // chunk as two 256-bit stores = two full 32-byte sectors, no shared-memory staging at all.  (The staged variants
          // -- TMA store or coalesced st.global -- cost 4-8 KB of shared-memory traffic per chunk and warp, in a mainloop that
          // is itself shared-memory-bandwidth bound.)
          if (row_ok && col0 < p.N) {
            if (p.bias != nullptr) {
              const int ncols = min(32, p.N - col0);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += p.bias[1LL * bz * p.strideBias + col0 + j];
            }
            __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(p.C) + crow + col0;
            auto put = [&](__nv_bfloat16* dst, const float (&f)[32]) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                if (col0 + 16 * k < p.N)
                  st_global_256(dst + 16 * k,
                                pack_bf16x8(f[16 * k], f[16 * k + 1], f[16 * k + 2], f[16 * k + 3], f[16 * k + 4],
                                            f[16 * k + 5], f[16 * k + 6], f[16 * k + 7]),
                                pack_bf16x8(f[16 * k + 8], f[16 * k + 9], f[16 * k + 10], f[16 * k + 11], f[16 * k + 12],
                                            f[16 * k + 13], f[16 * k + 14], f[16 * k + 15]));
              }
            };
            if (kMath && p.epi == EPI_ACT_DUAL) {
              __nv_bfloat16* d2 = reinterpret_cast<__nv_bfloat16*>(p.C2) + crow + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
              put(d1, v);
              if (p.act) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
              }
              put(d2, v);
            } else {
              put(d1, v);
            }
          }
        } else if (kMath && p.tma_store == 2) {
          // bf16 store staged through shared memory and written by the warp itself: the 32 x 32 box is transposed so that
          // four lanes write one full 64-byte row segment (fire-and-forget st.global, full sectors).  The math tails use
          // this instead of the TMA store: their stores queue behind the producer's prefetched operand loads in the TMA
          // unit, and waiting for a staging buffer to be read cost ~2 us per chunk (the first fused GELU epilogue ran at
          // 0.45x of the plain GEMM for that reason).
          const int row0 = (mb * kCtas + cta_rank) * kBlockM + q * 32;
          if (row0 < p.M && col0 < p.N) {  // warp-uniform
            if (p.bias != nullptr) {
              const int ncols = min(32, p.N - col0);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += p.bias[1LL * bz * p.strideBias + col0 + j];
            }
            const int sw = (lane >> 1) & 3;
            auto write_box = [&](const uint8_t* buf, void* base) {
              __syncwarp();
              const int pc = lane & 3;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = (lane >> 2) + 8 * i;
                const uint4 val = *reinterpret_cast<const uint4*>(buf + r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
                if (row0 + r < p.M && col0 + pc * 8 < p.N)
                  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + 1LL * bz * p.strideC +
                                            1LL * (row0 + r) * p.ldc + col0 + pc * 8) = val;
              }
              __syncwarp();
            };
            if (p.epi == EPI_ACT_DUAL) {
              // the activation is taken on the bf16-rounded pre-activation so that backward (which re-reads C)
              // differentiates the same function
              auto stage_dual = [&](auto tanh_tag) {
                constexpr bool kTanh = decltype(tanh_tag)::value;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float x[8], y[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    x[e] = bf16_round(v[8 * j + e]);
                    y[e] = gelu_fast<kTanh>(x[e]);
                  }
                  *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((j ^ sw) << 4)) =
                      pack_bf16x8(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
                  *reinterpret_cast<uint4*>(my_stage + Cfg::kStoreBufBytes + lane * 64 + ((j ^ sw) << 4)) =
                      pack_bf16x8(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7]);
                }
              };
              if (p.act) stage_dual(std::true_type{});
              else stage_dual(std::false_type{});
              write_box(my_stage, p.C);
              write_box(my_stage + Cfg::kStoreBufBytes, p.C2);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((j ^ sw) << 4)) =
                    pack_bf16x8(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3], v[8 * j + 4], v[8 * j + 5],
                                v[8 * j + 6], v[8 * j + 7]);
              write_box(my_stage, p.C);
            }
          }
        } else if (!kMath && p.tma_store) {
          // bf16 store through shared memory: each lane (= output row) drops its 64 bytes into a 64B-swizzled 32 x 32
          // box and one lane hands the box to the TMA, which writes full lines and clips at the tensor edge.  Direct
          // st.global from this layout is one half-sector request per lane per store and kept the L1->XBAR port ~70 %
          // busy (ncu, profiles/r01_ncu_gemm_full.md).
          const int row0 = (mb * kCtas + cta_rank) * kBlockM + q * 32;
          if (row0 < p.M && col0 < p.N) {  // warp-uniform
            if (p.bias != nullptr) {
              const int ncols = min(32, p.N - col0);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += p.bias[1LL * bz * p.strideBias + col0 + j];
            }
            const int sw = (lane >> 1) & 3;
            {
              uint8_t* buf = my_stage + (store_it & 1) * Cfg::kStoreBufBytes;
              if (lane == 0) tma_store_wait_read<1>();  // the store issued from this buffer two chunks ago has read it
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 a;
                a.x = pack_bf16(v[8 * j], v[8 * j + 1]);
                a.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                a.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
                a.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(buf + lane * 64 + ((j ^ sw) << 4)) = a;
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&tmC, buf, col0, row0, bz);
                tma_store_commit();
              }
              ++store_it;
            }
          }
        } else if (!kMath && p.split_ws != nullptr) {
          // deterministic split-K: this split's partial tile goes to the workspace [split][batch][M][N] with plain stores
          // (zeros for an empty split); splitk_reduce_kernel adds the splits up in a fixed order
          if (row_ok && col0 < p.N) {
            float* dst = p.split_ws + ((1LL * sp * p.batch + bz) * p.M + row) * p.N + col0;
            const int ncols = min(32, p.N - col0);
            if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = has_k ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) dst[j] = has_k ? v[j] : 0.f;
            }
          }
        } else
        // No divergent `continue`: every lane must reach the next (warp-aligned) tcgen05 instruction together.
        if (row_ok && col0 < p.N && has_k) {
          const int ncols = min(32, p.N - col0);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) v[j] += p.bias[1LL * bz * p.strideBias + col0 + j];
          }
          if (!kMath && p.epi == EPI_ATOMIC_F32) {
            float* dst = reinterpret_cast<float*>(p.C) + crow + col0;
            if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)  // static indices only: a runtime index would push v[] into local memory
                if (j < ncols) atomicAdd(dst + j, v[j]);
            }
          } else if (kMath && p.epi == EPI_ACT_DUAL) {
            // C = pre-activation (bf16), C2 = act(pre) (bf16); the activation is taken on the bf16-rounded
            // pre-activation so that backward (which re-reads C) differentiates the same function.
            __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(p.C) + crow + col0;
            __nv_bfloat16* d2 = reinterpret_cast<__nv_bfloat16*>(p.C2) + crow + col0;
            const bool vec = ncols == 32 && ((reinterpret_cast<uintptr_t>(d1) & 15) == 0) &&
                             ((reinterpret_cast<uintptr_t>(d2) & 15) == 0);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float x[8], y[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = bf16_round(v[j + e]);
              if (p.act) {
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = gelu_tanh_fast(x[e]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = gelu_erf_fast(x[e]);
              }
              const uint4 a = pack_bf16x8(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
              const uint4 g = pack_bf16x8(y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7]);
              if (vec) {
                *reinterpret_cast<uint4*>(d1 + j) = a;
                *reinterpret_cast<uint4*>(d2 + j) = g;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (j + e < ncols) {
                    const uint32_t aw = (e >> 1) == 0 ? a.x : (e >> 1) == 1 ? a.y : (e >> 1) == 2 ? a.z : a.w;
                    const uint32_t gw = (e >> 1) == 0 ? g.x : (e >> 1) == 1 ? g.y : (e >> 1) == 2 ? g.z : g.w;
                    const unsigned short ab = static_cast<unsigned short>((e & 1) ? (aw >> 16) : (aw & 0xffffu));
                    const unsigned short gb = static_cast<unsigned short>((e & 1) ? (gw >> 16) : (gw & 0xffffu));
                    d1[j + e] = __ushort_as_bfloat16(ab);
                    d2[j + e] = __ushort_as_bfloat16(gb);
                  }
              }
            }
          } else {
            if (!kMath && p.epi == EPI_RESID_F32) {
              if (p.C2 != nullptr) {  // bf16 copy of the raw GEMM result (needed by backward for d(gate))
                __nv_bfloat16* d2 = reinterpret_cast<__nv_bfloat16*>(p.C2) + crow + col0;
                if (ncols == 32 && ((reinterpret_cast<uintptr_t>(d2) & 15) == 0)) {
#pragma unroll
                  for (int j = 0; j < 32; j += 8) {
                    *reinterpret_cast<uint4*>(d2 + j) =
                        pack_bf16x8(v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (j < ncols) d2[j] = __float2bfloat16_rn(v[j]);
                }
              }
              if (gate_row != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) v[j] *= gate_row[col0 + j];
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += resv[j].x; v[4 * j + 1] += resv[j].y; v[4 * j + 2] += resv[j].z; v[4 * j + 3] += resv[j].w;
              }
            }
            if (kMath || p.epi == EPI_STORE_BF16 || p.epi == EPI_ACT_GRAD) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + crow + col0;
              if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  *reinterpret_cast<uint4*>(dst + j) =
                      pack_bf16x8(v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
              }
            } else {  // EPI_STORE_F32 / EPI_RESID_F32
              float* dst = reinterpret_cast<float*>(p.C) + crow + col0;
              if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) dst[j] = v[j];
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtas == 2) mbar_arrive_leader(&tempty_bar[as]);
        else mbar_arrive(&tempty_bar[as]);
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait<0>();  // staging memory must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kCtas == 2) cluster_sync_all();  // nobody exits / frees TMEM while the peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kCtas == 2) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// C[b][rowmap(m)][n] += sum_sp ws[sp][b][m][n], sp = 0, 1, ... (deterministic split-K, second pass)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ C, int splits, int batch, int M, int N,
                     long long ldc, long long strideC, int row_interleave) {
  const long long total = 1LL * batch * M * N;
  for (long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int n = static_cast<int>(i % N);
    const long long bm = i / N;
    const int m = static_cast<int>(bm % M);
    const int b = static_cast<int>(bm / M);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += ws[1LL * sp * total + i];
    int row = m;
    if (row_interleave > 0) {
      const int blk = m >> 6, in = m & 63;
      row = (in < 32 ? 0 : row_interleave) + 32 * blk + (in & 31);
    }
    C[1LL * b * strideC + 1LL * row * ldc + n] += s;
  }
}

// ---------------------------------------------------------------------------------------------- host

// bf16 tensor viewed as [batch][rows][cols] (cols contiguous), box = [1][box_rows][64], 128B swizzle.
static int make_map(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long batch,
                    long long ld, long long batch_stride, int box_rows) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 8) != 0 || (batch > 1 && (batch_stride % 8) != 0))
    return md_set_error(MD_ERR_INVALID, "gemm operand must be 16-byte aligned with ld %% 8 == 0");
  const TmapKey key = make_tmap_key(ptr, cols, rows, batch, ld, batch > 1 ? batch_stride : rows * ld, 64, box_rows,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  CUresult r = cached_tensor_map(map, key);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) cols=%lld rows=%lld ld=%lld", (int)r, cols,
             rows, ld);
    return md_set_error(MD_ERR_CUDA, msg);
  }
  return 0;
}

// bf16 output viewed as [batch][M][N]; box = [1][32 rows][32 cols], 64B swizzle (what the epilogue warps stage).
static int make_store_map(CUtensorMap* map, void* ptr, long long cols, long long rows, long long batch, long long ld,
                          long long batch_stride) {
  const TmapKey key = make_tmap_key(ptr, cols, rows, batch, ld, batch > 1 ? batch_stride : rows * ld, 32, 32,
                                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE);
  if (cached_tensor_map(map, key) != CUDA_SUCCESS)
    return md_set_error(MD_ERR_CUDA, "cuTensorMapEncodeTiled failed for the output map");
  return 0;
}

template <int BLOCK_N, bool kMN, int kCtas, int kEpiW = 4>
static int launch(const md_gemm_args* a, GemmDev dev, int sm_count, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, kCtas, kEpiW>;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if (!kMN) {
    rc = make_map(&tmA, a->A, a->K, a->M, a->batch, a->lda, a->strideA, kBlockM);
    if (rc) return rc;
    rc = make_map(&tmB, a->B, a->K, a->N, a->batch, a->ldb, a->strideB, BLOCK_N / kCtas);
    if (rc) return rc;
  } else {
    rc = make_map(&tmA, a->A, a->M, a->K, a->batch, a->lda, a->strideA, kBlockK);
    if (rc) return rc;
    rc = make_map(&tmB, a->B, a->N, a->K, a->batch, a->ldb, a->strideB, kBlockK);
    if (rc) return rc;
  }
  // store mode of the bf16-output epilogues (MD_GEMM_TMA_STORE; the math tails additionally MD_GEMM_MATH_STORE)
  static int tma_store_env = -1;
  if (tma_store_env == -1) {
    const char* e = getenv("MD_GEMM_TMA_STORE");
    tma_store_env = e ? atoi(e) : 3;  // 3 = 256-bit register-direct stores (default), 1 = TMA store, 0 = 16-byte direct
  }
  static int debug_env = -1;
  if (debug_env == -1) {
    const char* e = getenv("MD_GEMM_DEBUG");
    debug_env = e ? atoi(e) : 0;
  }
  dev.debug = debug_env;
  dev.tma_store = 0;
  const bool bf16_out = a->epilogue == EPI_STORE_BF16 || a->epilogue == EPI_ACT_GRAD || a->epilogue == EPI_ACT_DUAL;
  const bool dual = a->epilogue == EPI_ACT_DUAL;
  const bool aligned_out = (reinterpret_cast<uintptr_t>(a->C) & 15) == 0 && (a->ldc % 8) == 0 && (a->N % 8) == 0 &&
                           (a->batch == 1 || (a->strideC % 8) == 0) &&
                           (!dual || (reinterpret_cast<uintptr_t>(a->C2) & 15) == 0) &&
                           (a->epilogue != EPI_ACT_GRAD || (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0);
  // staged + coalesced st.global (tma_store = 2): the math tails' fallback when the view is only 16-byte addressable
  const bool math = a->epilogue == EPI_ACT_DUAL || a->epilogue == EPI_ACT_GRAD;
  // 32-byte pieces for the register-direct 256-bit path
  const bool aligned32 = aligned_out && (reinterpret_cast<uintptr_t>(a->C) & 31) == 0 && (a->ldc % 16) == 0 &&
                         (a->N % 16) == 0 && (a->batch == 1 || (a->strideC % 16) == 0) &&
                         (!dual || (reinterpret_cast<uintptr_t>(a->C2) & 31) == 0) &&
                         (a->epilogue != EPI_ACT_GRAD || (reinterpret_cast<uintptr_t>(a->aux) & 31) == 0);
  static int math_store_env = -1;  // MD_GEMM_MATH_STORE: 3 = 256-bit direct (default), 2 = staged + coalesced
  if (math_store_env == -1) {
    const char* e = getenv("MD_GEMM_MATH_STORE");
    math_store_env = e ? atoi(e) : 3;
  }
  if (tma_store_env && bf16_out && aligned32 && ((math && math_store_env == 3) || tma_store_env == 3)) {
    dev.tma_store = 3;
    tmC = tmA;
  } else if (tma_store_env && math && aligned_out) {
    dev.tma_store = 2;
    tmC = tmA;
  } else if (tma_store_env && a->epilogue == EPI_STORE_BF16 && (reinterpret_cast<uintptr_t>(a->C) & 15) == 0 &&
             (a->ldc % 8) == 0 && (a->batch == 1 || (a->strideC % 8) == 0)) {
    rc = make_store_map(&tmC, a->C, a->N, a->M, a->batch, a->ldc, a->strideC);
    if (rc) return rc;
    dev.tma_store = 1;
  } else {
    tmC = tmA;  // unused
  }
  auto kern = gemm_tcgen05_kernel<BLOCK_N, kMN, kCtas, kEpiW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  const long long m_blocks = (a->M + kBlockM * kCtas - 1) / (kBlockM * kCtas);
  const long long n_blocks = (a->N + BLOCK_N - 1) / BLOCK_N;
  const long long tiles = a->batch * dev.splits * m_blocks * n_blocks;
  const long long slots = sm_count / kCtas;  // CTAs (1-CTA mode) or CTA pairs resident at once
  const int grid = static_cast<int>((tiles < slots ? tiles : slots) * kCtas);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, dev);
  if (e != cudaSuccess) return md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}

}  // namespace md

extern "C" int md_gemm_bf16(const md_gemm_args* a, void* stream_) {
  using namespace md;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a == nullptr || a->A == nullptr || a->B == nullptr || a->C == nullptr)
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: null operand");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->batch <= 0) return 0;  // empty problem: nothing to do
  if (a->epilogue < 0 || a->epilogue >= EPI_COUNT) return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: bad epilogue");
  if (a->epilogue == EPI_RESID_F32 && a->res == nullptr)
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: residual epilogue needs res");
  if (a->epilogue == EPI_ACT_DUAL && a->C2 == nullptr)
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: activation epilogue needs C2");
  if (a->epilogue == EPI_ACT_GRAD && a->aux == nullptr)
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: activation-gradient epilogue needs aux (the saved pre-activation)");
  if (a->epilogue == EPI_ACT_GRAD && a->bias != nullptr)
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: the activation-gradient epilogue takes no bias");
  if (a->epilogue == EPI_SWIGLU || a->epilogue == EPI_SWIGLU_GRAD) {
    const bool fwd = a->epilogue == EPI_SWIGLU;
    const void* second = fwd ? a->C2 : a->aux;
    const int64_t ld2 = fwd ? (a->ldc2 > 0 ? a->ldc2 : a->N / 2) : a->ldc;
    if (second == nullptr || a->bias != nullptr || a->layout != MD_GEMM_NT || a->splits > 1)
      return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: SwiGLU epilogues need C2 (forward) / aux (backward), NT layout, no bias");
    if (a->N % (fwd ? 64 : 32) != 0 || (a->ldc % 16) != 0 || (ld2 % 16) != 0 ||
        ((reinterpret_cast<uintptr_t>(a->C) | reinterpret_cast<uintptr_t>(second)) & 31) != 0 ||
        (a->batch > 1 && ((a->strideC % 16) != 0 || (fwd && (a->strideC2 % 16) != 0))))
      return md_set_error(MD_ERR_UNSUPPORTED,
                          "md_gemm_bf16: SwiGLU epilogues need 32-byte aligned outputs, pitches % 16 == 0 and whole 32-column blocks");
  }
  if (a->row_interleave != 0 && (a->epilogue != EPI_ATOMIC_F32 || a->row_interleave % 32 != 0 || a->M != 2 * a->row_interleave))
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: row_interleave = f needs the atomic epilogue, f % 32 == 0 and M == 2 f");
  if (a->gate != nullptr && a->rows_per_gate <= 0)
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: gate needs rows_per_gate > 0");
  int splits = a->splits > 0 ? a->splits : 1;
  if (splits > 1 && a->epilogue != EPI_ATOMIC_F32)
    return md_set_error(MD_ERR_INVALID, "md_gemm_bf16: split-K needs the atomic epilogue");

  int dev_id = 0, sm_count = 0;
  cudaError_t e = cudaGetDevice(&dev_id);
  if (e != cudaSuccess) return md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
  static int cached_sm[64] = {0};
  if (dev_id < 64 && cached_sm[dev_id] > 0) sm_count = cached_sm[dev_id];
  else {
    int major = 0;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev_id);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev_id);
    if (major != 10) return md_set_error(MD_ERR_UNSUPPORTED, "md_gemm_bf16: requires an sm_100a device (B200)");
    if (dev_id < 64) cached_sm[dev_id] = sm_count;
  }

  if (a->sm_limit > 0 && a->sm_limit < sm_count) sm_count = a->sm_limit >= 2 ? a->sm_limit : 2;

  GemmDev dev;
  dev.C = a->C; dev.C2 = a->C2;
  dev.bias = reinterpret_cast<const float*>(a->bias);
  dev.res = reinterpret_cast<const float*>(a->res);
  dev.gate = reinterpret_cast<const float*>(a->gate);
  dev.aux = a->aux;
  dev.ldc2 = a->ldc2 > 0 ? a->ldc2 : a->N / 2;
  dev.strideC2 = a->strideC2;
  dev.row_interleave = static_cast<int>(a->row_interleave);
  dev.split_ws = nullptr;
  dev.M = static_cast<int>(a->M); dev.N = static_cast<int>(a->N); dev.K = static_cast<int>(a->K);
  dev.batch = static_cast<int>(a->batch); dev.splits = splits;
  dev.ldc = a->ldc; dev.strideC = a->strideC; dev.strideBias = a->strideBias;
  dev.ldgate = a->ldgate; dev.rows_per_gate = static_cast<int>(a->rows_per_gate > 0 ? a->rows_per_gate : 1);
  dev.epi = a->epilogue;
  dev.res_mod = static_cast<int>(a->res_mod);
  dev.act = a->act;
  dev.alpha = a->alpha == 0.0f ? 1.0f : a->alpha;

  const bool mn = a->layout == MD_GEMM_TN;
  const long long m_blocks = (a->M + kBlockM - 1) / kBlockM;
  const long long kb_total = (a->K + kBlockK - 1) / kBlockK;
  auto tiles_for = [&](int bn) { return a->batch * m_blocks * ((a->N + bn - 1) / bn); };
  // Tile-N: 256 halves the B-operand smem traffic per MMA (a 128x128 tile is smem-bandwidth bound), so prefer it
  // whenever the padding waste is small and there is enough work to spread over the SMs.
  const long long n256 = (a->N + 255) / 256 * 256;
  // <= 20 % padded columns: N = 640 / 896 (the 0.625 / 0.875 attention widths of MicroDiT_XL_2) run 3 / 4 tiles of 256 with
  // a partly empty last one at ~0.8 of the full-tile rate; the all-128 fallback measured 0.45 (profiles/r01_per_op_c2_final.csv)
  bool use256 = (a->N >= 256) && (n256 - a->N) * 5 <= a->N;
  if (a->splits == 0 && a->epilogue == EPI_ATOMIC_F32) {
    // auto split of the reduction: minimise  waves * (k-blocks per split + fixed per-tile cost)
    const long long t = tiles_for(use256 ? 256 : 128);
    const long long smax = kb_total / 8 > 0 ? (kb_total / 8 < 64 ? kb_total / 8 : 64) : 1;
    double best = 1e30;
    for (long long sp = 1; sp <= smax; ++sp) {
      const long long units = t * sp;
      const long long waves = (units + sm_count - 1) / sm_count;
      const double cost = static_cast<double>(waves) * (static_cast<double>((kb_total + sp - 1) / sp) + 10.0);
      if (cost < best - 1e-9) { best = cost; splits = static_cast<int>(sp); }
    }
    dev.splits = splits;
  } else if (use256 && tiles_for(256) * splits < sm_count && tiles_for(128) * splits > tiles_for(256) * splits) {
    use256 = false;  // too few 256-wide tiles to occupy the machine: smaller tiles win
  }
  // CTA pairs (cta_group::2, 256-row tiles) whenever there are at least two 128-row blocks; MD_GEMM_CTAS=1|2 overrides.
  static int ctas_forced = -1;
  if (ctas_forced == -1) {
    const char* e = getenv("MD_GEMM_CTAS");
    ctas_forced = e ? atoi(e) : 0;
  }
  const bool pair = ctas_forced == 2 || (ctas_forced == 0 && a->M > kBlockM);
  // eight epilogue warps for the tails that do real math per element (NT only; the 4-warp kernels do not carry them)
  const bool swiglu = a->epilogue == EPI_SWIGLU || a->epilogue == EPI_SWIGLU_GRAD;  // only the 8-warp kernels carry them
  const bool math_tail = swiglu || a->epilogue == EPI_ACT_DUAL || a->epilogue == EPI_ACT_GRAD;
  if (math_tail && mn) return md_set_error(MD_ERR_UNSUPPORTED, "md_gemm_bf16: the activation / SwiGLU epilogues need the NT layout");
  // deterministic mode: a split reduction goes through per-split partial tiles in the workspace and a fixed-order second
  // pass; if the workspace cannot hold them the reduction is not split (one writer per element: deterministic, slower)
  if (det_enabled() && a->epilogue == EPI_ATOMIC_F32 && dev.splits > 1) {
    dev.split_ws = det_workspace(sizeof(float) * static_cast<size_t>(dev.splits) * a->batch * a->M * a->N);
    if (dev.split_ws == nullptr) dev.splits = 1;
  }
  auto run = [&]() -> int {
    if (pair) {
      if (mn) return use256 ? launch<256, true, 2>(a, dev, sm_count, stream) : launch<128, true, 2>(a, dev, sm_count, stream);
      if (math_tail)
        return use256 ? launch<256, false, 2, 8>(a, dev, sm_count, stream) : launch<128, false, 2, 8>(a, dev, sm_count, stream);
      return use256 ? launch<256, false, 2>(a, dev, sm_count, stream) : launch<128, false, 2>(a, dev, sm_count, stream);
    }
    if (mn) return use256 ? launch<256, true, 1>(a, dev, sm_count, stream) : launch<128, true, 1>(a, dev, sm_count, stream);
    if (math_tail)
      return use256 ? launch<256, false, 1, 8>(a, dev, sm_count, stream) : launch<128, false, 1, 8>(a, dev, sm_count, stream);
    return use256 ? launch<256, false, 1>(a, dev, sm_count, stream) : launch<128, false, 1>(a, dev, sm_count, stream);
  };
  if (int rc = run()) return rc;
  if (dev.split_ws != nullptr) {
    const long long total = 1LL * a->batch * a->M * a->N;
    const int blocks = static_cast<int>(total / 256 + 1 < 148LL * 16 ? total / 256 + 1 : 148LL * 16);
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(dev.split_ws, reinterpret_cast<float*>(a->C), dev.splits, dev.batch, dev.M,
                                                     dev.N, dev.ldc, dev.strideC, dev.row_interleave);
    return check_launch("md_gemm_bf16 (deterministic split-K reduction)");
  }
  return 0;
}
