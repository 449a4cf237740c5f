// Attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA) for head_dim 64 -- forward.
//
// Contract = md_attn_fwd: softmax(Q K^T / sqrt(hd)) V, non-causal (F.scaled_dot_product_attention at reference
// utils.py:188-193 self, 127-132 cross); q / k / v are column slices of the packed projection buffers; lse in the
// log2 domain.  Envelope: all keys of one (sample, head) fit one S tile, i.e. Tk <= 256 (every sequence of the res-256
// configs: 64, 77, 256) -- longer sequences stay on the mma.sync kernels (attn.cu).
//
// Persistent, warp-specialised, double-buffered.  One CTA per SM loops over tiles = (sample, head or head pair,
// 128-query block); 12 warps:
//   warp 0      TMA producer: Q / K (one barrier) and V (another) of tile i+2 as soon as tile i's PV has retired
//   warp 1      one thread issues S = Q K^T (M 128, N = keys, K 64) and O = P V (M 128, N 64, K = keys)
//   warp 2      TMEM allocator (512 columns: two 256-column buffers; O of a tile overwrites the first 64 columns of
//               its own S once the softmax has consumed it)
//   warps 4-7   softmax group 0 (even tiles), warps 8-11 softmax group 1 (odd tiles): one query row per thread,
//               whole-row softmax straight from TMEM in two passes (max, then exp2 / sum -- no online rescale because
//               all of S is resident), P written as bf16 into 128B-swizzled K-major atoms that overwrite the dead Q / K
//               tiles in shared memory, then O read back, scaled by 1 / rowsum and stored.
// While group 0 does the softmax of tile i the tensor core runs S of tile i+1 and PV of tile i-1.
//
// Tq <= 64 ("packed" mode, the 64-token backbone of mask 0.75): two heads share one 128-row tile -- rows 0-63 are head
// h, rows 64-127 head h+1, the keys of both heads are stacked along N (S is 128 x 2*keys, each row uses its own head's
// column range) and P is written block-diagonal so that one PV chain over the stacked V tiles serves both heads.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tensormap.cuh"

namespace md {
namespace attn_tc {

constexpr int kQ = 128;               // query rows per tile (UMMA M)
constexpr int kHd = 64;               // head_dim == one 128-byte swizzle row of bf16
constexpr int kMaxCols = 256;         // S columns per tile (keys, or 2 x keys in packed mode)
constexpr int kThreads = 384;            // backward: 4 control warps + 2 groups of 4
constexpr int kFwdThreads = 640;         // forward: 4 control warps + 2 groups of 8 (two threads per query row)
constexpr int kTmemCols = 512;
constexpr int kBytesQ = kQ * kHd * 2;            // 16 KB
constexpr int kBytesPAtom = kQ * 64 * 2;         // 16 KB: 128 rows x 64 keys
// A shared-memory stage = [Q (16 KB) | K (keys x 128 B)] -- later overwritten by P (one 16 KB atom per 64 keys) -- and V
// (keys x 128 B), sized at launch from the sequence lengths: 96 KB for 256 keys (two stages), 43-68 KB for the 64 / 77-key
// shapes (three stages: the TMA load of tile i+1 then no longer waits for the PV MMA of tile i-1 to free a stage, which had
// put the whole load latency on every tile's critical path).  TMEM stays double-buffered (tile parity).
constexpr int kMaxStages = 3;
constexpr int kOffBar = 208 * 1024;               // >= stages x stage bytes for every shape (host checks)
constexpr int kOffXch = kOffBar + 256;            // 2 groups x [max | sum] x 2 halves x 128 rows of fp32 = 4 KB
constexpr int kSmemBytes = kOffXch + 4096 + 1024;

// Optional phase timeline (MD_ATTN_DEBUG=1): block 0 logs clock64() at fixed points; md_attn_debug_dump() reads it back.
__device__ long long g_dbg[8 * 512];
#define DBG(slot, idx) \
  do { if (p.dbg && blockIdx.x == 0 && (idx) < 512) g_dbg[(slot) * 512 + (idx)] = clock64(); } while (0)

struct FwdParams {
  int dbg, o256;
  int ns, stage_bytes, v_off;   // shared-memory ring: stages, bytes per stage, offset of V inside a stage
  __nv_bfloat16* o;
  long long ldo;
  float* lse;
  int H, Tq, Tk, n_pad, packed, q_blocks, head_tiles;
  long long tiles;
  float scale_log2;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void decode_tile(long long t, const FwdParams& p, int& b, int& h0, int& qb) {
  qb = static_cast<int>(t % p.q_blocks);
  t /= p.q_blocks;
  const int ht = static_cast<int>(t % p.head_tiles);
  b = static_cast<int>(t / p.head_tiles);
  h0 = p.packed ? 2 * ht : ht;
}

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* qk_full = bars;        // [3] TMA bytes of Q + K             (indexed by shared-memory stage)
  uint64_t* v_full = bars + 3;     // [3] TMA bytes of V
  uint64_t* st_free = bars + 6;    // [3] PV of the tile retired: the stage's shared memory is free (tcgen05.commit)
  uint64_t* s_full = bars + 9;     // [2] S in TMEM (tcgen05.commit)     (indexed by tile parity = TMEM buffer = group)
  uint64_t* p_full = bars + 11;    // [2] P in shared memory, S consumed (8 warps)
  uint64_t* o_full = bars + 13;    // [2] O in TMEM (tcgen05.commit)
  uint64_t* o_free = bars + 15;    // [2] O read back: the TMEM buffer is free (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  const int ns = p.ns;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pad = p.n_pad;                          // keys of one head, padded to 16
  const int Tk = p.Tk;
  const float sl2 = p.scale_log2;
  const int n_s = p.packed ? 2 * n_pad : n_pad;       // S columns / PV reduction length of a tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&st_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 8);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t it = 0, st = 0, round = 0;   // st = it % ns, round = it / ns
      for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        DBG(0, 2 * it);
        if (round > 0) mbar_wait_sleep(&st_free[st], (round - 1) & 1);  // PV of the tile that last used this stage has retired
        DBG(0, 2 * it + 1);
        int b, h0, qb;
        decode_tile(t, p, b, h0, qb);
        uint8_t* sQ = smem + st * p.stage_bytes;
        uint8_t* sK = sQ + kBytesQ;
        uint8_t* sV = sQ + p.v_off;
        mbar_expect_tx(&qk_full[st], kBytesQ + n_s * 128);
        if (!p.packed) {
          tma_load_3d(&tmQ, &qk_full[st], sQ, h0 * kHd, qb * kQ, b);
          tma_load_3d(&tmK, &qk_full[st], sK, h0 * kHd, 0, b);
        } else {
          tma_load_3d(&tmQ, &qk_full[st], sQ, h0 * kHd, 0, b);
          tma_load_3d(&tmQ, &qk_full[st], sQ + kBytesQ / 2, (h0 + 1) * kHd, 0, b);
          tma_load_3d(&tmK, &qk_full[st], sK, h0 * kHd, 0, b);
          tma_load_3d(&tmK, &qk_full[st], sK + n_pad * 128, (h0 + 1) * kHd, 0, b);
        }
        mbar_expect_tx(&v_full[st], n_s * 128);
        tma_load_3d(&tmV, &v_full[st], sV, h0 * kHd, 0, b);
        if (p.packed) tma_load_3d(&tmV, &v_full[st], sV + n_pad * 128, (h0 + 1) * kHd, 0, b);
        if (++st == static_cast<uint32_t>(ns)) { st = 0; ++round; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp walks the loop (waits, descriptor arithmetic: warp-uniform values stay in uniform registers) and one
    // lane issues: with the role gated on a single lane the compiler rebuilt every operand through R2UR broadcast loops and
    // a UTCHMMA issue cost ~100 cycles -- more than a 128x64x16 MMA takes to execute.
    {
      const uint32_t idesc_s = umma_idesc_bf16(kQ, n_s, false, false);   // Q, K both K-major (head_dim contiguous)
      const uint32_t idesc_o = umma_idesc_bf16(kQ, kHd, false, true);    // P K-major, V MN-major (head_dim contiguous)
      const int ksteps = n_s / 16;
      // S of tile i and PV of tile i-1 are both outstanding most of the time; the issuer takes whichever has its inputs
      // ready instead of a fixed order (with S first, PV(i-1) sat behind the TMA load of tile i, whose stage had just been
      // freed by PV(i-2): the whole load latency ended up between P and O of every tile).
      uint32_t n_s_issued = 0, n_pv_issued = 0;       // tiles whose S / PV have been issued
      uint32_t st_s = 0, rd_s = 0, st_pv = 0, rd_pv = 0;  // shared-memory stage / round of the next S and the next PV tile
      const uint32_t my_tiles = blockIdx.x < p.tiles ? static_cast<uint32_t>((p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
      auto s_ready = [&]() {
        const uint32_t it = n_s_issued;
        if (!mbar_test(&qk_full[st_s], rd_s & 1)) return false;
        return it < 2 || mbar_test(&o_free[it & 1], ((it >> 1) - 1) & 1);  // O of the tile two back has left this TMEM buffer
      };
      auto pv_ready = [&]() {
        const uint32_t j = n_pv_issued;
        return mbar_test(&p_full[j & 1], (j >> 1) & 1) && mbar_test(&v_full[st_pv], rd_pv & 1);
      };
      auto issue_s = [&]() {
        const uint32_t it = n_s_issued;
        DBG(1, 6 * it + 1);
        tc_fence_after();
        const uint32_t aq = smem_u32(smem + st_s * p.stage_bytes);
        const uint64_t dq0 = umma_smem_desc(aq, 16, 1024), dk0 = umma_smem_desc(aq + kBytesQ, 16, 1024);
#pragma unroll
        for (int ks = 0; ks < kHd / 16; ++ks)
          if (elect_one()) umma_bf16(tmem_base + (it & 1) * kMaxCols, dq0 + 2 * ks, dk0 + 2 * ks, idesc_s, ks > 0 ? 1u : 0u);
        if (elect_one()) umma_commit(&s_full[it & 1]);
        DBG(1, 6 * it + 2);
        ++n_s_issued;
        if (++st_s == static_cast<uint32_t>(ns)) { st_s = 0; ++rd_s; }
      };
      auto issue_pv = [&]() {
        const uint32_t j = n_pv_issued;
        DBG(1, 6 * j + 4);
        tc_fence_after();
        const uint32_t ap = smem_u32(smem + st_pv * p.stage_bytes);
        const uint32_t av = ap + p.v_off;
        // descriptors advance by plain adds on the 14-bit (address >> 4) field: one UTCHMMA costs the issuing thread a few
        // instructions instead of a full descriptor rebuild (which made the issue loop, not the tensor pipe, the limit)
        const uint64_t da0 = umma_smem_desc(ap, 16, 1024);
        const uint64_t db0 = umma_smem_desc(av, 64 * 128, 1024);
        // Consecutive MMAs into the SAME accumulator serialise on the accumulator's read-modify-write latency (~100 cycles,
        // three times what a 128x64x16 MMA takes to execute): the reduction over the keys is therefore spread over four
        // accumulators (k-step mod 4 -> columns 64 * j of the dead S buffer) that the epilogue adds up.
        const uint32_t d = tmem_base + (j & 1) * kMaxCols;
        for (int at = 0; at * 4 < ksteps; ++at) {
          const uint64_t da = da0 + static_cast<uint64_t>(at) * (kBytesPAtom >> 4);
          const uint64_t db = db0 + static_cast<uint64_t>(at) * (4 * 16 * 128 >> 4);
          const uint32_t acc = at > 0 ? 1u : 0u;
          if (elect_one()) umma_bf16(d, da, db, idesc_o, acc);
          if (at * 4 + 1 < ksteps) { if (elect_one()) umma_bf16(d + 64, da + 2, db + 128, idesc_o, acc); }
          if (at * 4 + 2 < ksteps) { if (elect_one()) umma_bf16(d + 128, da + 4, db + 256, idesc_o, acc); }
          if (at * 4 + 3 < ksteps) { if (elect_one()) umma_bf16(d + 192, da + 6, db + 384, idesc_o, acc); }
        }
        if (elect_one()) umma_commit(&o_full[j & 1]);
        if (elect_one()) umma_commit(&st_free[st_pv]);   // the stage's shared memory (P, V) is free again
        DBG(1, 6 * j + 5);
        ++n_pv_issued;
        if (++st_pv == static_cast<uint32_t>(ns)) { st_pv = 0; ++rd_pv; }
      };
      uint32_t spins = 0;
      while (n_pv_issued < my_tiles) {
        // all lanes evaluate the same barrier states (warp-uniform control flow)
        const bool can_s = n_s_issued < my_tiles && n_s_issued < n_pv_issued + 2;   // TMEM: at most two tiles past PV
        if (can_s && s_ready()) { issue_s(); spins = 0; continue; }
        if (n_pv_issued < n_s_issued && pv_ready()) { issue_pv(); spins = 0; continue; }
        if (++spins > 400000000u) {
          if (lane == 0) printf("md: attention forward issuer stalled (block %d, S %u PV %u of %u)\n", blockIdx.x, n_s_issued, n_pv_issued, my_tiles);
          __trap();
        }
        __nanosleep(20);
      }
    }
  } else if (warp >= 4) {
    // ================================ softmax / epilogue groups ================================
    // A group = 8 warps = two threads per query row: warps 0-3 of the group take the first half of the row's key columns,
    // warps 4-7 the second half (each warp may only touch the TMEM lane quarter warp % 4).  One warp issuing MUFU.EX2 back
    // to back gets one every ~18 cycles while the pipe takes one every 8: with four softmax warps per scheduler (two groups
    // x two halves) the exp pipe stays busy.  The halves exchange the partial row maximum and row sum through shared memory.
    const int gw = (warp - 4) & 7;
    const int grp = (warp - 4) >> 3;              // group g handles tiles with (iteration & 1) == g -> stage g
    const int half = gw >> 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch (hardware rule)
    const int row = quarter * 32 + lane;          // row inside the tile == TMEM lane
    const int hi = (p.packed && row >= 64) ? 1 : 0;
    const int split = min(n_pad, ((n_pad >> 1) + 15) & ~15);   // columns [0, split) -> half 0, [split, n_pad) -> half 1
    const int my0 = half ? split : 0;             // this thread's columns inside its head's range: [my0, my1)
    const int my1 = half ? n_pad : split;
    const int col0 = (hi ? n_pad : 0) + my0;      // first S column of this thread
    const int chunks = (my1 - my0 + 31) / 32;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + grp * kMaxCols;
    const uint32_t prow0 = smem_u32(smem) + row * 128;   // this row of P inside stage 0 (shared-window address)
    const int sw = row & 7;
    float* xch = reinterpret_cast<float*>(smem + kOffXch) + grp * 512;   // [max | sum][half][128 rows]
    const int s = grp;
    const int nacc = min(4, n_s / 16);            // PV accumulators in use
    const uint32_t bar_id = 1 + grp;
    uint32_t n = 0;
    uint32_t st = grp % ns;   // shared-memory stage of this group's next tile: (2 n + grp) % ns
    for (long long t = blockIdx.x + 1LL * grp * gridDim.x; t < p.tiles; t += 2LL * gridDim.x, ++n) {
      const uint32_t prow = prow0 + st * p.stage_bytes;
      st = (st + 2) % ns;
      int b, h0, qb;
      decode_tile(t, p, b, h0, qb);
      const int h = h0 + hi;
      const int qrow = p.packed ? (row & 63) : qb * kQ + row;
      const bool q_ok = qrow < p.Tq;

      const int dslot = 2 + grp, dlog = (gw == 0 && lane == 0);
      if (dlog) DBG(dslot, 10 * n);
      mbar_wait_sleep(&s_full[s], n & 1);
      if (dlog) DBG(dslot, 10 * n + 1);
      tc_fence_after();
      // ---- pass 1: maximum over this thread's columns, then over both halves
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains
      for (int c = 0; c < chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(trow + col0 + c * 32, r);
        tmem_ld_wait();
        const int kc = my0 + c * 32;              // key index of column 0 of this chunk
        if (kc + 32 <= my1 && kc + 32 <= Tk) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (kc + j < my1 && kc + j < Tk) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(r[j]));
        }
      }
      float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      xch[half * 128 + row] = m;
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      m = fmaxf(m, xch[(half ^ 1) * 128 + row]);
      const float m2 = m * sl2;
      if (dlog) DBG(dslot, 10 * n + 2);
      // ---- pass 2: P = exp2(s * scale - max), partial row sum, bf16 P into the swizzled K-major atoms
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(trow + col0 + c * 32, r);
        tmem_ld_wait();
        float pv[32];
        const int kc = my0 + c * 32;
        if (kc + 32 <= my1 && kc + 32 <= Tk) {
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = ex2_approx(fmaf(__uint_as_float(r[j]), sl2, -m2));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            pv[j] = (kc + j < my1 && kc + j < Tk) ? ex2_approx(fmaf(__uint_as_float(r[j]), sl2, -m2)) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) ls[j & 3] += pv[j];
        const int k8 = (col0 >> 3) + c * 4;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (kc + g * 8 < my1) {                 // split and n_pad are multiples of 16: whole 8-key groups
            const int key8 = k8 + g;
            st_shared_v4(prow + (key8 >> 3) * kBytesPAtom + (((key8 & 7) ^ sw) << 4), pack2(pv[8 * g + 0], pv[8 * g + 1]),
                         pack2(pv[8 * g + 2], pv[8 * g + 3]), pack2(pv[8 * g + 4], pv[8 * g + 5]),
                         pack2(pv[8 * g + 6], pv[8 * g + 7]));
          }
        }
      }
      const float lpart = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      xch[256 + half * 128 + row] = lpart;        // read by the other half after o_full (ordered by the barrier chain)
      if (dlog) DBG(dslot, 10 * n + 3);
      if (p.packed) {  // block-diagonal P: zero this thread's share of the row's columns of the other head
        const int z0 = ((hi ? 0 : n_pad) + my0) >> 3;
        for (int g = 0; g < ((my1 - my0) >> 3); ++g) {
          const int key8 = z0 + g;
          st_shared_v4(prow + (key8 >> 3) * kBytesPAtom + (((key8 & 7) ^ sw) << 4), 0u, 0u, 0u, 0u);
        }
      }
      fence_proxy_async_smem();  // the UMMA reads P through the async proxy
      if (dlog) DBG(dslot, 10 * n + 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[s]);
      if (dlog) DBG(dslot, 10 * n + 5);

      // ---- epilogue: this half's 32 columns of O / rowsum
      mbar_wait_sleep(&o_full[s], n & 1);
      if (dlog) DBG(dslot, 10 * n + 6);
      tc_fence_after();
      uint32_t r0[32];
      tmem_ld_32x32(trow + half * 32, r0);
      const float l = lpart + xch[256 + (half ^ 1) * 128 + row];
      tmem_ld_wait();
      for (int j = 1; j < nacc; ++j) {            // the other PV accumulators (k-step mod 4)
        uint32_t rj[32];
        tmem_ld_32x32(trow + j * 64 + half * 32, rj);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) r0[e] = __float_as_uint(__uint_as_float(r0[e]) + __uint_as_float(rj[e]));
      }
      if (dlog) DBG(dslot, 10 * n + 7);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[s]);
      if (dlog) DBG(dslot, 10 * n + 8);
      if (q_ok) {
        const float inv = __fdividef(1.f, l);
        __nv_bfloat16* dst = p.o + (static_cast<long long>(b) * p.Tq + qrow) * p.ldo + h * kHd + half * 32;
        uint4 w[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          w[g].x = pack2(__uint_as_float(r0[8 * g + 0]) * inv, __uint_as_float(r0[8 * g + 1]) * inv);
          w[g].y = pack2(__uint_as_float(r0[8 * g + 2]) * inv, __uint_as_float(r0[8 * g + 3]) * inv);
          w[g].z = pack2(__uint_as_float(r0[8 * g + 4]) * inv, __uint_as_float(r0[8 * g + 5]) * inv);
          w[g].w = pack2(__uint_as_float(r0[8 * g + 6]) * inv, __uint_as_float(r0[8 * g + 7]) * inv);
        }
        if (p.o256) {  // this thread's 64 bytes of the row as two full 32-byte sectors
          st_global_256(dst, w[0], w[1]);
          st_global_256(dst + 16, w[2], w[3]);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(dst + g * 8) = w[g];
        }
        if (half == 0) p.lse[(static_cast<long long>(b) * p.H + h) * p.Tq + qrow] = m2 + __log2f(l);
      }
      if (dlog) DBG(dslot, 10 * n + 9);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// bf16 [batch][rows][cols] view, box = [1][box_rows][64 columns], 128B swizzle (the encoding of the GEMM operand maps).
int make_map(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long batch, long long ld,
             int box_rows) {
  const TmapKey key = make_tmap_key(ptr, cols, rows, batch, ld, rows * ld, 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (cached_tensor_map(map, key) != CUDA_SUCCESS)
    return md_set_error(MD_ERR_CUDA, "attention (tcgen05): cuTensorMapEncodeTiled failed");
  return 0;
}

int sm_count_cached() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

}  // namespace attn_tc
}  // namespace md

// diagnostics: copy the phase timeline of the last MD_ATTN_DEBUG=1 launch (8 slots x 512 clock64 stamps) to the host
extern "C" int md_attn_debug_dump(long long* out, int64_t n) {
  if (!out || n <= 0 || n > 8 * 512) return md::md_set_error(MD_ERR_INVALID, "md_attn_debug_dump: bad argument");
  cudaError_t e = cudaMemcpyFromSymbol(out, md::attn_tc::g_dbg, n * sizeof(long long));
  return e == cudaSuccess ? 0 : md::md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
}

extern "C" int md_attn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                              int64_t ldo, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t hd,
                              void* stream) {
  using namespace md;
  using namespace md::attn_tc;
  if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0) return B == 0 ? 0 : md_set_error(MD_ERR_INVALID, "md_attn_fwd_tc: bad sizes");
  if (hd != kHd || Tk > kMaxCols)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_attn_fwd_tc: needs head_dim 64 and Tk <= 256");
  if (!q || !k || !v || !o || !lse) return md_set_error(MD_ERR_INVALID, "md_attn_fwd_tc: null pointer");
  const uintptr_t align = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                          reinterpret_cast<uintptr_t>(o);
  if ((align & 15) != 0 || ((ldq | ldk | ldv | ldo) % 8) != 0)
    return md_set_error(MD_ERR_INVALID, "md_attn_fwd_tc: operands must be 16-byte aligned with pitches % 8 == 0");
  FwdParams p;
  static int dbg_env = -1;
  if (dbg_env < 0) { const char* e = getenv("MD_ATTN_DEBUG"); dbg_env = e ? atoi(e) : 0; }
  p.dbg = dbg_env;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.o256 = ((reinterpret_cast<uintptr_t>(o) & 31) == 0 && (ldo % 16) == 0) ? 1 : 0;
  p.lse = lse;
  p.H = static_cast<int>(H); p.Tq = static_cast<int>(Tq); p.Tk = static_cast<int>(Tk);
  p.n_pad = (p.Tk + 15) & ~15;
  p.packed = (Tq <= 64 && (H % 2) == 0 && 2 * p.n_pad <= kMaxCols) ? 1 : 0;
  p.q_blocks = p.packed ? 1 : static_cast<int>((Tq + kQ - 1) / kQ);
  p.head_tiles = p.packed ? p.H / 2 : p.H;
  p.tiles = static_cast<long long>(B) * p.head_tiles * p.q_blocks;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(hd));
  {  // shared-memory ring: stage = max(Q + K, P) + V for this shape's key count; as many stages (<= 3) as fit
    const int n_s = p.packed ? 2 * p.n_pad : p.n_pad;
    const int qk = kBytesQ + n_s * 128, pb = ((n_s + 63) / 64) * kBytesPAtom;
    p.v_off = ((qk > pb ? qk : pb) + 1023) & ~1023;
    p.stage_bytes = p.v_off + ((n_s * 128 + 1023) & ~1023);
    p.ns = kOffBar / p.stage_bytes;
    if (p.ns > kMaxStages) p.ns = kMaxStages;
    static int ns_env = -1;   // MD_ATTN_STAGES=2 restores the two-stage ring (A/B)
    if (ns_env < 0) { const char* e = getenv("MD_ATTN_STAGES"); ns_env = e ? atoi(e) : 0; }
    if (ns_env >= 2 && ns_env < p.ns) p.ns = ns_env;
    if (p.ns < 2) return md_set_error(MD_ERR_UNSUPPORTED, "md_attn_fwd_tc: shared-memory stage too large");
  }
  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_map(&tmQ, q, H * hd, Tq, B, ldq, p.packed ? 64 : kQ)) return rc;
  if (int rc = make_map(&tmK, k, H * hd, Tk, B, ldk, p.n_pad)) return rc;
  if (int rc = make_map(&tmV, v, H * hd, Tk, B, ldv, p.n_pad)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
    attr = true;
  }
  const long long sms = sm_count_cached();
  const unsigned grid = static_cast<unsigned>(p.tiles < sms ? p.tiles : sms);
  attn_fwd_tc_kernel<<<grid, kFwdThreads, kSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  return check_launch("md_attn_fwd_tc");
}

// ================================================================================================ backward
// dQ, dK, dV of the same attention (the reference gets them from autograd through F.scaled_dot_product_attention).
// delta = rowsum(dO * O) is derived here from o and dout (no scratch buffer).  Same envelope (head_dim 64, Tk <= 256).
//
// Persistent, warp-specialised.  A CTA loops over "iterations" = (sample, head [pair], 128-key block kb, 128-query block
// qb), kb outer / qb inner, so that dK / dV of the resident key block accumulate in TMEM over the query blocks and are
// read out once; dQ of a query block is produced per key block and summed across key blocks through global memory by
// the one thread that owns that row segment (deterministic).  Per iteration the five GEMMs are
//     S = Q K^T, dP = dO V^T            (M 128 queries, N = keys of the block, K 64)      -> TMEM
//     dV += P^T dO, dK += dS^T Q        (M 128 keys,    N 64, K 128 queries)              -> TMEM accumulators
//     dQ  = dS K                        (M 128 queries, N 64, K = keys of the block)      -> TMEM
// and the element-wise part  P = exp2(S * scale - lse),  dS = P * (dP - delta)  runs on two groups of four warps, each
// owning HALF of the block's key columns (chunk 0 / chunk 1: separate TMEM buffers and barriers), one query row per
// thread.  P and dS go to shared memory once, as bf16 [query][key] in 128B-swizzled atoms of 64 keys: that buffer is the
// K-major A operand of dQ and -- read through the MN-major descriptor -- the A operand of dV and dK.
// The issuer runs one iteration ahead with S / dP (double-buffered Q / dO and K / V stages), so the tensor core computes
// S / dP of iteration g+1 and the three gradient GEMMs of iteration g while the groups work on iteration g+1.
// Tq <= 64 and Tk <= 64 ("packed"): two heads share the 128-row tile; chunk c holds the keys of head c and rows of the
// other head write zeros there (block-diagonal P / dS).
namespace md {
namespace attn_tc {

constexpr int kBwdOffK = 0;                       // [2] x 16 KB : 128 keys x 64
constexpr int kBwdOffV = 32 * 1024;               // [2] x 16 KB
constexpr int kBwdOffQ = 64 * 1024;               // [2] x 16 KB : 128 queries x 64
constexpr int kBwdOffdO = 96 * 1024;              // [2] x 16 KB
constexpr int kBwdOffP = 128 * 1024;              // 32 KB : 2 atoms (128 queries x 64 keys)
constexpr int kBwdOffdS = 160 * 1024;             // 32 KB
constexpr int kBwdOffBar = 192 * 1024;
constexpr int kBwdOffStats = kBwdOffBar + 256;          // 2 x 128 x (lse, delta)
constexpr int kBwdOffOld = kBwdOffStats + 2048;         // 256 threads x 64 B: dQ partial of the previous key block
constexpr int kBwdSmemBytes = kBwdOffOld + 16384 + 1024;
constexpr int kTile = 16 * 1024;
constexpr int kBwdMaxKeys = 4096;
// TMEM columns
constexpr uint32_t kColS = 0, kColdP = 64, kColBuf = 128, kColdQ = 256, kColdK = 320, kColdV = 384;

struct BwdParams {
  int dbg, st256;
  const __nv_bfloat16* dout; long long lddo;
  const __nv_bfloat16* o; long long ldo;
  const float* lse;
  __nv_bfloat16* dq; long long lddq;
  __nv_bfloat16* dk; long long lddk;
  __nv_bfloat16* dv; long long lddv;
  int H, Tq, Tk, n_pad, packed, QB, KB, head_tiles, kv_box;
  long long items;
  float scale, scale_log2;
};

// Position of an iteration inside this CTA's sequence (sample / head tile outer, key block, query block inner), advanced
// incrementally: the only integer divisions happen once per (sample, head tile).
struct BwdPos {
  int b, h0, kb, qb, n_kb, n0, n1;
  uint32_t u;       // key-block phase counter (one per (item, kb))
  long long item;
  bool first, last; // first / last query block of the phase
};
__device__ __forceinline__ void bwd_pos_item(BwdPos& s, const BwdParams& p) {
  const int ht = static_cast<int>(s.item % p.head_tiles);
  s.b = static_cast<int>(s.item / p.head_tiles);
  s.h0 = p.packed ? 2 * ht : ht;
}
__device__ __forceinline__ void bwd_pos_block(BwdPos& s, const BwdParams& p) {
  if (p.packed) {
    s.n_kb = 2 * p.n_pad; s.n0 = p.n_pad; s.n1 = p.n_pad;
  } else {
    s.n_kb = min(128, p.n_pad - s.kb * 128);
    s.n0 = ((s.n_kb >> 1) + 15) & ~15;
    s.n1 = s.n_kb - s.n0;
  }
  s.first = s.qb == 0;
  s.last = s.qb == p.QB - 1;
}
__device__ __forceinline__ void bwd_pos_init(BwdPos& s, const BwdParams& p) {
  s.item = blockIdx.x; s.kb = 0; s.qb = 0; s.u = 0;
  bwd_pos_item(s, p);
  bwd_pos_block(s, p);
}
__device__ __forceinline__ void bwd_pos_next(BwdPos& s, const BwdParams& p) {
  if (++s.qb == p.QB) {
    s.qb = 0;
    ++s.u;
    if (++s.kb == p.KB) {
      s.kb = 0;
      s.item += gridDim.x;
      bwd_pos_item(s, p);
    }
  }
  bwd_pos_block(s, p);
}

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float (&f)[8]) {
  uint4 w;
  w.x = pack2(f[0], f[1]);
  w.y = pack2(f[2], f[3]);
  w.z = pack2(f[4], f[5]);
  w.w = pack2(f[6], f[7]);
  *reinterpret_cast<uint4*>(dst) = w;
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdOffBar);
  uint64_t* kv_full = bars;          // [2] TMA bytes of a K / V block
  uint64_t* kv_free = bars + 2;      // [2] every MMA of the phase that used this K / V stage has retired (commit)
  uint64_t* qdo_full = bars + 4;     // [2] TMA bytes of a Q / dO block
  uint64_t* mma_done = bars + 6;     // [2] gradient GEMMs of an iteration retired (commit): Q/dO stage, P/dS, dQ ready
  uint64_t* sdp_full = bars + 8;     // [2] S / dP chunk in TMEM (commit)
  uint64_t* sdp_free = bars + 10;    // [2] chunk consumed by its group (4 warps)
  uint64_t* pds_full = bars + 12;    // P / dS of the iteration in shared memory (8 warps)
  uint64_t* dq_free = bars + 13;     // dQ of the previous iteration read back (8 warps)
  uint64_t* acc_free = bars + 14;    // dK / dV of the previous phase read back (8 warps)
  uint64_t* st_full = bars + 15;     // [2] lse / delta of an iteration's 128 query rows in shared memory (2 warps)
  uint64_t* st_free = bars + 17;     // [2] ... read by the groups (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  float2* stats = reinterpret_cast<float2*>(smem + kBwdOffStats);  // [2][128] (lse in the log2 domain, delta)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t my_items = blockIdx.x < p.items ? static_cast<uint32_t>((p.items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const uint32_t total = my_items * p.KB * p.QB;
  const int Tq = p.Tq, Tk = p.Tk, H = p.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_free[i], 1);
      mbar_init(&qdo_full[i], 1);
      mbar_init(&mma_done[i], 1);
      mbar_init(&sdp_full[i], 1);
      mbar_init(&sdp_free[i], 4);
      mbar_init(&st_full[i], 2);
      mbar_init(&st_free[i], 8);
    }
    mbar_init(pds_full, 8);
    mbar_init(dq_free, 8);
    mbar_init(acc_free, 8);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      BwdPos it;
      bwd_pos_init(it, p);
      for (uint32_t g = 0; g < total; ++g, bwd_pos_next(it, p)) {
        if (it.first) {
          const int ks = it.u & 1;
          if (it.u >= 2) mbar_wait_sleep(&kv_free[ks], ((it.u >> 1) - 1) & 1);
          uint8_t* sK = smem + kBwdOffK + ks * kTile;
          uint8_t* sV = smem + kBwdOffV + ks * kTile;
          if (!p.packed) {
            mbar_expect_tx(&kv_full[ks], 2 * p.kv_box * 128);
            tma_load_3d(&tmK, &kv_full[ks], sK, it.h0 * kHd, it.kb * 128, it.b);
            tma_load_3d(&tmV, &kv_full[ks], sV, it.h0 * kHd, it.kb * 128, it.b);
          } else {
            mbar_expect_tx(&kv_full[ks], 4 * p.kv_box * 128);
            tma_load_3d(&tmK, &kv_full[ks], sK, it.h0 * kHd, 0, it.b);
            tma_load_3d(&tmK, &kv_full[ks], sK + p.n_pad * 128, (it.h0 + 1) * kHd, 0, it.b);
            tma_load_3d(&tmV, &kv_full[ks], sV, it.h0 * kHd, 0, it.b);
            tma_load_3d(&tmV, &kv_full[ks], sV + p.n_pad * 128, (it.h0 + 1) * kHd, 0, it.b);
          }
        }
        const int st = g & 1;
        if (g >= 2) mbar_wait_sleep(&mma_done[st], ((g >> 1) - 1) & 1);
        uint8_t* sQ = smem + kBwdOffQ + st * kTile;
        uint8_t* sdO = smem + kBwdOffdO + st * kTile;
        mbar_expect_tx(&qdo_full[st], 2 * kTile);
        if (!p.packed) {
          tma_load_3d(&tmQ, &qdo_full[st], sQ, it.h0 * kHd, it.qb * kQ, it.b);
          tma_load_3d(&tmdO, &qdo_full[st], sdO, it.h0 * kHd, it.qb * kQ, it.b);
        } else {
          tma_load_3d(&tmQ, &qdo_full[st], sQ, it.h0 * kHd, 0, it.b);
          tma_load_3d(&tmQ, &qdo_full[st], sQ + kTile / 2, (it.h0 + 1) * kHd, 0, it.b);
          tma_load_3d(&tmdO, &qdo_full[st], sdO, it.h0 * kHd, 0, it.b);
          tma_load_3d(&tmdO, &qdo_full[st], sdO + kTile / 2, (it.h0 + 1) * kHd, 0, it.b);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (total > 0) {  // whole warp walks the loop, one lane issues (see the forward kernel)
      const uint32_t idesc_mn = umma_idesc_bf16(kQ, kHd, true, true);    // dV, dK: A and B MN-major
      const uint32_t idesc_dq = umma_idesc_bf16(kQ, kHd, false, true);   // dQ: A K-major, B MN-major
      auto issue_sdp = [&](uint32_t g, const BwdPos& it) {
        const int st = g & 1, ks = it.u & 1;
        DBG(4, 8 * g);
        mbar_wait_sleep(&qdo_full[st], (g >> 1) & 1);
        if (it.first) mbar_wait_sleep(&kv_full[ks], (it.u >> 1) & 1);
        DBG(4, 8 * g + 1);
        const uint32_t aQ = smem_u32(smem + kBwdOffQ + st * kTile), adO = smem_u32(smem + kBwdOffdO + st * kTile);
        const uint32_t aK = smem_u32(smem + kBwdOffK + ks * kTile), aV = smem_u32(smem + kBwdOffV + ks * kTile);
        // both chunks' buffers must be free; then the four GEMMs (S0, dP0, S1, dP1) are issued interleaved, one UMMA_K
        // step each in turn: successive MMAs into one accumulator serialise on its ~100-cycle update latency, independent
        // accumulators pipeline
        if (g > 0) {
          mbar_wait_sleep(&sdp_free[0], (g - 1) & 1);
          DBG(4, 8 * g + 2);
          mbar_wait_sleep(&sdp_free[1], (g - 1) & 1);
        }
        DBG(4, 8 * g + 3);
        tc_fence_after();
        const uint64_t dq0 = umma_smem_desc(aQ, 16, 1024), ddo0 = umma_smem_desc(adO, 16, 1024);
        const uint64_t dk0 = umma_smem_desc(aK, 16, 1024), dv0 = umma_smem_desc(aV, 16, 1024);
        const uint64_t dk1 = dk0 + static_cast<uint64_t>(it.n0 * 128 >> 4), dv1 = dv0 + static_cast<uint64_t>(it.n0 * 128 >> 4);
        const uint32_t id0 = umma_idesc_bf16(kQ, it.n0, false, false);
        const uint32_t id1 = umma_idesc_bf16(kQ, it.n1 > 0 ? it.n1 : 16, false, false);
#pragma unroll
        for (int ks4 = 0; ks4 < kHd / 16; ++ks4) {
          const uint32_t acc = ks4 > 0 ? 1u : 0u;
          if (elect_one()) umma_bf16(tmem_base + kColS, dq0 + 2 * ks4, dk0 + 2 * ks4, id0, acc);
          if (elect_one()) umma_bf16(tmem_base + kColdP, ddo0 + 2 * ks4, dv0 + 2 * ks4, id0, acc);
          if (it.n1 > 0) {
            if (elect_one()) umma_bf16(tmem_base + kColBuf + kColS, dq0 + 2 * ks4, dk1 + 2 * ks4, id1, acc);
            if (elect_one()) umma_bf16(tmem_base + kColBuf + kColdP, ddo0 + 2 * ks4, dv1 + 2 * ks4, id1, acc);
          }
        }
        if (elect_one()) umma_commit(&sdp_full[0]);
        if (elect_one()) umma_commit(&sdp_full[1]);
      };
      auto issue_grad = [&](uint32_t g, const BwdPos& it) {
        const int st = g & 1, ks = it.u & 1;
        DBG(4, 8 * g + 4);
        mbar_wait_sleep(pds_full, g & 1);
        DBG(4, 8 * g + 5);
        if (g > 0) mbar_wait_sleep(dq_free, (g - 1) & 1);
        if (it.first && it.u > 0) mbar_wait_sleep(acc_free, (it.u - 1) & 1);
        DBG(4, 8 * g + 6);
        tc_fence_after();
        const uint32_t aQ = smem_u32(smem + kBwdOffQ + st * kTile), adO = smem_u32(smem + kBwdOffdO + st * kTile);
        const uint32_t aK = smem_u32(smem + kBwdOffK + ks * kTile);
        const uint32_t aP = smem_u32(smem + kBwdOffP), adS = smem_u32(smem + kBwdOffdS);
        // dV += P^T dO, dK += dS^T Q (reduction over the 128 queries of the block) and dQ = dS K (reduction over the keys
        // of the block: A = K-major dS atoms, B = the K tile read MN-major), issued interleaved -- three independent
        // accumulators pipeline, one accumulator alone serialises on its update latency.  Descriptors advance by adds.
        {
          const uint64_t db_do = umma_smem_desc(adO, 64 * 128, 1024), db_q = umma_smem_desc(aQ, 64 * 128, 1024);
          const uint64_t da_p = umma_smem_desc(aP, kTile, 1024), da_ds = umma_smem_desc(adS, kTile, 1024);
          const uint64_t dqa0 = umma_smem_desc(adS, 16, 1024), dqb0 = umma_smem_desc(aK, 64 * 128, 1024);
          const uint32_t acc0 = it.first ? 0u : 1u;
          const int qsteps = it.n_kb / 16;
#pragma unroll
          for (int kk = 0; kk < kQ / 16; ++kk) {
            if (elect_one()) umma_bf16(tmem_base + kColdV, da_p + 128 * kk, db_do + 128 * kk, idesc_mn, kk > 0 ? 1u : acc0);
            if (elect_one()) umma_bf16(tmem_base + kColdK, da_ds + 128 * kk, db_q + 128 * kk, idesc_mn, kk > 0 ? 1u : acc0);
            if (kk < qsteps) {
              const uint64_t da = dqa0 + static_cast<uint64_t>(kk >> 2) * (kTile >> 4) + 2 * (kk & 3);
              if (elect_one()) umma_bf16(tmem_base + kColdQ, da, dqb0 + 128 * kk, idesc_dq, kk > 0 ? 1u : 0u);
            }
          }
        }
        if (elect_one()) umma_commit(&mma_done[st]);
        if (it.last) if (elect_one()) umma_commit(&kv_free[ks]);
        DBG(4, 8 * g + 7);
      };
      BwdPos cur, nxt;
      bwd_pos_init(cur, p);
      nxt = cur;
      issue_sdp(0, cur);
      for (uint32_t g = 0; g < total; ++g) {
        if (g + 1 < total) {
          bwd_pos_next(nxt, p);
          issue_sdp(g + 1, nxt);
        }
        issue_grad(g, cur);
        cur = nxt;
      }
    }
  } else if (warp < 4) {
    // ================================ row statistics (warps 2, 3) ================================
    // lse (log2 domain) and delta = sum_d dO * O of the 128 query rows of iteration g, two rows per thread, one
    // iteration ahead of the groups; +inf lse -> P = 0 for padded query rows.
    BwdPos it;
    bwd_pos_init(it, p);
    for (uint32_t g = 0; g < total; ++g, bwd_pos_next(it, p)) {
      const int sb = g & 1;
      if (g >= 2) mbar_wait_sleep(&st_free[sb], ((g >> 1) - 1) & 1);
      for (int e = 0; e < 2; ++e) {
        const int row = (warp - 2) * 64 + e * 32 + lane;
        const int hi = (p.packed && row >= 64) ? 1 : 0;
        const int h = it.h0 + hi;
        const int qidx = p.packed ? (row & 63) : it.qb * kQ + row;
        float lrow = INFINITY, delta = 0.f;
        if (qidx < Tq) {
          lrow = p.lse[(static_cast<long long>(it.b) * H + h) * Tq + qidx];
          const uint4* po = reinterpret_cast<const uint4*>(p.o + (static_cast<long long>(it.b) * Tq + qidx) * p.ldo + h * kHd);
          const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (static_cast<long long>(it.b) * Tq + qidx) * p.lddo + h * kHd);
          uint4 a[kHd / 8], d[kHd / 8];
#pragma unroll
          for (int j = 0; j < kHd / 8; ++j) { a[j] = po[j]; d[j] = pd[j]; }
          float dl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < kHd / 8; ++j) {
            const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a[j]);
            const __nv_bfloat162* dh = reinterpret_cast<const __nv_bfloat162*>(&d[j]);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float2 fa = __bfloat1622float2(ah[q4]), fd = __bfloat1622float2(dh[q4]);
              dl[q4] = fmaf(fa.x, fd.x, dl[q4]);
              dl[q4] = fmaf(fa.y, fd.y, dl[q4]);
            }
          }
          delta = (dl[0] + dl[1]) + (dl[2] + dl[3]);
        }
        stats[sb * 128 + row] = make_float2(lrow, delta);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&st_full[sb]);
    }
  } else {
    // ================================ element-wise groups ================================
    const int c = (warp - 4) >> 2;                // chunk (half of the key block) this group owns
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;          // query row inside the tile == TMEM lane (== key row for dK / dV)
    const int hi = (p.packed && row >= 64) ? 1 : 0;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t sP = smem_u32(smem + kBwdOffP) + row * 128;     // this row of P / dS (shared-window addresses)
    const uint32_t sdS = smem_u32(smem + kBwdOffdS) + row * 128;
    const int sw = row & 7;
    const float sl2 = p.scale_log2, scale = p.scale;
    const bool active = !p.packed || hi == c;     // warp-uniform (rows 0-63 / 64-127 are whole warps)
    uint8_t* old_slot = smem + kBwdOffOld + (c * 128 + row) * 64;

    // read back dQ of iteration `it` (this thread: 32 columns of its row) and, after the last query block of a phase,
    // dK (group 0) or dV (group 1) of the key row this thread owns
    const int dr_log = (quarter == 0 && lane == 0), dr_slot = c ? 3 : 7;
    auto drain = [&](uint32_t g, const BwdPos& it) {
      if (dr_log) DBG(dr_slot, 8 * g);
      mbar_wait_sleep(&mma_done[g & 1], (g >> 1) & 1);
      if (dr_log) DBG(dr_slot, 8 * g + 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(trow + kColdQ + c * 32, r);
      tmem_ld_wait();
      if (dr_log) DBG(dr_slot, 8 * g + 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      const int qidx = p.packed ? (row & 63) : it.qb * kQ + row;
      if (qidx < Tq) {
        __nv_bfloat16* dst = p.dq + (static_cast<long long>(it.b) * Tq + qidx) * p.lddq + (it.h0 + hi) * kHd + c * 32;
        uint4 old[4];
        if (it.kb > 0) {  // sum over key blocks: the partial of the previous block was written by this very thread and has
                          // been prefetched (cp.async, issued before this iteration's element-wise work) into its slot
          asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 4; ++e) old[e] = *reinterpret_cast<const uint4*>(old_slot + 16 * e);
        }
        uint4 w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[8 * e + j]) * scale;
          if (it.kb > 0) {
            const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old[e]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 of = __bfloat1622float2(oh[j]);
              f[2 * j] += of.x;
              f[2 * j + 1] += of.y;
            }
          }
          w[e] = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
        }
        // 16-byte stores from this one-row-per-lane layout touch 32 half-used sectors per instruction and took ~1500 cycles
        // of every iteration (timeline); 256-bit stores write whole sectors with half the instructions
        if (p.st256) {
          st_global_256(dst, w[0], w[1]);
          st_global_256(dst + 16, w[2], w[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) *reinterpret_cast<uint4*>(dst + 8 * e) = w[e];
        }
      }
      if (dr_log) DBG(dr_slot, 8 * g + 3);
      if (it.last) {
        uint32_t r2[32];
        const uint32_t col = c == 0 ? kColdK : kColdV;
        tmem_ld_32x32(trow + col, r);
        tmem_ld_32x32(trow + col + 32, r2);
        tmem_ld_wait();
        if (dr_log) DBG(dr_slot, 8 * g + 4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);
        int key, kh;
        if (p.packed) { kh = row >= p.n_pad ? 1 : 0; key = row - kh * p.n_pad; if (row >= 2 * p.n_pad) key = Tk; }
        else { kh = 0; key = it.kb * 128 + row; }
        if (key < Tk) {
          const float mul = c == 0 ? scale : 1.0f;
          __nv_bfloat16* base = c == 0 ? p.dk + (static_cast<long long>(it.b) * Tk + key) * p.lddk
                                       : p.dv + (static_cast<long long>(it.b) * Tk + key) * p.lddv;
          __nv_bfloat16* dst = base + (it.h0 + kh) * kHd;
          uint4 w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t* rr = e < 4 ? r : r2;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(rr[8 * (e & 3) + j]) * mul;
            w[e] = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
          }
          if (p.st256) {
#pragma unroll
            for (int e = 0; e < 4; ++e) st_global_256(dst + 16 * e, w[2 * e], w[2 * e + 1]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) *reinterpret_cast<uint4*>(dst + 8 * e) = w[e];
          }
        }
      }
      if (dr_log) DBG(dr_slot, 8 * g + 5);
    };

    // the dQ partial that drain() of iteration `pv` adds to (written by this very thread one key block earlier) is fetched
    // into the thread's shared-memory slot ahead of time, under the element-wise work of the next iteration
    auto prefetch_old = [&](const BwdPos& pv) {
      if (pv.kb == 0) return;
      const int pq = p.packed ? (row & 63) : pv.qb * kQ + row;
      if (pq < Tq) {
        const __nv_bfloat16* src = p.dq + (static_cast<long long>(pv.b) * Tq + pq) * p.lddq + (pv.h0 + hi) * kHd + c * 32;
        const uint32_t sdst = smem_u32(old_slot);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + 16 * e), "l"(src + 8 * e) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    BwdPos it, prev;
    bwd_pos_init(it, p);
    prev = it;
    for (uint32_t g = 0; g < total; ++g) {
      const int n_c = c ? it.n1 : it.n0;
      const int colbase = c ? it.n0 : 0;
      // key index (inside its head) of column 0 of this chunk
      const int key0 = p.packed ? 0 : it.kb * 128 + colbase;

      const int dlog = (quarter == 0 && lane == 0), dslot = 5 + c;
      if (dlog) DBG(dslot, 8 * g);
      if (g > 0) prefetch_old(prev);
      mbar_wait_sleep(&st_full[g & 1], (g >> 1) & 1);
      if (dlog) DBG(dslot, 8 * g + 1);
      const float2 stv = stats[(g & 1) * 128 + row];
      __syncwarp();
      if (lane == 0) mbar_arrive(&st_free[g & 1]);
      const float lrow = stv.x, delta = stv.y;

      uint32_t Pp[32], dSp[32];  // 64 columns of P and dS, packed bf16x2
#pragma unroll
      for (int j = 0; j < 32; ++j) Pp[j] = dSp[j] = 0u;
      mbar_wait_sleep(&sdp_full[c], g & 1);
      if (dlog) DBG(dslot, 8 * g + 2);
      tc_fence_after();
      if (active && n_c > 0) {
        const uint32_t tb = trow + c * kColBuf;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf * 32 < n_c) {
            uint32_t rs[32], rp[32];
            tmem_ld_32x32(tb + kColS + hf * 32, rs);
            tmem_ld_32x32(tb + kColdP + hf * 32, rp);
            tmem_ld_wait();
            const int kbase = key0 + hf * 32;
            if (kbase + 32 <= Tk) {   // whole half inside the sequence: no masks
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(rs[j]), sl2, -lrow));
                const float p1 = ex2_approx(fmaf(__uint_as_float(rs[j + 1]), sl2, -lrow));
                Pp[hf * 16 + (j >> 1)] = pack2(p0, p1);
                dSp[hf * 16 + (j >> 1)] = pack2(p0 * (__uint_as_float(rp[j]) - delta), p1 * (__uint_as_float(rp[j + 1]) - delta));
              }
            } else {                  // zero-filled K / V rows beyond Tk (and stale TMEM beyond the MMA's N)
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(rs[j]), sl2, -lrow));
                float p1 = ex2_approx(fmaf(__uint_as_float(rs[j + 1]), sl2, -lrow));
                float d0 = p0 * (__uint_as_float(rp[j]) - delta);
                float d1 = p1 * (__uint_as_float(rp[j + 1]) - delta);
                if (kbase + j >= Tk) { p0 = 0.f; d0 = 0.f; }
                if (kbase + j + 1 >= Tk) { p1 = 0.f; d1 = 0.f; }
                Pp[hf * 16 + (j >> 1)] = pack2(p0, p1);
                dSp[hf * 16 + (j >> 1)] = pack2(d0, d1);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sdp_free[c]);
      if (dlog) DBG(dslot, 8 * g + 3);

      // P / dS of the previous iteration must have been consumed (its gradient GEMMs retired) before they are overwritten
      if (g > 0) mbar_wait_sleep(&mma_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
      if (dlog) DBG(dslot, 8 * g + 4);
      const int k8 = colbase >> 3;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        if (g8 * 8 < n_c) {
          const int key8 = k8 + g8;
          const int off = (key8 >> 3) * kTile + (((key8 & 7) ^ sw) << 4);
          st_shared_v4(sP + off, Pp[4 * g8], Pp[4 * g8 + 1], Pp[4 * g8 + 2], Pp[4 * g8 + 3]);
          st_shared_v4(sdS + off, dSp[4 * g8], dSp[4 * g8 + 1], dSp[4 * g8 + 2], dSp[4 * g8 + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (dlog) DBG(dslot, 8 * g + 5);
      if (g > 0) drain(g - 1, prev);
      if (dlog) DBG(dslot, 8 * g + 6);
      prev = it;
      bwd_pos_next(it, p);
    }
    if (total > 0) {
      prefetch_old(prev);
      drain(total - 1, prev);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace attn_tc
}  // namespace md

extern "C" int md_attn_bwd_tc(const void* dout, int64_t lddo, const void* q, int64_t ldq, const void* k, int64_t ldk,
                              const void* v, int64_t ldv, const void* o, int64_t ldo, const float* lse, void* dq,
                              int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t B, int64_t H,
                              int64_t Tq, int64_t Tk, int64_t hd, void* stream) {
  using namespace md;
  using namespace md::attn_tc;
  if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0) return B == 0 ? 0 : md_set_error(MD_ERR_INVALID, "md_attn_bwd_tc: bad sizes");
  // the backward walks the keys in resident blocks of 128 (dQ partials summed across blocks), so it has no 256-key limit
  if (hd != kHd || Tk > kBwdMaxKeys)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_attn_bwd_tc: needs head_dim 64 and Tk <= 4096");
  if (!dout || !q || !k || !v || !o || !lse || !dq || !dk || !dv)
    return md_set_error(MD_ERR_INVALID, "md_attn_bwd_tc: null pointer");
  const uintptr_t align = reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                          reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dq) |
                          reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv);
  if ((align & 15) != 0 || ((lddo | ldq | ldk | ldv | ldo | lddq | lddk | lddv) % 8) != 0)
    return md_set_error(MD_ERR_INVALID, "md_attn_bwd_tc: operands must be 16-byte aligned with pitches % 8 == 0");
  BwdParams p;
  static int dbg_env = -1;
  if (dbg_env < 0) { const char* e = getenv("MD_ATTN_DEBUG"); dbg_env = e ? atoi(e) : 0; }
  p.dbg = dbg_env;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout); p.lddo = lddo;
  p.o = reinterpret_cast<const __nv_bfloat16*>(o); p.ldo = ldo;
  p.lse = lse;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.lddq = lddq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.lddk = lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.lddv = lddv;
  p.st256 = (((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 31) == 0 &&
             ((lddq | lddk | lddv) % 16) == 0) ? 1 : 0;
  p.H = static_cast<int>(H); p.Tq = static_cast<int>(Tq); p.Tk = static_cast<int>(Tk);
  p.n_pad = (p.Tk + 15) & ~15;
  p.packed = (Tq <= 64 && (H % 2) == 0 && 2 * p.n_pad <= 128) ? 1 : 0;
  p.QB = p.packed ? 1 : static_cast<int>((Tq + kQ - 1) / kQ);
  p.KB = p.packed ? 1 : (p.n_pad + 127) / 128;
  p.head_tiles = p.packed ? p.H / 2 : p.H;
  p.kv_box = p.packed ? p.n_pad : (p.KB > 1 ? 128 : p.n_pad);
  p.items = static_cast<long long>(B) * p.head_tiles;
  p.scale = 1.f / sqrtf(static_cast<float>(hd));
  p.scale_log2 = 1.4426950408889634f * p.scale;
  CUtensorMap tmQ, tmdO, tmK, tmV;
  if (int rc = make_map(&tmQ, q, H * hd, Tq, B, ldq, p.packed ? 64 : kQ)) return rc;
  if (int rc = make_map(&tmdO, dout, H * hd, Tq, B, lddo, p.packed ? 64 : kQ)) return rc;
  if (int rc = make_map(&tmK, k, H * hd, Tk, B, ldk, p.kv_box)) return rc;
  if (int rc = make_map(&tmV, v, H * hd, Tk, B, ldv, p.kv_box)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes);
    if (e != cudaSuccess) return md_set_error(MD_ERR_CUDA, cudaGetErrorString(e));
    attr = true;
  }
  const long long sms = sm_count_cached();
  const unsigned grid = static_cast<unsigned>(p.items < sms ? p.items : sms);
  attn_bwd_tc_kernel<<<grid, kThreads, kBwdSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmdO, tmK, tmV, p);
  return check_launch("md_attn_bwd_tc");
}
