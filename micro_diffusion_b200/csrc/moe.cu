// Expert-choice MoE routing (FeedForwardECMoe.forward, dit.py:126-143) around the grouped expert
// GEMMs: gate + softmax, per-(sample, expert) top-k over tokens, dispatch gather, weighted combine
// fused with the gated residual, and the matching backward pieces.  The reference materialises a
// one-hot (n,e,k,t) tensor and contracts with it (dit.py:133-134,140); here dispatch and combine are
// index gathers (deterministic: an inverse slot table replaces scatter-add atomics).
#include "act.cuh"

namespace md {

constexpr int kMaxE = 16;
constexpr int kChunk = 4;  // 16-byte chunks a lane keeps in flight (D <= 1024 in one slab)

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __low2float(h[e]);
    f[2 * e + 1] = __high2float(h[e]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return r;
}

// ---------------------------------------------------------------------------------------- gate fwd
// A warp owns kRowsPerWarp consecutive rows; gate weights staged in shared memory (E*D fp32 <= 64 KB).  The weights
// of a column chunk are read from shared memory once and used for all the warp's rows: with one row per warp the
// kernel was shared-memory-bandwidth bound (16 LDS.128 per 64 FMA), 5x off the HBM time of reading x once.
constexpr int kRowsPerWarp = 4;
constexpr int kGChunk = 2;  // 16-byte chunks per row a lane keeps in flight (x 4 rows)
template <int ME, typename AT>
__global__ void __launch_bounds__(256, ME <= 8 ? 2 : 1)
moe_gate_fwd_kernel(const AT* __restrict__ x, const float* __restrict__ wg, float* __restrict__ probs, long long rows,
                    int D, int E) {
  extern __shared__ float swg[];  // [E][D]
  {  // 16-byte loads, several in flight per thread: a scalar copy loop exposed one L2 round trip per element
    const int n4 = (E * D) >> 2;
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
      reinterpret_cast<float4*>(swg)[i] = reinterpret_cast<const float4*>(wg)[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  for (long long row0 = (1LL * blockIdx.x * 8 + warp) * kRowsPerWarp; row0 < rows;
       row0 += 1LL * gridDim.x * 8 * kRowsPerWarp) {
    float acc[kRowsPerWarp][ME];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
      for (int e = 0; e < ME; ++e) acc[r][e] = 0.f;
    for (int base = 0; base < nvec; base += 32 * kGChunk) {
      V8<AT> raw[kGChunk][kRowsPerWarp];
#pragma unroll
      for (int j = 0; j < kGChunk; ++j) {  // all loads of this slab first
        const int i = base + lane + 32 * j;
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r)
          raw[j][r] = (i < nvec && row0 + r < rows) ? ldv8(x + (row0 + r) * D + 8 * i) : zerov8<AT>();
      }
#pragma unroll
      for (int j = 0; j < kGChunk; ++j) {
        const int i = base + lane + 32 * j;
        if (i < nvec) {
          float xv[kRowsPerWarp][8];
#pragma unroll
          for (int r = 0; r < kRowsPerWarp; ++r) unpackv8(raw[j][r], xv[r]);
#pragma unroll
          for (int e = 0; e < ME; ++e) {
            if (e < E) {
              const float4 w0 = *reinterpret_cast<const float4*>(swg + e * D + 8 * i);
              const float4 w1 = *reinterpret_cast<const float4*>(swg + e * D + 8 * i + 4);
#pragma unroll
              for (int r = 0; r < kRowsPerWarp; ++r)
                acc[r][e] += xv[r][0] * w0.x + xv[r][1] * w0.y + xv[r][2] * w0.z + xv[r][3] * w0.w + xv[r][4] * w1.x +
                             xv[r][5] * w1.y + xv[r][6] * w1.z + xv[r][7] * w1.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < ME; ++e) {
        if (e < E) {
          acc[r][e] = warp_sum(acc[r][e]);
          mx = fmaxf(mx, acc[r][e]);
        }
      }
      float den = 0.f;
#pragma unroll
      for (int e = 0; e < ME; ++e) {
        if (e < E) {
          acc[r][e] = expf(acc[r][e] - mx);
          den += acc[r][e];
        }
      }
      if (lane == 0 && row0 + r < rows) {
#pragma unroll
        for (int e = 0; e < ME; ++e)
          if (e < E) probs[(row0 + r) * E + e] = acc[r][e] / den;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------- top-k
// block per (sample, expert): bitonic sort of (prob, token) descending; ties broken by lower token id.
__global__ void __launch_bounds__(1024)
moe_topk_kernel(const float* __restrict__ probs, int32_t* __restrict__ idx, float* __restrict__ gval,
                int32_t* __restrict__ inv, int T, int E, int k, int n) {
  extern __shared__ unsigned long long keys[];
  const int e = blockIdx.x % E;
  const long long b = blockIdx.x / E;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned long long key = ~0ULL;  // padding sorts last
    if (i < T) {
      // probs are in [0,1]: the raw bit pattern is monotone.  Descending by prob == ascending by ~bits.
      const unsigned int u = ~__float_as_uint(probs[(b * T + i) * E + e]);
      key = (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned int>(i);
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) {
          keys[lo] = c;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    const int tok = static_cast<int>(keys[j] & 0xffffffffu);
    if (j < k) {
      idx[(b * E + e) * k + j] = tok;
      gval[(b * E + e) * k + j] = __uint_as_float(~static_cast<unsigned int>(keys[j] >> 32));
    }
    inv[(b * T + tok) * E + e] = j < k ? j : -1;
  }
}

// ------------------------------------------------------------------------------------------ gather
// warp per slot: xin[e][b*k + j][:] = x[b*T + idx[b,e,j]][:]
template <typename AT>
__global__ void __launch_bounds__(256)
moe_gather_kernel(const AT* __restrict__ x, const int32_t* __restrict__ idx, AT* __restrict__ xin, int B, int T, int E,
                  int k, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long slots = 1LL * B * E * k;
  const int nvec = D >> 3;
  for (long long s = 1LL * blockIdx.x * 8 + warp; s < slots; s += 1LL * gridDim.x * 8) {
    const int j = static_cast<int>(s % k);
    const int e = static_cast<int>((s / k) % E);
    const long long b = s / (1LL * k * E);
    const int tok = idx[s];  // idx is [B,E,k] == s ordering
    const AT* src = x + (b * T + tok) * D;
    AT* dst = xin + ((1LL * e * B + b) * k + j) * D;
    for (int i = lane; i < nvec; i += 32) {
      float v[8];
      ld8(src + 8 * i, v);
      st8(dst + 8 * i, v);
    }
  }
}

// ------------------------------------------------------------------------------------- combine fwd
// warp per token: ymoe = sum_e g*h2[slot]; xout = xres + gate*ymoe
template <int ME, typename AT>
__global__ void __launch_bounds__(256)
moe_combine_fwd_kernel(const AT* __restrict__ h2, const float* __restrict__ gval, const int32_t* __restrict__ inv,
                       const float* __restrict__ xres, const float* __restrict__ gate, long long ldmod,
                       float* __restrict__ xout, AT* __restrict__ ymoe, int B, int T, int E, int k, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = 1LL * B * T;
  const int nvec = D >> 3;
  for (long long row = 1LL * blockIdx.x * 8 + warp; row < rows; row += 1LL * gridDim.x * 8) {
    const long long b = row / T;
    int slot[ME];
    float g[ME];
#pragma unroll
    for (int e = 0; e < ME; ++e) {
      slot[e] = -1;
      g[e] = 0.f;
      if (e < E) {
        slot[e] = inv[row * E + e];
        if (slot[e] >= 0) g[e] = gval[(b * E + e) * k + slot[e]];
      }
    }
    const float* gt = gate ? gate + b * ldmod : nullptr;
    for (int i = lane; i < nvec; i += 32) {
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      V8<AT> raw[ME];
#pragma unroll
      for (int e = 0; e < ME; ++e)  // every selected expert's row chunk in flight before the first use
        if (e < E && slot[e] >= 0) raw[e] = ldv8(h2 + ((1LL * e * B + b) * k + slot[e]) * D + 8 * i);
#pragma unroll
      for (int e = 0; e < ME; ++e) {
        if (e < E && slot[e] >= 0) {
          float hv[8];
          unpackv8(raw[e], hv);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] += g[e] * hv[q];
        }
      }
      if (ymoe) st8(ymoe + row * D + 8 * i, acc);
      if (xout) {
#pragma unroll
        for (int q = 0; q < 8; q += 4) {
          float4 r = *reinterpret_cast<const float4*>(xres + row * D + 8 * i + q);
          float4 gg = make_float4(1.f, 1.f, 1.f, 1.f);
          if (gt) gg = *reinterpret_cast<const float4*>(gt + 8 * i + q);
          r.x += gg.x * acc[q]; r.y += gg.y * acc[q + 1]; r.z += gg.z * acc[q + 2]; r.w += gg.w * acc[q + 3];
          *reinterpret_cast<float4*>(xout + row * D + 8 * i + q) = r;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------- combine bwd
// warp per slot: dh2 = g * dy[token]; dgval = <h2, dy[token]>
template <typename AT>
__global__ void __launch_bounds__(256)
moe_combine_bwd_kernel(const AT* __restrict__ dy, const AT* __restrict__ h2, const float* __restrict__ gval,
                       const int32_t* __restrict__ idx, AT* __restrict__ dh2, float* __restrict__ dgval, int B, int T,
                       int E, int k, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long slots = 1LL * B * E * k;
  const int nvec = D >> 3;
  for (long long s = 1LL * blockIdx.x * 8 + warp; s < slots; s += 1LL * gridDim.x * 8) {
    const int j = static_cast<int>(s % k);
    const int e = static_cast<int>((s / k) % E);
    const long long b = s / (1LL * k * E);
    const int tok = idx[s];
    const float g = gval[s];
    const AT* dyr = dy + (b * T + tok) * D;
    const long long hrow = ((1LL * e * B + b) * k + j) * D;
    float dot = 0.f;
    for (int i = lane; i < nvec; i += 32) {
      float dv[8], hv[8], o[8];
      ld8(dyr + 8 * i, dv);
      ld8(h2 + hrow + 8 * i, hv);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        dot += dv[q] * hv[q];
        o[q] = g * dv[q];
      }
      st8(dh2 + hrow + 8 * i, o);
    }
    dot = warp_sum(dot);
    if (lane == 0) dgval[s] = dot;
  }
}

// ------------------------------------------------------------------------------------------ dx bwd
// Half-warp per token: dscores (softmax backward of the selected gate values) and
// dx = sum_e dxin[slot] + dscores . Wg.  Lanes l and l + 16 work on the same columns of two consecutive tokens, so
// their gate-weight reads hit the same shared-memory addresses (one wavefront pair instead of four per LDS.128):
// with a full warp per token the weight reads, not HBM, set the pace.
template <int ME, typename AT>
__global__ void __launch_bounds__(256)
moe_dx_bwd_kernel(const AT* __restrict__ dxin, const int32_t* __restrict__ inv, const float* __restrict__ dgval,
                  const float* __restrict__ probs, const float* __restrict__ wg, float* __restrict__ dscores,
                  AT* __restrict__ dx, int B, int T, int E, int k, int D) {
  extern __shared__ float swg[];  // [E][D]
  {  // 16-byte loads, several in flight per thread: a scalar copy loop exposed one L2 round trip per element
    const int n4 = (E * D) >> 2;
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
      reinterpret_cast<float4*>(swg)[i] = reinterpret_cast<const float4*>(wg)[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hl = lane & 15;
  const long long rows = 1LL * B * T;
  const int nvec = D >> 3;
  for (long long row = 2 * (1LL * blockIdx.x * 8 + warp) + (lane >> 4); row < rows; row += 2LL * gridDim.x * 8) {
    const long long b = row / T;
    int slot[ME];
    float ds[ME];
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < ME; ++e) {
      slot[e] = -1;
      ds[e] = 0.f;
      if (e < E) {
        slot[e] = inv[row * E + e];
        const float pe = probs[row * E + e];
        const float dp = slot[e] >= 0 ? dgval[(b * E + e) * k + slot[e]] : 0.f;
        ds[e] = dp;
        dot += pe * dp;
      }
    }
#pragma unroll
    for (int e = 0; e < ME; ++e)
      if (e < E) ds[e] = probs[row * E + e] * (ds[e] - dot);
    if (hl == 0) {
#pragma unroll
      for (int e = 0; e < ME; ++e)
        if (e < E) dscores[row * E + e] = ds[e];
    }
    for (int i = hl; i < nvec; i += 16) {
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      V8<AT> raw[ME];
#pragma unroll
      for (int e = 0; e < ME; ++e)
        if (e < E && slot[e] >= 0) raw[e] = ldv8(dxin + ((1LL * e * B + b) * k + slot[e]) * D + 8 * i);
#pragma unroll
      for (int e = 0; e < ME; ++e) {
        if (e < E) {
          const float4 w0 = *reinterpret_cast<const float4*>(swg + e * D + 8 * i);
          const float4 w1 = *reinterpret_cast<const float4*>(swg + e * D + 8 * i + 4);
          acc[0] += ds[e] * w0.x; acc[1] += ds[e] * w0.y; acc[2] += ds[e] * w0.z; acc[3] += ds[e] * w0.w;
          acc[4] += ds[e] * w1.x; acc[5] += ds[e] * w1.y; acc[6] += ds[e] * w1.z; acc[7] += ds[e] * w1.w;
          if (slot[e] >= 0) {
            float dv[8];
            unpackv8(raw[e], dv);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += dv[q];
          }
        }
      }
      st8(dx + row * D + 8 * i, acc);
    }
  }
}

// -------------------------------------------------------------------------------------- gate wgrad
// dwg[e][c] += sum_r ds[r][e] * x[r][c].  Persistent blocks of 256 threads walk 64-row slabs; a thread owns the
// column pairs (2t, 2t+1) + 512 j and keeps their E partial sums in registers over ALL its slabs, so the atomics
// are paid once per block (the first version paid them per 128 rows and read ds with one LDS per FMA).
constexpr int kWgSlab = 64;
template <int ME, int NJ, typename AT>
__global__ void __launch_bounds__(256)
moe_gate_wgrad_kernel(const float* __restrict__ dscores, const AT* __restrict__ x, float* __restrict__ dwg,
                      long long rows, int D, int E, float* __restrict__ ws) {
  __shared__ __align__(16) float sds[kWgSlab * ME];
  const int c0 = 2 * threadIdx.x;
  float acc[NJ][2][ME];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int e = 0; e < ME; ++e) acc[j][0][e] = acc[j][1][e] = 0.f;
  const long long slabs = (rows + kWgSlab - 1) / kWgSlab;
  for (long long sl = blockIdx.x; sl < slabs; sl += gridDim.x) {
    const long long r0 = sl * kWgSlab;
    const int nr = static_cast<int>(min(static_cast<long long>(kWgSlab), rows - r0));
    __syncthreads();
    for (int i = threadIdx.x; i < kWgSlab * ME; i += blockDim.x) {
      const int r = i / ME, e = i % ME;
      sds[i] = (r < nr && e < E) ? dscores[(r0 + r) * E + e] : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < nr; r += 8) {
      float xv[8][NJ][2];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int c = c0 + 512 * j;
          const bool ok = (r + u < nr) && (c < D);
          const float2 t = ok ? ld2a(x + (r0 + r + u) * D + c) : make_float2(0.f, 0.f);
          xv[u][j][0] = t.x;
          xv[u][j][1] = t.y;
        }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float d[ME];
#pragma unroll
        for (int e = 0; e < ME; e += 4) {
          const float4 t = *reinterpret_cast<const float4*>(sds + (r + u) * ME + e);  // rows >= nr hold zeros
          d[e] = t.x; d[e + 1] = t.y; d[e + 2] = t.z; d[e + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int e = 0; e < ME; ++e) {
            acc[j][0][e] += d[e] * xv[u][j][0];
            acc[j][1][e] += d[e] * xv[u][j][1];
          }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = c0 + 512 * j;
    if (c < D) {
#pragma unroll
      for (int e = 0; e < ME; ++e)
        if (e < E) {
          if (ws != nullptr) {  // deterministic mode: this block's partial, reduced over blocks in a fixed order afterwards
            ws[(1LL * blockIdx.x * E + e) * D + c] = acc[j][0][e];
            ws[(1LL * blockIdx.x * E + e) * D + c + 1] = acc[j][1][e];
          } else {
            atomicAdd(dwg + e * D + c, acc[j][0][e]);
            atomicAdd(dwg + e * D + c + 1, acc[j][1][e]);
          }
        }
    }
  }
}

// kernels that stage the gate weights in shared memory: at most two resident blocks per SM (register budget), so a grid of
// 2 x SMs pays the fill once per block
static int staged_grid(long long warps_needed) {
  long long blocks = (warps_needed + 7) / 8;
  if (blocks > 148LL * 2) blocks = 148LL * 2;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}
static int warp_grid(long long warps_needed) {
  long long blocks = (warps_needed + 7) / 8;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}
static int check_moe(const char* what, long long D, long long E) {
  if (E < 1 || E > kMaxE || D % 8 != 0 || D <= 0) {
    char buf[128];
    snprintf(buf, sizeof(buf), "%s: need 1 <= E <= %d and D %% 8 == 0 (E=%lld D=%lld)", what, kMaxE, E, D);
    return md_set_error(MD_ERR_UNSUPPORTED, buf);
  }
  return 0;
}

}  // namespace md

using namespace md;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int md_moe_gate_fwd(const void* x, const float* wg, float* probs, int64_t rows, int64_t D, int64_t E,
                               int prec, void* stream) {
  if (int rc = check_moe("md_moe_gate_fwd", D, E)) return rc;
  if (rows == 0) return 0;
  if (!x || !wg || !probs) return md_set_error(MD_ERR_INVALID, "md_moe_gate_fwd: null pointer");
  const size_t smem = E * D * sizeof(float);
  if (smem > 200 * 1024) return md_set_error(MD_ERR_UNSUPPORTED, "md_moe_gate_fwd: E*D too large for shared memory");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(moe_gate_fwd_kernel<8, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(moe_gate_fwd_kernel<8, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(moe_gate_fwd_kernel<16, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(moe_gate_fwd_kernel<16, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  const int grid = staged_grid((rows + kRowsPerWarp - 1) / kRowsPerWarp);
  if (E <= 8)
    MD_WITH_ACT(prec, moe_gate_fwd_kernel<8, AT><<<grid, 256, smem, ST(stream)>>>(CAP(AT, x), wg, probs, rows, (int)D, (int)E));
  else
    MD_WITH_ACT(prec, moe_gate_fwd_kernel<16, AT><<<grid, 256, smem, ST(stream)>>>(CAP(AT, x), wg, probs, rows, (int)D, (int)E));
  return check_launch("md_moe_gate_fwd");
}

extern "C" int md_moe_topk(const float* probs, int32_t* idx, float* gval, int32_t* inv, int64_t B, int64_t T, int64_t E,
                           int64_t k, void* stream) {
  if (B == 0) return 0;
  if (!probs || !idx || !gval || !inv) return md_set_error(MD_ERR_INVALID, "md_moe_topk: null pointer");
  if (T > 4096 || T < 1 || k > T || k < 0 || E < 1 || E > kMaxE)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_moe_topk: need 1 <= T <= 4096, k <= T, E <= 16");
  int n = 2;
  while (n < T) n <<= 1;
  const int threads = n / 2 < 32 ? 32 : (n / 2 > 1024 ? 1024 : n / 2);
  moe_topk_kernel<<<(unsigned)(B * E), threads, n * sizeof(unsigned long long), ST(stream)>>>(probs, idx, gval, inv,
                                                                                             (int)T, (int)E, (int)k, n);
  return check_launch("md_moe_topk");
}

extern "C" int md_moe_gather(const void* x, const int32_t* idx, void* xin, int64_t B, int64_t T, int64_t E, int64_t k,
                             int64_t D, int prec, void* stream) {
  if (int rc = check_moe("md_moe_gather", D, E)) return rc;
  if (B * k == 0) return 0;
  if (!x || !idx || !xin) return md_set_error(MD_ERR_INVALID, "md_moe_gather: null pointer");
  MD_WITH_ACT(prec, moe_gather_kernel<AT><<<warp_grid(B * E * k), 256, 0, ST(stream)>>>(CAP(AT, x), idx, AP(AT, xin), (int)B,
                                                                                        (int)T, (int)E, (int)k, (int)D));
  return check_launch("md_moe_gather");
}

extern "C" int md_moe_combine_fwd(const void* h2, const float* gval, const int32_t* inv, const float* xres,
                                  const float* gate, int64_t ldmod, float* xout, void* ymoe, int64_t B, int64_t T,
                                  int64_t E, int64_t k, int64_t D, int prec, void* stream) {
  if (int rc = check_moe("md_moe_combine_fwd", D, E)) return rc;
  if (B * T == 0) return 0;
  if (!h2 || !gval || !inv || (xout && !xres)) return md_set_error(MD_ERR_INVALID, "md_moe_combine_fwd: null pointer");
  if (E <= 8)
    MD_WITH_ACT(prec, moe_combine_fwd_kernel<8, AT><<<warp_grid(B * T), 256, 0, ST(stream)>>>(
                          CAP(AT, h2), gval, inv, xres, gate, ldmod, xout, AP(AT, ymoe), (int)B, (int)T, (int)E, (int)k, (int)D));
  else
    MD_WITH_ACT(prec, moe_combine_fwd_kernel<16, AT><<<warp_grid(B * T), 256, 0, ST(stream)>>>(
                          CAP(AT, h2), gval, inv, xres, gate, ldmod, xout, AP(AT, ymoe), (int)B, (int)T, (int)E, (int)k, (int)D));
  return check_launch("md_moe_combine_fwd");
}

extern "C" int md_moe_combine_bwd(const void* dy, const void* h2, const float* gval, const int32_t* idx, void* dh2,
                                  float* dgval, int64_t B, int64_t T, int64_t E, int64_t k, int64_t D, int prec,
                                  void* stream) {
  if (int rc = check_moe("md_moe_combine_bwd", D, E)) return rc;
  if (B * k == 0) return 0;
  if (!dy || !h2 || !gval || !idx || !dh2 || !dgval)
    return md_set_error(MD_ERR_INVALID, "md_moe_combine_bwd: null pointer");
  MD_WITH_ACT(prec, moe_combine_bwd_kernel<AT><<<warp_grid(B * E * k), 256, 0, ST(stream)>>>(
                        CAP(AT, dy), CAP(AT, h2), gval, idx, AP(AT, dh2), dgval, (int)B, (int)T, (int)E, (int)k, (int)D));
  return check_launch("md_moe_combine_bwd");
}

extern "C" int md_moe_dx_bwd(const void* dxin, const int32_t* inv, const float* dgval, const float* probs,
                             const float* wg, float* dscores, void* dx, int64_t B, int64_t T, int64_t E, int64_t k,
                             int64_t D, int prec, void* stream) {
  if (int rc = check_moe("md_moe_dx_bwd", D, E)) return rc;
  if (B * T == 0) return 0;
  if (!dxin || !inv || !dgval || !probs || !wg || !dscores || !dx)
    return md_set_error(MD_ERR_INVALID, "md_moe_dx_bwd: null pointer");
  const size_t smem = E * D * sizeof(float);
  if (smem > 200 * 1024) return md_set_error(MD_ERR_UNSUPPORTED, "md_moe_dx_bwd: E*D too large for shared memory");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(moe_dx_bwd_kernel<8, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(moe_dx_bwd_kernel<16, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(moe_dx_bwd_kernel<8, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(moe_dx_bwd_kernel<16, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  if (E <= 8)
    MD_WITH_ACT(prec, moe_dx_bwd_kernel<8, AT><<<staged_grid((B * T + 1) / 2), 256, smem, ST(stream)>>>(
                          CAP(AT, dxin), inv, dgval, probs, wg, dscores, AP(AT, dx), (int)B, (int)T, (int)E, (int)k, (int)D));
  else
    MD_WITH_ACT(prec, moe_dx_bwd_kernel<16, AT><<<staged_grid((B * T + 1) / 2), 256, smem, ST(stream)>>>(
                          CAP(AT, dxin), inv, dgval, probs, wg, dscores, AP(AT, dx), (int)B, (int)T, (int)E, (int)k, (int)D));
  return check_launch("md_moe_dx_bwd");
}

extern "C" int md_moe_gate_wgrad(const float* dscores, const void* x, float* dwg, int64_t rows, int64_t D, int64_t E,
                                 int prec, void* stream) {
  if (int rc = check_moe("md_moe_gate_wgrad", D, E)) return rc;
  if (rows == 0) return 0;
  if (!dscores || !x || !dwg) return md_set_error(MD_ERR_INVALID, "md_moe_gate_wgrad: null pointer");
  if (D > 2048) return md_set_error(MD_ERR_UNSUPPORTED, "md_moe_gate_wgrad: D > 2048");
  const long long slabs = (rows + kWgSlab - 1) / kWgSlab;
  const unsigned grid = (unsigned)(slabs < 148 * 2 ? slabs : 148 * 2);
  float* ws = nullptr;
  if (det_enabled()) {
    ws = det_workspace(static_cast<size_t>(grid) * E * D * sizeof(float));
    if (ws == nullptr) return md_set_error(MD_ERR_INVALID, "md_moe_gate_wgrad: deterministic workspace too small");
  }
#define WGRAD(ME, NJ) \
  MD_WITH_ACT(prec, moe_gate_wgrad_kernel<ME, NJ, AT><<<grid, 256, 0, ST(stream)>>>(dscores, CAP(AT, x), dwg, rows, (int)D, (int)E, ws))
  if (E <= 8) {
    if (D <= 512) WGRAD(8, 1);
    else if (D <= 1024) WGRAD(8, 2);
    else WGRAD(8, 4);
  } else {
    if (D <= 512) WGRAD(16, 1);
    else if (D <= 1024) WGRAD(16, 2);
    else WGRAD(16, 4);
  }
#undef WGRAD
  if (int rc = check_launch("md_moe_gate_wgrad")) return rc;
  return ws ? det_reduce(ws, dwg, grid, E * D, 1, ST(stream)) : 0;
}
