// LayerNorm(+adaLN modulate) forward/backward, QK row-norm, gated-residual backward.
// All HBM-bound: one warp per row, 16-byte vector loads, warp-shuffle reductions, fp32 statistics.
// Reference semantics: create_norm (utils.py:71-78), modulate (utils.py:28-30), DiTBlock.forward
// (dit.py:232-239), ln_q/ln_k over the full hidden width (utils.py:183-186, 122-125).
#include "act.cuh"

namespace md {

constexpr int kMaxVec = 16;  // float4 groups per lane  -> D <= 2048

__device__ __forceinline__ float4 load4(const void* base, bool bf16, long long elem_off) {
  if (bf16) {
    const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
}
__device__ __forceinline__ void store4_bf16(void* base, long long elem_off, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = raw;
}

// All row kernels issue every global load of a row BEFORE the first dependent use (separate load / compute
// loops, compile-time trip counts, no per-load predicates on the fast path): one warp then has VEC x 16-byte
// (or more) requests in flight instead of one, which is what these latency-bound kernels need to approach the
// HBM roofline.  EXACT: D == 128 * VEC (the model widths 512 / 768 / 1024); otherwise a predicated generic path.

template <bool XBF>
__device__ __forceinline__ float4 ld4(const void* base, long long elem_off) {
  if constexpr (XBF) {
    const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  } else {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
  }
}
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ------------------------------------------------------------------------------------------ ln_fwd
// grid (ceil(T / rpb), samples): a block works on rows of ONE sample, so the per-sample vectors -- gamma * (1 + scale),
// shift and the residual gate -- are staged in shared memory once per block.  (With a flat row grid every row re-read up
// to 24 float4 of them from L2 AFTER its statistics were known: a second exposed latency per row, ~25 % of the kernel.)
template <int VEC, bool EXACT, bool XBF, typename AT>
__global__ void __launch_bounds__(128)
ln_fwd_kernel(const void* __restrict__ x, const int32_t* __restrict__ src_rows, const AT* __restrict__ yadd,
              const float* __restrict__ gadd, float* __restrict__ xnew, const float* __restrict__ gamma,
              const float* __restrict__ shift, const float* __restrict__ scale, long long ldmod, long long T,
              AT* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int D,
              float eps, int rpb) {
  extern __shared__ float sp[];   // [3][D]: gamma * (1 + scale) | shift | gate of the pending residual
  float* sgw = sp;
  float* ssh = sp + D;
  float* sga = sp + 2 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  const long long smp = blockIdx.y;
  const long long t0 = 1LL * blockIdx.x * rpb;
  const long long t1 = min(T, t0 + rpb);
  {
    const float* sh = shift ? shift + smp * ldmod : nullptr;
    const float* sc = scale ? scale + smp * ldmod : nullptr;
    const float* gt = gadd ? gadd + smp * ldmod : nullptr;
    for (int c = threadIdx.x; c < D; c += 128) {
      sgw[c] = (gamma ? gamma[c] : 1.f) * (sc ? 1.f + sc[c] : 1.f);
      ssh[c] = sh ? sh[c] : 0.f;
      sga[c] = gt ? gt[c] : 1.f;
    }
  }
  __syncthreads();
  for (long long t = t0 + warp; t < t1; t += 4) {
    const long long row = smp * T + t;
    const long long src = src_rows ? src_rows[row] : row;
    float4 v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      v[j] = (EXACT || i < nvec) ? ld4<XBF>(x, src * D + 4LL * i) : f4zero();
    }
    if (yadd != nullptr) {  // fused residual update  x_new = x + gate[sample] * y  (dit.py:236-238), written back
      float4 ya[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int i = lane + 32 * j;
        ya[j] = (EXACT || i < nvec) ? ld4a(yadd + src * D + 4LL * i) : f4zero();
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int i = lane + 32 * j;
        if (EXACT || i < nvec) {
          const float4 g = *reinterpret_cast<const float4*>(sga + 4 * i);
          v[j].x += g.x * ya[j].x; v[j].y += g.y * ya[j].y; v[j].z += g.z * ya[j].z; v[j].w += g.w * ya[j].w;
          *reinterpret_cast<float4*>(xnew + src * D + 4LL * i) = v[j];
        }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) s += v[j].x + v[j].y + v[j].z + v[j].w;
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (EXACT || lane + 32 * j < nvec) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        ss += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      if (EXACT || i < nvec) {
        const float4 gw = *reinterpret_cast<const float4*>(sgw + 4 * i);
        const float4 sh = *reinterpret_cast<const float4*>(ssh + 4 * i);
        float4 o;
        o.x = fmaf((v[j].x - mean) * rstd, gw.x, sh.x); o.y = fmaf((v[j].y - mean) * rstd, gw.y, sh.y);
        o.z = fmaf((v[j].z - mean) * rstd, gw.z, sh.z); o.w = fmaf((v[j].w - mean) * rstd, gw.w, sh.w);
        st4a(y + row * D + 4LL * i, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ ln_bwd
// grid (ceil(T / rpb), samples); 4 warps share the block's rows.  Per-column partials A = sum dy,
// Bc = sum dy*xhat over the block's rows (all of one sample, so scale is constant): dshift += A,
// dscale += gamma*Bc, dgamma += (1+scale)*Bc.
template <int VEC, bool EXACT, bool XBF, typename AT>
__global__ void __launch_bounds__(128)
ln_bwd_kernel(const AT* __restrict__ dy, const void* __restrict__ x, const int32_t* __restrict__ src_rows,
              const float* __restrict__ gamma, const float* __restrict__ scale, long long ldmod, long long T,
              const float* __restrict__ mean, const float* __restrict__ rstd, void* __restrict__ dx, int dx_mode,
              float* __restrict__ dgamma, float* __restrict__ dshift, float* __restrict__ dscale, long long rows,
              int D, int rpb, float* __restrict__ dgamma_ws) {
  extern __shared__ float red[];  // [4][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long smp = blockIdx.y;
  const long long t0 = 1LL * blockIdx.x * rpb;
  const long long t1 = min(T, t0 + rpb);
  const int nvec = D >> 2;
  const float* sc = scale ? scale + smp * ldmod : nullptr;
  const bool need_cols = (dgamma != nullptr) || (dshift != nullptr) || (dscale != nullptr);

  // per-column weight d xhat / d y : gamma * (1 + scale), constant over the block
  float4 w[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int i = lane + 32 * j;
    w[j] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (EXACT || i < nvec) {
      if (gamma) w[j] = *reinterpret_cast<const float4*>(gamma + 4 * i);
      if (sc) {
        const float4 a = *reinterpret_cast<const float4*>(sc + 4 * i);
        w[j].x *= 1.f + a.x; w[j].y *= 1.f + a.y; w[j].z *= 1.f + a.z; w[j].w *= 1.f + a.w;
      }
    }
  }
  float4 accA[VEC], accB[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) accA[j] = accB[j] = f4zero();

  for (long long t = t0 + warp; t < t1; t += 4) {
    const long long row = smp * T + t;
    if (row >= rows) break;
    const long long src = src_rows ? src_rows[row] : row;
    const long long drow = (dx_mode == 2) ? src : row;
    float4 d[VEC], xh[VEC], old[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      const bool ok = EXACT || i < nvec;
      d[j] = ok ? ld4a(dy + row * D + 4LL * i) : f4zero();
      xh[j] = ok ? ld4<XBF>(x, src * D + 4LL * i) : f4zero();
    }
    if (dx != nullptr && dx_mode != 1) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int i = lane + 32 * j;
        old[j] = (EXACT || i < nvec) ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dx) + drow * D + 4LL * i)
                                     : f4zero();
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (EXACT || lane + 32 * j < nvec) {
        xh[j] = make_float4((xh[j].x - mu) * rs, (xh[j].y - mu) * rs, (xh[j].z - mu) * rs, (xh[j].w - mu) * rs);
        accA[j].x += d[j].x; accA[j].y += d[j].y; accA[j].z += d[j].z; accA[j].w += d[j].w;
        accB[j].x += d[j].x * xh[j].x; accB[j].y += d[j].y * xh[j].y;
        accB[j].z += d[j].z * xh[j].z; accB[j].w += d[j].w * xh[j].w;
        d[j] = make_float4(d[j].x * w[j].x, d[j].y * w[j].y, d[j].z * w[j].z, d[j].w * w[j].w);  // d loss / d xhat
        s1 += d[j].x + d[j].y + d[j].z + d[j].w;
        s2 += d[j].x * xh[j].x + d[j].y * xh[j].y + d[j].z * xh[j].z + d[j].w * xh[j].w;
      }
    }
    const float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D;
    if (dx != nullptr) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int i = lane + 32 * j;
        if (EXACT || i < nvec) {
          float4 o;
          o.x = rs * (d[j].x - m1 - xh[j].x * m2); o.y = rs * (d[j].y - m1 - xh[j].y * m2);
          o.z = rs * (d[j].z - m1 - xh[j].z * m2); o.w = rs * (d[j].w - m1 - xh[j].w * m2);
          if (dx_mode == 1) {
            st4a(reinterpret_cast<AT*>(dx) + drow * D + 4LL * i, o);
          } else {
            o.x += old[j].x; o.y += old[j].y; o.z += old[j].z; o.w += old[j].w;
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(dx) + drow * D + 4LL * i) = o;
          }
        }
      }
    }
  }
  if (!need_cols) return;
  // cross-warp reduction of the column partials, one quantity at a time through red[4][D]
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      if (EXACT || i < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * i) = pass == 0 ? accA[j] : accB[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const float v = red[c] + red[D + c] + red[2 * D + c] + red[3 * D + c];
      if (pass == 0) {
        if (dshift) atomicAdd(dshift + smp * ldmod + c, v);
      } else {
        if (dscale) atomicAdd(dscale + smp * ldmod + c, v * (gamma ? gamma[c] : 1.f));
        if (dgamma) {
          // deterministic mode (one block per sample): the per-sample term goes to the workspace, summed over samples later
          if (dgamma_ws) dgamma_ws[smp * D + c] = v * (sc ? 1.f + sc[c] : 1.f);
          else atomicAdd(dgamma + c, v * (sc ? 1.f + sc[c] : 1.f));
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------- ln_bwd (teams)
// The model widths (512 / 768 / 1024): a row is shared by a TEAM of 2 or 4 warps (D = 128 * TEAM * V: V float4 groups per
// lane), NT teams per block.  Cutting the per-lane footprint is what makes room for the fused tail
// below (one warp per row sat at 245 registers, 8 warps per SM); the partial row sums cross through a double-buffered
// shared-memory slot and a named barrier over the team.
//
// Fused tail (dy_next != NULL): dx_new is the gradient entering the NEXT branch of the backward chain (DiTBlock,
// dit.py:236-238 read backwards), whose first step used to be a separate pass over dx (md_gate_bwd):
//   dy_next = bf16(gate_next[sample] * dx_new),   dgate_next[sample] += sum_t dx_new * y_next.
template <int V, int TEAM, int NT, int kMinBlocks, bool kPrefetch, bool XBF, typename AT>
__global__ void __launch_bounds__(32 * TEAM * NT, kMinBlocks)
ln_bwd_team_kernel(const AT* __restrict__ dy, const void* __restrict__ x, const int32_t* __restrict__ src_rows,
                   const float* __restrict__ gamma, const float* __restrict__ scale, long long ldmod, long long T,
                   const float* __restrict__ mean, const float* __restrict__ rstd, void* __restrict__ dx, int dx_mode,
                   float* __restrict__ dgamma, float* __restrict__ dshift, float* __restrict__ dscale,
                   const AT* __restrict__ y_next, const float* __restrict__ gate_next, float* __restrict__ dgate_next,
                   AT* __restrict__ dy_next, int rpb, float* __restrict__ dgamma_ws) {
  constexpr int D = 128 * TEAM * V;
  constexpr int kThreads = 32 * TEAM * NT;  // NT teams per block
  extern __shared__ float sm[];
  float* red = sm;             // [NT][D]
  float* sw = sm + NT * D;     // [D] gamma * (1 + scale): d xhat / d y, constant over the block (one sample)
  float* sg = sw + D;          // [D] gate_next
  float* xch = sg + D;         // [NT][2 parities][TEAM warps][2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / TEAM, wt = warp % TEAM;
  const long long smp = blockIdx.y;
  const long long t0 = 1LL * blockIdx.x * rpb;
  const long long t1 = min(T, t0 + rpb);
  const float* sc = scale ? scale + smp * ldmod : nullptr;
  const float* gt = gate_next ? gate_next + smp * ldmod : nullptr;
  const bool need_cols = (dgamma != nullptr) || (dshift != nullptr) || (dscale != nullptr);
  const bool need_gate = (dgate_next != nullptr) && (y_next != nullptr);
  for (int c = threadIdx.x; c < D; c += kThreads) {
    sw[c] = (gamma ? gamma[c] : 1.f) * (sc ? 1.f + sc[c] : 1.f);
    sg[c] = gt ? gt[c] : 1.f;
  }
  __syncthreads();

  float4 accA[V], accB[V], accG[V];
#pragma unroll
  for (int j = 0; j < V; ++j) accA[j] = accB[j] = accG[j] = f4zero();
  const int i0 = wt * 32 + lane;  // float4 group index of this lane: i0 + 32 * TEAM * j
  const bool want_old = dx != nullptr && dx_mode != 1;
  // One row ahead: the loads of the team's next row are issued before this row's reductions and barrier, so a team
  // keeps two rows in flight (with four warps per row only four rows per SM were in flight otherwise, half of what
  // the one-warp-per-row kernel had, and the fused kernel ran at 0.65 of the HBM roofline).
  float4 nd[V], nx[V], nold[V], ny[V];
  float nmu = 0.f, nrs = 0.f;
  long long ndrow = 0;
  auto prefetch = [&](long long t) {
    const long long row = smp * T + t;
    const long long src = src_rows ? src_rows[row] : row;
    ndrow = (dx_mode == 2) ? src : row;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int i = i0 + 32 * TEAM * j;
      nd[j] = ld4a(dy + row * D + 4LL * i);
      nx[j] = ld4<XBF>(x, src * D + 4LL * i);
    }
    if (want_old) {
#pragma unroll
      for (int j = 0; j < V; ++j)
        nold[j] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dx) + ndrow * D + 4LL * (i0 + 32 * TEAM * j));
    }
    if (need_gate) {
#pragma unroll
      for (int j = 0; j < V; ++j) ny[j] = ld4a(y_next + row * D + 4LL * (i0 + 32 * TEAM * j));
    }
    nmu = mean[row];
    nrs = rstd[row];
  };
  if (t0 + team < t1) prefetch(t0 + team);
  int it = 0;
  for (long long t = t0 + team; t < t1; t += NT, ++it) {
    const long long row = smp * T + t;
    const long long drow = ndrow;
    float4 d[V], xh[V], old[V], yv[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      d[j] = nd[j];
      xh[j] = nx[j];
      old[j] = nold[j];
      yv[j] = ny[j];
    }
    const float mu = nmu, rs = nrs;
    if (kPrefetch && t + NT < t1) prefetch(t + NT);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(sw + 4 * (i0 + 32 * TEAM * j));
      xh[j] = make_float4((xh[j].x - mu) * rs, (xh[j].y - mu) * rs, (xh[j].z - mu) * rs, (xh[j].w - mu) * rs);
      accA[j].x += d[j].x; accA[j].y += d[j].y; accA[j].z += d[j].z; accA[j].w += d[j].w;
      accB[j].x += d[j].x * xh[j].x; accB[j].y += d[j].y * xh[j].y;
      accB[j].z += d[j].z * xh[j].z; accB[j].w += d[j].w * xh[j].w;
      d[j] = make_float4(d[j].x * w.x, d[j].y * w.y, d[j].z * w.z, d[j].w * w.w);  // d loss / d xhat
      s1 += d[j].x + d[j].y + d[j].z + d[j].w;
      s2 += d[j].x * xh[j].x + d[j].y * xh[j].y + d[j].z * xh[j].z + d[j].w * xh[j].w;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    float* slot = xch + ((team * 2 + (it & 1)) * TEAM) * 2;
    if (lane == 0) {
      slot[wt * 2] = s1;
      slot[wt * 2 + 1] = s2;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(32 * TEAM) : "memory");
    s1 = 0.f;
    s2 = 0.f;
#pragma unroll
    for (int u = 0; u < TEAM; ++u) {  // fixed order: every warp of the team gets bit-identical sums
      s1 += slot[u * 2];
      s2 += slot[u * 2 + 1];
    }
    const float m1 = s1 * (1.f / D), m2 = s2 * (1.f / D);
    if (dx != nullptr) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int i = i0 + 32 * TEAM * j;
        float4 o;
        o.x = rs * (d[j].x - m1 - xh[j].x * m2); o.y = rs * (d[j].y - m1 - xh[j].y * m2);
        o.z = rs * (d[j].z - m1 - xh[j].z * m2); o.w = rs * (d[j].w - m1 - xh[j].w * m2);
        if (dx_mode == 1) {
          st4a(reinterpret_cast<AT*>(dx) + drow * D + 4LL * i, o);
        } else {
          o.x += old[j].x; o.y += old[j].y; o.z += old[j].z; o.w += old[j].w;
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(dx) + drow * D + 4LL * i) = o;
          if (dy_next != nullptr) {
            const float4 g = *reinterpret_cast<const float4*>(sg + 4 * i);
            st4a(dy_next + row * D + 4LL * i, make_float4(o.x * g.x, o.y * g.y, o.z * g.z, o.w * g.w));
            if (need_gate) {
              accG[j].x += o.x * yv[j].x; accG[j].y += o.y * yv[j].y;
              accG[j].z += o.z * yv[j].z; accG[j].w += o.w * yv[j].w;
            }
          }
        }
      }
    }
    if (!kPrefetch && t + NT < t1) prefetch(t + NT);
  }
  if (!need_cols && !need_gate) return;
  // cross-team reduction of the column partials, one quantity at a time through red[NT][D]
  for (int pass = 0; pass < 3; ++pass) {
    if (pass < 2 && !need_cols) continue;
    if (pass == 2 && !need_gate) continue;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < V; ++j)
      *reinterpret_cast<float4*>(red + team * D + 4 * (i0 + 32 * TEAM * j)) =
          pass == 0 ? accA[j] : (pass == 1 ? accB[j] : accG[j]);
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += kThreads) {
      float v = 0.f;
#pragma unroll
      for (int u = 0; u < NT; ++u) v += red[u * D + c];
      if (pass == 0) {
        if (dshift) atomicAdd(dshift + smp * ldmod + c, v);
      } else if (pass == 1) {
        if (dscale) atomicAdd(dscale + smp * ldmod + c, v * (gamma ? gamma[c] : 1.f));
        if (dgamma) {
          if (dgamma_ws) dgamma_ws[smp * D + c] = v * (sc ? 1.f + sc[c] : 1.f);   // deterministic mode, see ln_bwd_kernel
          else atomicAdd(dgamma + c, v * (sc ? 1.f + sc[c] : 1.f));
        }
      } else {
        atomicAdd(dgate_next + smp * ldmod + c, v);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------- rownorm
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
// VEC = uint4 groups per lane (8 bf16 each): W <= 256 * VEC
template <int VEC, typename AT>
__global__ void __launch_bounds__(128)
rownorm_fwd_kernel(AT* __restrict__ x, long long ld, float* __restrict__ rstd_out, long long rows, int W, int nslice,
                   float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = W >> 3;
  // virtual row = (row, slice): the slices of one row are adjacent in memory and go to consecutive warps
  for (long long vr = 1LL * blockIdx.x * 4 + warp; vr < rows * nslice; vr += 1LL * gridDim.x * 4) {
    const long long row = vr / nslice;
    const int sl = static_cast<int>(vr - row * nslice);
    AT* p = x + row * ld + sl * W;
    V8<AT> raw[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      raw[j] = i < nvec ? ldv8(p + 8 * i) : zerov8<AT>();
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float v[8];
      unpackv8(raw[j], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[e];
    }
    const float mean = warp_sum(s) / W;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (lane + 32 * j < nvec) {
        float v[8];
        unpackv8(raw[j], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += (v[e] - mean) * (v[e] - mean);
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / W + eps);
    if (lane == 0) rstd_out[sl * rows + row] = rstd;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        float v[8];
        unpackv8(raw[j], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean) * rstd;
        st8(p + 8 * i, v);
      }
    }
  }
}

template <int VEC, typename AT>
__global__ void __launch_bounds__(128)
rownorm_bwd_kernel(AT* __restrict__ dy, long long ld_dy, const AT* __restrict__ xhat, long long ld_x,
                   const float* __restrict__ rstd, long long rows, int W, int nslice) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = W >> 3;
  for (long long vr = 1LL * blockIdx.x * 4 + warp; vr < rows * nslice; vr += 1LL * gridDim.x * 4) {
    const long long row = vr / nslice;
    const int sl = static_cast<int>(vr - row * nslice);
    AT* pd = dy + row * ld_dy + sl * W;
    const AT* px = xhat + row * ld_x + sl * W;
    V8<AT> rd[VEC], rx[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      rd[j] = i < nvec ? ldv8(pd + 8 * i) : zerov8<AT>();
      rx[j] = i < nvec ? ldv8(px + 8 * i) : zerov8<AT>();
    }
    const float rs = rstd[sl * rows + row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float d[8], xh[8];
      unpackv8(rd[j], d);
      unpackv8(rx[j], xh);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s1 += d[e];
        s2 += d[e] * xh[e];
      }
    }
    const float m1 = warp_sum(s1) / W, m2 = warp_sum(s2) / W;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        float d[8], xh[8];
        unpackv8(rd[j], d);
        unpackv8(rx[j], xh);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = rs * (d[e] - m1 - xh[e] * m2);
        st8(pd + 8 * i, d);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------- gate_bwd
template <int VEC, bool EXACT, typename AT>
__global__ void __launch_bounds__(128)
gate_bwd_kernel(const float* __restrict__ dres, const AT* __restrict__ y, const float* __restrict__ gate,
                long long ldmod, long long T, AT* __restrict__ dy, float* __restrict__ dgate, long long rows, int D,
                int rpb) {
  extern __shared__ float red[];  // [4][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long smp = blockIdx.y;
  const long long t0 = 1LL * blockIdx.x * rpb;
  const long long t1 = min(T, t0 + rpb);
  const int nvec = D >> 2;
  const float* gt = gate ? gate + smp * ldmod : nullptr;
  const bool need = (dgate != nullptr) && (y != nullptr);
  float4 g[VEC], acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int i = lane + 32 * j;
    acc[j] = f4zero();
    g[j] = (gt && (EXACT || i < nvec)) ? *reinterpret_cast<const float4*>(gt + 4 * i) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  for (long long t = t0 + warp; t < t1; t += 4) {
    const long long row = smp * T + t;
    if (row >= rows) break;
    float4 d[VEC], yv[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      const bool ok = EXACT || i < nvec;
      d[j] = ok ? *reinterpret_cast<const float4*>(dres + row * D + 4LL * i) : f4zero();
      if (need) yv[j] = ok ? ld4a(y + row * D + 4LL * i) : f4zero();
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      if (EXACT || i < nvec) {
        if (need) {
          acc[j].x += d[j].x * yv[j].x; acc[j].y += d[j].y * yv[j].y;
          acc[j].z += d[j].z * yv[j].z; acc[j].w += d[j].w * yv[j].w;
        }
        st4a(dy + row * D + 4LL * i, make_float4(d[j].x * g[j].x, d[j].y * g[j].y, d[j].z * g[j].z, d[j].w * g[j].w));
      }
    }
  }
  if (!need) return;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int i = lane + 32 * j;
    if (EXACT || i < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * i) = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    atomicAdd(dgate + smp * ldmod + c, red[c] + red[D + c] + red[2 * D + c] + red[3 * D + c]);
}

static int check_ln_dims(const char* what, long long rows, long long D, long long T) {
  if (rows < 0 || D <= 0 || T <= 0) return md_set_error(MD_ERR_INVALID, what);
  if (D % 8 != 0 || D > 4 * 32 * kMaxVec) {
    char buf[128];
    snprintf(buf, sizeof(buf), "%s: D=%lld unsupported (need D %% 8 == 0 and D <= %d)", what, D, 4 * 32 * kMaxVec);
    return md_set_error(MD_ERR_UNSUPPORTED, buf);
  }
  return 0;
}

}  // namespace md

using namespace md;

static int row_grid(long long rows) {
  long long blocks = (rows + 3) / 4;
  const long long cap = 148LL * 12;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}
static int rows_per_block(long long T, long long samples) {
  // enough CTAs to fill the machine a few times over, few enough to keep the column atomics cheap
  int rpb = 32;
  while (rpb > 4 && ((T + rpb - 1) / rpb) * samples < 148LL * 6) rpb >>= 1;
  return rpb;
}

extern "C" int md_ln_fwd(const void* x, int x_bf16, const int32_t* src_rows, const void* y_add, const float* gate_add,
                         float* x_new, const float* gamma, const float* shift, const float* scale, int64_t ldmod,
                         int64_t T, void* y, float* mean, float* rstd, int64_t rows, int64_t D, float eps, int prec,
                         void* stream) {
  if (int rc = check_ln_dims("md_ln_fwd", rows, D, T)) return rc;
  if (rows == 0) return 0;
  if (!x || !y) return md_set_error(MD_ERR_INVALID, "md_ln_fwd: null pointer");
  if (y_add != nullptr && (x_new == nullptr || x_bf16)) return md_set_error(MD_ERR_INVALID, "md_ln_fwd: residual add needs x_new and f32 x");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (rows % T != 0) return md_set_error(MD_ERR_INVALID, "md_ln_fwd: rows must be a multiple of T");
  const int rpb = rows_per_block(T, rows / T);
  dim3 grid(static_cast<unsigned>((T + rpb - 1) / rpb), static_cast<unsigned>(rows / T));
  const size_t smem = 3 * D * sizeof(float);
#define LN_FWD(VEC, EXACT)                                                                                              \
  MD_WITH_ACT(prec, do {                                                                                                \
    if (x_bf16)                                                                                                         \
      ln_fwd_kernel<VEC, EXACT, true, AT><<<grid, 128, smem, st>>>(x, src_rows, CAP(AT, y_add), gate_add, x_new, gamma, \
                                                                   shift, scale, ldmod, T, AP(AT, y), mean, rstd, rows, \
                                                                   static_cast<int>(D), eps, rpb);                     \
    else                                                                                                                \
      ln_fwd_kernel<VEC, EXACT, false, AT><<<grid, 128, smem, st>>>(x, src_rows, CAP(AT, y_add), gate_add, x_new, gamma,\
                                                                    shift, scale, ldmod, T, AP(AT, y), mean, rstd,     \
                                                                    rows, static_cast<int>(D), eps, rpb);              \
  } while (0))
  if (D == 1024) LN_FWD(8, true);
  else if (D == 768) LN_FWD(6, true);
  else if (D == 512) LN_FWD(4, true);
  else if (D <= 1024) LN_FWD(8, false);
  else LN_FWD(16, false);
#undef LN_FWD
  return check_launch("md_ln_fwd");
}

extern "C" int md_ln_bwd(const void* dy, const void* x, int x_bf16, const int32_t* src_rows, const float* gamma,
                         const float* scale, int64_t ldmod, int64_t T, const float* mean, const float* rstd, void* dx,
                         int dx_mode, float* dgamma, float* dshift, float* dscale, const void* y_next,
                         const float* gate_next, float* dgate_next, void* dy_next, int64_t rows, int64_t D, int prec,
                         void* stream) {
  if (int rc = check_ln_dims("md_ln_bwd", rows, D, T)) return rc;
  if (rows == 0) return 0;
  if (!dy || !x || !mean || !rstd) return md_set_error(MD_ERR_INVALID, "md_ln_bwd: null pointer");
  if (rows % T != 0) return md_set_error(MD_ERR_INVALID, "md_ln_bwd: rows must be a multiple of T");
  if (dx_mode < 0 || dx_mode > 2 || (dx_mode == 2 && !src_rows))
    return md_set_error(MD_ERR_INVALID, "md_ln_bwd: bad dx_mode");
  if (dy_next != nullptr && (dx == nullptr || dx_mode != 0))
    return md_set_error(MD_ERR_INVALID, "md_ln_bwd: the fused next-branch tail needs dx with dx_mode 0");
  if ((y_next != nullptr || gate_next != nullptr || dgate_next != nullptr) && dy_next == nullptr)
    return md_set_error(MD_ERR_INVALID, "md_ln_bwd: y_next / gate_next / dgate_next come with dy_next");
  // deterministic mode: one block per sample (a single contributor to every per-sample atomic) and the cross-sample sum
  // of d gamma through the workspace in a fixed order
  const bool det = det_enabled();
  const int rpb = det ? static_cast<int>(T) : rows_per_block(T, rows / T);
  dim3 grid(static_cast<unsigned>((T + rpb - 1) / rpb), static_cast<unsigned>(rows / T));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* gws = nullptr;
  if (det && dgamma != nullptr) {
    gws = det_workspace(static_cast<size_t>(rows / T) * D * sizeof(float));
    if (gws == nullptr) return md_set_error(MD_ERR_INVALID, "md_ln_bwd: deterministic workspace too small");
  }
  auto finish = [&]() -> int {
    if (int rc = check_launch("md_ln_bwd")) return rc;
    return gws ? det_reduce(gws, dgamma, rows / T, D, 1, st) : 0;
  };
  if (D == 1024 || D == 768 || D == 512) {
    // width -> (float4 groups per lane, warps per row, rows in flight per block, blocks per SM, one-row-ahead loads)
#define LN_BWD_TEAM(V, TEAM, NT, MINB, PF)                                                                              \
  MD_WITH_ACT(prec, do {                                                                                               \
    const size_t smem = ((NT + 2) * D + 64) * sizeof(float);                                                          \
    if (x_bf16)                                                                                                        \
      ln_bwd_team_kernel<V, TEAM, NT, MINB, PF, true, AT><<<grid, 32 * TEAM * NT, smem, st>>>(                          \
          CAP(AT, dy), x, src_rows, gamma, scale, ldmod, T, mean, rstd, dx, dx_mode, dgamma, dshift, dscale,           \
          CAP(AT, y_next), gate_next, dgate_next, AP(AT, dy_next), rpb, gws);                                          \
    else                                                                                                               \
      ln_bwd_team_kernel<V, TEAM, NT, MINB, PF, false, AT><<<grid, 32 * TEAM * NT, smem, st>>>(                         \
          CAP(AT, dy), x, src_rows, gamma, scale, ldmod, T, mean, rstd, dx, dx_mode, dgamma, dshift, dscale,           \
          CAP(AT, y_next), gate_next, dgate_next, AP(AT, dy_next), rpb, gws);                                          \
  } while (0))
    if (D == 1024) LN_BWD_TEAM(2, 4, 2, 2, true);
    else if (D == 768) LN_BWD_TEAM(2, 3, 2, 2, true);
    else LN_BWD_TEAM(2, 2, 4, 2, true);
#undef LN_BWD_TEAM
    return finish();
  }
  const size_t smem = 4 * D * sizeof(float);
#define LN_BWD(VEC, EXACT)                                                                                             \
  MD_WITH_ACT(prec, do {                                                                                               \
    if (x_bf16)                                                                                                        \
      ln_bwd_kernel<VEC, EXACT, true, AT><<<grid, 128, smem, st>>>(CAP(AT, dy), x, src_rows, gamma, scale, ldmod, T,  \
                                                                   mean, rstd, dx, dx_mode, dgamma, dshift, dscale,   \
                                                                   rows, static_cast<int>(D), rpb, gws);              \
    else                                                                                                               \
      ln_bwd_kernel<VEC, EXACT, false, AT><<<grid, 128, smem, st>>>(CAP(AT, dy), x, src_rows, gamma, scale, ldmod, T, \
                                                                    mean, rstd, dx, dx_mode, dgamma, dshift, dscale,  \
                                                                    rows, static_cast<int>(D), rpb, gws);             \
  } while (0))
  if (D <= 1024) LN_BWD(8, false);
  else LN_BWD(16, false);
#undef LN_BWD
  if (int rc = finish()) return rc;
  // other widths: the next branch's gate backward runs as its own pass over the updated dx
  if (dy_next != nullptr)
    return md_gate_bwd(reinterpret_cast<const float*>(dx), y_next, gate_next, ldmod, T, dy_next, dgate_next, rows, D, prec,
                       stream);
  return 0;
}

extern "C" int md_rownorm_fwd(void* x, int64_t ld, float* rstd, int64_t rows, int64_t W, int64_t nslice, float eps,
                              int prec, void* stream) {
  if (rows == 0) return 0;
  if (nslice < 1 || nslice > 4) return md_set_error(MD_ERR_INVALID, "md_rownorm_fwd: need 1 <= nslice <= 4");
  if (!x || !rstd) return md_set_error(MD_ERR_INVALID, "md_rownorm_fwd: null pointer");
  if (W % 8 != 0 || W > 2048 || ld % 8 != 0)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_rownorm_fwd: need W % 8 == 0, W <= 2048, ld % 8 == 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (W <= 1024)
    MD_WITH_ACT(prec, rownorm_fwd_kernel<4, AT><<<row_grid(rows * nslice), 128, 0, st>>>(AP(AT, x), ld, rstd, rows, static_cast<int>(W), static_cast<int>(nslice), eps));
  else
    MD_WITH_ACT(prec, rownorm_fwd_kernel<8, AT><<<row_grid(rows * nslice), 128, 0, st>>>(AP(AT, x), ld, rstd, rows, static_cast<int>(W), static_cast<int>(nslice), eps));
  return check_launch("md_rownorm_fwd");
}

extern "C" int md_rownorm_bwd(void* dy, int64_t ld_dy, const void* xhat, int64_t ld_x, const float* rstd, int64_t rows,
                              int64_t W, int64_t nslice, int prec, void* stream) {
  if (rows == 0) return 0;
  if (nslice < 1 || nslice > 4) return md_set_error(MD_ERR_INVALID, "md_rownorm_bwd: need 1 <= nslice <= 4");
  if (!dy || !xhat || !rstd) return md_set_error(MD_ERR_INVALID, "md_rownorm_bwd: null pointer");
  if (W % 8 != 0 || W > 2048 || ld_dy % 8 != 0 || ld_x % 8 != 0)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_rownorm_bwd: need W % 8 == 0, W <= 2048, ld % 8 == 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (W <= 1024)
    MD_WITH_ACT(prec, rownorm_bwd_kernel<4, AT><<<row_grid(rows * nslice), 128, 0, st>>>(AP(AT, dy), ld_dy, CAP(AT, xhat), ld_x, rstd, rows, static_cast<int>(W), static_cast<int>(nslice)));
  else
    MD_WITH_ACT(prec, rownorm_bwd_kernel<8, AT><<<row_grid(rows * nslice), 128, 0, st>>>(AP(AT, dy), ld_dy, CAP(AT, xhat), ld_x, rstd, rows, static_cast<int>(W), static_cast<int>(nslice)));
  return check_launch("md_rownorm_bwd");
}

extern "C" int md_gate_bwd(const float* dres, const void* y, const float* gate, int64_t ldmod, int64_t T, void* dy,
                           float* dgate, int64_t rows, int64_t D, int prec, void* stream) {
  if (int rc = check_ln_dims("md_gate_bwd", rows, D, T)) return rc;
  if (rows == 0) return 0;
  if (!dres || !dy) return md_set_error(MD_ERR_INVALID, "md_gate_bwd: null pointer");
  if (rows % T != 0) return md_set_error(MD_ERR_INVALID, "md_gate_bwd: rows must be a multiple of T");
  const int rpb = det_enabled() ? static_cast<int>(T) : rows_per_block(T, rows / T);  // deterministic: one block per sample
  dim3 grid(static_cast<unsigned>((T + rpb - 1) / rpb), static_cast<unsigned>(rows / T));
  const size_t smem = 4 * D * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define GATE(VEC, EXACT)                                                                                               \
  MD_WITH_ACT(prec, gate_bwd_kernel<VEC, EXACT, AT><<<grid, 128, smem, st>>>(dres, CAP(AT, y), gate, ldmod, T, AP(AT, dy), \
                                                                             dgate, rows, static_cast<int>(D), rpb))
  if (D == 1024) GATE(8, true);
  else if (D == 768) GATE(6, true);
  else if (D == 512) GATE(4, true);
  else if (D <= 1024) GATE(8, false);
  else GATE(16, false);
#undef GATE
  return check_launch("md_gate_bwd");
}
