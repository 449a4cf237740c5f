// LayerNorm(+adaLN modulate) forward/backward, QK row-norm, gated-residual backward.
// All HBM-bound: one warp per row, 16-byte vector loads, warp-shuffle reductions, fp32 statistics.
// Reference semantics: create_norm (utils.py:71-78), modulate (utils.py:28-30), DiTBlock.forward
// (dit.py:232-239), ln_q/ln_k over the full hidden width (utils.py:183-186, 122-125).
#include "common.cuh"

namespace md {

constexpr int kMaxVec = 16;  // float4 groups per lane  -> D <= 2048
constexpr int kRowsPerBlock = 32;

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

// ------------------------------------------------------------------------------------------ ln_fwd
__global__ void __launch_bounds__(128)
ln_fwd_kernel(const void* __restrict__ x, int x_bf16, const int32_t* __restrict__ src_rows,
              const float* __restrict__ gamma, const float* __restrict__ shift, const float* __restrict__ scale,
              long long ldmod, long long T, void* __restrict__ y, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, long long rows, int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = 1LL * blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const long long src = src_rows ? src_rows[row] : row;
  const int nvec = D >> 2;
  float4 v[kMaxVec];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      v[j] = load4(x, x_bf16, src * D + 4LL * i);
      s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  }
  const float mean = warp_sum(s) / D;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const long long smp = row / T;
  const float* sh = shift ? shift + smp * ldmod : nullptr;
  const float* sc = scale ? scale + smp * ldmod : nullptr;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      float4 o;
      o.x = (v[j].x - mean) * rstd; o.y = (v[j].y - mean) * rstd;
      o.z = (v[j].z - mean) * rstd; o.w = (v[j].w - mean) * rstd;
      if (gamma) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * i);
        o.x *= g.x; o.y *= g.y; o.z *= g.z; o.w *= g.w;
      }
      if (sc) {
        const float4 a = *reinterpret_cast<const float4*>(sc + 4 * i);
        o.x *= 1.f + a.x; o.y *= 1.f + a.y; o.z *= 1.f + a.z; o.w *= 1.f + a.w;
      }
      if (sh) {
        const float4 a = *reinterpret_cast<const float4*>(sh + 4 * i);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      store4_bf16(y, row * D + 4LL * i, o);
    }
  }
}

// ------------------------------------------------------------------------------------------ ln_bwd
// grid (ceil(T / 32), samples); 4 warps, each 8 rows.  Per-column partials A = sum dy, Bc = sum dy*xhat
// over the block's rows (all of one sample, so scale is constant): dshift += A, dscale += gamma*Bc,
// dgamma += (1+scale)*Bc.
template <int VEC>
__global__ void __launch_bounds__(128)
ln_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, int x_bf16,
              const int32_t* __restrict__ src_rows, const float* __restrict__ gamma,
              const float* __restrict__ scale, long long ldmod, long long T, const float* __restrict__ mean,
              const float* __restrict__ rstd, void* __restrict__ dx, int dx_mode, float* __restrict__ dgamma,
              float* __restrict__ dshift, float* __restrict__ dscale, long long rows, int D) {
  extern __shared__ float red[];  // [4][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long smp = blockIdx.y;
  const long long t0 = 1LL * blockIdx.x * kRowsPerBlock;
  const long long t1 = min(T, t0 + kRowsPerBlock);
  const int nvec = D >> 2;
  const float* sc = scale ? scale + smp * ldmod : nullptr;
  const bool need_cols = (dgamma != nullptr) || (dshift != nullptr) || (dscale != nullptr);

  float4 accA[VEC], accB[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    accA[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    accB[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long t = t0 + warp; t < t1; t += 4) {
    const long long row = smp * T + t;
    if (row >= rows) break;
    const long long src = src_rows ? src_rows[row] : row;
    const float mu = mean[row], rs = rstd[row];
    float4 g[VEC], xh[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        const float4 d = load4(dy, true, row * D + 4LL * i);
        const float4 xv = load4(x, x_bf16, src * D + 4LL * i);
        xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        accA[j].x += d.x; accA[j].y += d.y; accA[j].z += d.z; accA[j].w += d.w;
        accB[j].x += d.x * xh[j].x; accB[j].y += d.y * xh[j].y;
        accB[j].z += d.z * xh[j].z; accB[j].w += d.w * xh[j].w;
        float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
        if (gamma) w = *reinterpret_cast<const float4*>(gamma + 4 * i);
        if (sc) {
          const float4 a = *reinterpret_cast<const float4*>(sc + 4 * i);
          w.x *= 1.f + a.x; w.y *= 1.f + a.y; w.z *= 1.f + a.z; w.w *= 1.f + a.w;
        }
        g[j] = make_float4(d.x * w.x, d.y * w.y, d.z * w.z, d.w * w.w);  // d loss / d xhat
        s1 += g[j].x + g[j].y + g[j].z + g[j].w;
        s2 += g[j].x * xh[j].x + g[j].y * xh[j].y + g[j].z * xh[j].z + g[j].w * xh[j].w;
      }
    }
    const float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D;
    if (dx != nullptr) {
      const long long drow = (dx_mode == 2) ? src : row;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int i = lane + 32 * j;
        if (i < nvec) {
          float4 o;
          o.x = rs * (g[j].x - m1 - xh[j].x * m2); o.y = rs * (g[j].y - m1 - xh[j].y * m2);
          o.z = rs * (g[j].z - m1 - xh[j].z * m2); o.w = rs * (g[j].w - m1 - xh[j].w * m2);
          if (dx_mode == 1) {
            store4_bf16(dx, drow * D + 4LL * i, o);
          } else {
            float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(dx) + drow * D + 4LL * i);
            float4 c = *p;
            c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
            *p = c;
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
      if (i < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * i) = pass == 0 ? accA[j] : accB[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const float v = red[c] + red[D + c] + red[2 * D + c] + red[3 * D + c];
      if (pass == 0) {
        if (dshift) atomicAdd(dshift + smp * ldmod + c, v);
      } else {
        if (dscale) atomicAdd(dscale + smp * ldmod + c, v * (gamma ? gamma[c] : 1.f));
        if (dgamma) atomicAdd(dgamma + c, v * (sc ? 1.f + sc[c] : 1.f));
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
constexpr int kRnVec = 8;  // uint4 groups per lane -> W <= 2048

__global__ void __launch_bounds__(128)
rownorm_fwd_kernel(__nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ rstd_out, long long rows, int W,
                   float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = 1LL * blockIdx.x * 4 + warp;
  if (row >= rows) return;
  __nv_bfloat16* p = x + row * ld;
  const int nvec = W >> 3;
  float v[kRnVec][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kRnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      unpack8(*reinterpret_cast<const uint4*>(p + 8 * i), v[j]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[j][e];
    }
  }
  const float mean = warp_sum(s) / W;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kRnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[j][e] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / W + eps);
  if (lane == 0) rstd_out[row] = rstd;
#pragma unroll
  for (int j = 0; j < kRnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mean) * rstd;
      *reinterpret_cast<uint4*>(p + 8 * i) = pack8(o);
    }
  }
}

__global__ void __launch_bounds__(128)
rownorm_bwd_kernel(__nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ xhat,
                   long long ld_x, const float* __restrict__ rstd, long long rows, int W) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = 1LL * blockIdx.x * 4 + warp;
  if (row >= rows) return;
  __nv_bfloat16* pd = dy + row * ld_dy;
  const __nv_bfloat16* px = xhat + row * ld_x;
  const int nvec = W >> 3;
  float d[kRnVec][8], xh[kRnVec][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < kRnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      unpack8(*reinterpret_cast<const uint4*>(pd + 8 * i), d[j]);
      unpack8(*reinterpret_cast<const uint4*>(px + 8 * i), xh[j]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s1 += d[j][e];
        s2 += d[j][e] * xh[j][e];
      }
    }
  }
  const float m1 = warp_sum(s1) / W, m2 = warp_sum(s2) / W;
  const float rs = rstd[row];
#pragma unroll
  for (int j = 0; j < kRnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rs * (d[j][e] - m1 - xh[j][e] * m2);
      *reinterpret_cast<uint4*>(pd + 8 * i) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------- gate_bwd
template <int VEC>
__global__ void __launch_bounds__(128)
gate_bwd_kernel(const float* __restrict__ dres, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
                long long ldmod, long long T, __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate,
                long long rows, int D) {
  extern __shared__ float red[];  // [4][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long smp = blockIdx.y;
  const long long t0 = 1LL * blockIdx.x * kRowsPerBlock;
  const long long t1 = min(T, t0 + kRowsPerBlock);
  const int nvec = D >> 2;
  const float* gt = gate ? gate + smp * ldmod : nullptr;
  const bool need = (dgate != nullptr) && (y != nullptr);
  float4 acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long t = t0 + warp; t < t1; t += 4) {
    const long long row = smp * T + t;
    if (row >= rows) break;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        const float4 d = *reinterpret_cast<const float4*>(dres + row * D + 4LL * i);
        if (need) {
          const float4 yv = load4(y, true, row * D + 4LL * i);
          acc[j].x += d.x * yv.x; acc[j].y += d.y * yv.y; acc[j].z += d.z * yv.z; acc[j].w += d.w * yv.w;
        }
        float4 o = d;
        if (gt) {
          const float4 g = *reinterpret_cast<const float4*>(gt + 4 * i);
          o.x *= g.x; o.y *= g.y; o.z *= g.z; o.w *= g.w;
        }
        store4_bf16(dy, row * D + 4LL * i, o);
      }
    }
  }
  if (!need) return;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * i) = acc[j];
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

extern "C" int md_ln_fwd(const void* x, int x_bf16, const int32_t* src_rows, const float* gamma, const float* shift,
                         const float* scale, int64_t ldmod, int64_t T, void* y, float* mean, float* rstd,
                         int64_t rows, int64_t D, float eps, void* stream) {
  if (int rc = check_ln_dims("md_ln_fwd", rows, D, T)) return rc;
  if (rows == 0) return 0;
  if (!x || !y) return md_set_error(MD_ERR_INVALID, "md_ln_fwd: null pointer");
  ln_fwd_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, x_bf16, src_rows, gamma, shift, scale, ldmod, T, y, mean, rstd, rows, static_cast<int>(D), eps);
  return check_launch("md_ln_fwd");
}

extern "C" int md_ln_bwd(const void* dy, const void* x, int x_bf16, const int32_t* src_rows, const float* gamma,
                         const float* scale, int64_t ldmod, int64_t T, const float* mean, const float* rstd, void* dx,
                         int dx_mode, float* dgamma, float* dshift, float* dscale, int64_t rows, int64_t D,
                         void* stream) {
  if (int rc = check_ln_dims("md_ln_bwd", rows, D, T)) return rc;
  if (rows == 0) return 0;
  if (!dy || !x || !mean || !rstd) return md_set_error(MD_ERR_INVALID, "md_ln_bwd: null pointer");
  if (rows % T != 0) return md_set_error(MD_ERR_INVALID, "md_ln_bwd: rows must be a multiple of T");
  if (dx_mode < 0 || dx_mode > 2 || (dx_mode == 2 && !src_rows))
    return md_set_error(MD_ERR_INVALID, "md_ln_bwd: bad dx_mode");
  dim3 grid(static_cast<unsigned>((T + kRowsPerBlock - 1) / kRowsPerBlock), static_cast<unsigned>(rows / T));
  const size_t smem = 4 * D * sizeof(float);
  if (D <= 1024)
    ln_bwd_kernel<8><<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        dy, x, x_bf16, src_rows, gamma, scale, ldmod, T, mean, rstd, dx, dx_mode, dgamma, dshift, dscale, rows,
        static_cast<int>(D));
  else
    ln_bwd_kernel<16><<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        dy, x, x_bf16, src_rows, gamma, scale, ldmod, T, mean, rstd, dx, dx_mode, dgamma, dshift, dscale, rows,
        static_cast<int>(D));
  return check_launch("md_ln_bwd");
}

extern "C" int md_rownorm_fwd(void* x, int64_t ld, float* rstd, int64_t rows, int64_t W, float eps, void* stream) {
  if (rows == 0) return 0;
  if (!x || !rstd) return md_set_error(MD_ERR_INVALID, "md_rownorm_fwd: null pointer");
  if (W % 8 != 0 || W > 8 * 32 * kRnVec || ld % 8 != 0)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_rownorm_fwd: need W % 8 == 0, W <= 2048, ld % 8 == 0");
  rownorm_fwd_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(x), ld, rstd, rows, static_cast<int>(W), eps);
  return check_launch("md_rownorm_fwd");
}

extern "C" int md_rownorm_bwd(void* dy, int64_t ld_dy, const void* xhat, int64_t ld_x, const float* rstd, int64_t rows,
                              int64_t W, void* stream) {
  if (rows == 0) return 0;
  if (!dy || !xhat || !rstd) return md_set_error(MD_ERR_INVALID, "md_rownorm_bwd: null pointer");
  if (W % 8 != 0 || W > 8 * 32 * kRnVec || ld_dy % 8 != 0 || ld_x % 8 != 0)
    return md_set_error(MD_ERR_UNSUPPORTED, "md_rownorm_bwd: need W % 8 == 0, W <= 2048, ld % 8 == 0");
  rownorm_bwd_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(dy), ld_dy, reinterpret_cast<const __nv_bfloat16*>(xhat), ld_x, rstd, rows,
      static_cast<int>(W));
  return check_launch("md_rownorm_bwd");
}

extern "C" int md_gate_bwd(const float* dres, const void* y, const float* gate, int64_t ldmod, int64_t T, void* dy,
                           float* dgate, int64_t rows, int64_t D, void* stream) {
  if (int rc = check_ln_dims("md_gate_bwd", rows, D, T)) return rc;
  if (rows == 0) return 0;
  if (!dres || !dy) return md_set_error(MD_ERR_INVALID, "md_gate_bwd: null pointer");
  if (rows % T != 0) return md_set_error(MD_ERR_INVALID, "md_gate_bwd: rows must be a multiple of T");
  dim3 grid(static_cast<unsigned>((T + kRowsPerBlock - 1) / kRowsPerBlock), static_cast<unsigned>(rows / T));
  const size_t smem = 4 * D * sizeof(float);
  if (D <= 1024)
    gate_bwd_kernel<8><<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        dres, reinterpret_cast<const __nv_bfloat16*>(y), gate, ldmod, T, reinterpret_cast<__nv_bfloat16*>(dy), dgate,
        rows, static_cast<int>(D));
  else
    gate_bwd_kernel<16><<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        dres, reinterpret_cast<const __nv_bfloat16*>(y), gate, ldmod, T, reinterpret_cast<__nv_bfloat16*>(dy), dgate,
        rows, static_cast<int>(D));
  return check_launch("md_gate_bwd");
}
