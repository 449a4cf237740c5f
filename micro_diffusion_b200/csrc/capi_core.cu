// Error bookkeeping and version query of the C ABI.
#include <string.h>

#include "common.cuh"

namespace md {
static thread_local char g_err[512] = "";
int md_set_error(int code, const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return code;
}

// ---- deterministic mode (md_set_deterministic): a caller-owned scratch buffer replaces every cross-block atomic
// accumulation by per-block partials + a fixed-order reduction.  One buffer, reused by consecutive launches: all
// deterministic-mode calls must be issued on ONE stream.
static float* g_det_ws = nullptr;
static size_t g_det_bytes = 0;
bool det_enabled() { return g_det_ws != nullptr; }
float* det_workspace(size_t need_bytes) { return (g_det_ws != nullptr && need_bytes <= g_det_bytes) ? g_det_ws : nullptr; }

// out[i * out_stride] += sum_p ws[p * n + i], p in a fixed order (8 interleaved sub-sums, combined 0..7)
__global__ void __launch_bounds__(256)
det_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, long long parts, long long n, long long out_stride) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long i = 1LL * blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < n)
    for (long long q = ty; q < parts; q += 8) s += ws[q * n + i];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += part[g][tx];
    out[i * out_stride] += t;
  }
}
int det_reduce(const float* ws, float* out, long long parts, long long n, long long out_stride, cudaStream_t stream) {
  if (parts <= 0 || n <= 0) return 0;
  det_reduce_kernel<<<static_cast<unsigned>((n + 31) / 32), 256, 0, stream>>>(ws, out, parts, n, out_stride);
  return check_launch("deterministic reduction");
}
}  // namespace md

extern "C" int md_set_deterministic(void* workspace, int64_t bytes) {
  if (workspace != nullptr && (bytes < (1 << 20) || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0))
    return md::md_set_error(MD_ERR_INVALID, "md_set_deterministic: workspace must be >= 1 MiB and 256-byte aligned (NULL turns the mode off)");
  md::g_det_ws = reinterpret_cast<float*>(workspace);
  md::g_det_bytes = workspace ? static_cast<size_t>(bytes) : 0;
  return 0;
}
extern "C" const char* md_last_error(void) { return md::g_err; }
extern "C" int md_abi_version(void) { return 4; }
