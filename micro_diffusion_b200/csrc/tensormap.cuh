// cuTensorMapEncodeTiled behind a small per-thread cache.  A tensor map is a pure function of (base pointer, extents,
// pitches, box, swizzle), and the training step relaunches the same few hundred operand shapes at the same addresses
// every step (torch's caching allocator hands the same blocks back), so ~10 k driver calls per step (1-2 us each: a
// measurable share of the host time at 8 GPUs) become table look-ups.  Entries are overwritten on collision.
#pragma once
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace md {

struct TmapKey {
  const void* ptr;
  long long cols, rows, batch, ld, batch_stride;
  int box_cols, box_rows, swizzle, l2promo;
};

inline CUresult encode_tiled_bf16_3d(CUtensorMap* map, const TmapKey& k) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return CUDA_ERROR_NOT_FOUND;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(k.cols), static_cast<cuuint64_t>(k.rows), static_cast<cuuint64_t>(k.batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(k.ld) * 2, static_cast<cuuint64_t>(k.batch_stride) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(k.box_cols), static_cast<cuuint32_t>(k.box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(k.ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, static_cast<CUtensorMapSwizzle>(k.swizzle),
            static_cast<CUtensorMapL2promotion>(k.l2promo), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// bf16 [batch][rows][cols] view (cols contiguous, row pitch ld, batch pitch batch_stride, elements) with a
// [1][box_rows][box_cols] box.  Returns CUDA_SUCCESS or the driver's error.
inline CUresult cached_tensor_map(CUtensorMap* out, const TmapKey& key) {
  constexpr int kSlots = 2048;
  struct Slot { TmapKey key; CUtensorMap map; bool used; };
  static thread_local Slot* table = nullptr;
  if (table == nullptr) table = static_cast<Slot*>(calloc(kSlots, sizeof(Slot)));
  unsigned long long h = 1469598103934665603ULL;
  const unsigned char* kb = reinterpret_cast<const unsigned char*>(&key);
  for (size_t i = 0; i < sizeof(TmapKey); ++i) h = (h ^ kb[i]) * 1099511628211ULL;
  Slot& s = table[h % kSlots];
  if (s.used && memcmp(&s.key, &key, sizeof(TmapKey)) == 0) {
    *out = s.map;
    return CUDA_SUCCESS;
  }
  CUresult r = encode_tiled_bf16_3d(out, key);
  if (r == CUDA_SUCCESS) {
    s.key = key;
    s.map = *out;
    s.used = true;
  }
  return r;
}

inline TmapKey make_tmap_key(const void* ptr, long long cols, long long rows, long long batch, long long ld,
                             long long batch_stride, int box_cols, int box_rows, int swizzle, int l2promo) {
  TmapKey k;
  memset(&k, 0, sizeof(k));  // padding bytes take part in the hash / compare
  k.ptr = ptr; k.cols = cols; k.rows = rows; k.batch = batch; k.ld = ld; k.batch_stride = batch_stride;
  k.box_cols = box_cols; k.box_rows = box_rows; k.swizzle = swizzle; k.l2promo = l2promo;
  return k;
}

}  // namespace md
