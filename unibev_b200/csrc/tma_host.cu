// Host side of ub_tma.cuh: cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup, with a
// small cache (the encoder is called with the same few tensors every frame).
#include <cstring>
#include <mutex>
#include <vector>

#include "ub_tma.cuh"

namespace ub {

namespace {
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Key {
  int dtype, rank, swizzle;
  const void* base;
  uint64_t dims[5], strides[5];
  uint32_t box[5];
};
struct Entry {
  Key key;
  alignas(64) CUtensorMap map;
};
std::mutex g_mu;
std::vector<Entry> g_cache;
size_t g_next = 0;
constexpr size_t kCacheSize = 128;

EncodeFn encoder() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}
}  // namespace

int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  Key k;
  memset(&k, 0, sizeof(k));
  k.dtype = (int)dtype, k.rank = rank, k.swizzle = (int)swizzle, k.base = base;
  for (int i = 0; i < rank; ++i) k.dims[i] = dims[i], k.box[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides_bytes[i];
  std::lock_guard<std::mutex> lock(g_mu);
  for (const Entry& e : g_cache)
    if (memcmp(&e.key, &k, sizeof(Key)) == 0) {
      *out = e.map;
      return UB_OK;
    }
  EncodeFn fn = encoder();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return UB_ECUDA;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i], gbox[i] = box[i], estr[i] = 1;
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  alignas(64) CUtensorMap m;
  const CUresult r = fn(&m, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, gbox, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return UB_ECUDA;
  }
  Entry e;
  e.key = k;
  e.map = m;
  if (g_cache.size() < kCacheSize)
    g_cache.push_back(e);
  else
    g_cache[g_next++ % kCacheSize] = e;
  *out = m;
  return UB_OK;
}

}  // namespace ub
