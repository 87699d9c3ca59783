// Shared device/host helpers for libunibev_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "unibev_b200.h"

namespace ub {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs (documentation; launch code asks the device: sm_count())

// Per-device facts and kernel attributes (a process may drive several GPUs): SM count of the CURRENT device, and the
// dynamic-shared-memory opt-in of `kernel` on the current device raised to at least `smem` bytes (cudaFuncSetAttribute is
// per device; remembered per (kernel, device)).  Callers make the tensors' device current before calling into the library.
int sm_count();
bool smem_config_needed(const void* kernel, size_t smem);   // true: not yet configured for `smem` on the current device
void smem_config_done(const void* kernel, size_t smem);

void set_error(const char* fmt, ...);
void count_launch();
int unsupported();   // counts the event (ub_unsupported_count) and returns UB_EUNSUPPORTED

#define UB_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      ub::set_error(__VA_ARGS__);             \
      return UB_EINVAL;                       \
    }                                         \
  } while (0)

#define UB_REQUIRE_ALIGNED16(p)                                         \
  do {                                                                  \
    if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) {                  \
      ub::set_error("%s: pointer %s is not 16-byte aligned", __func__, #p); \
      return UB_EALIGN;                                                 \
    }                                                                   \
  } while (0)

// Programmatic dependent launch (see pdl_trigger / pdl_wait below): on unless ub_set_pdl(0).
bool pdl_enabled();
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-serialization attribute when enabled
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename K>
inline int ensure_smem(K kernel, size_t smem, const char* fn) {
  const void* key = reinterpret_cast<const void*>(kernel);
  if (smem <= 48 * 1024 || !smem_config_needed(key, smem)) return UB_OK;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("%s: cannot reserve %zu bytes of shared memory", fn, smem);
    cudaGetLastError();
    return UB_ECUDA;
  }
  smem_config_done(key, smem);
  return UB_OK;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UB_ECUDA;
  }
  count_launch();
  return UB_OK;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (read-once / write-once) accesses: keep them out of L1 so the gathered value map stays resident
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// 32 contiguous bytes in one request (LDG.256, sm_100+): p must be 32-byte aligned
__device__ __forceinline__ void ld_stream8(const float* p, float (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ float2 ld_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream is still draining.  pdl_trigger() lets the successor's CTAs be scheduled as soon as SMs
// free up; pdl_wait() blocks until the predecessor grid has completed and its memory is visible -- call it before the
// first access to anything a predecessor may have written (or may still be reading, for buffers this kernel overwrites).
// Both are no-ops for kernels launched without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}

// One bilinear sample of a (fH, fW, stride) fp32 map at pixel coords (h_im, w_im), mmcv semantics:
// contributes iff h_im > -1 && w_im > -1 && h_im < fH && w_im < fW; each corner bounds-checked.
// `base` points at channel slice of pixel (0,0); `pix_stride` floats between neighbouring pixels.
__device__ __forceinline__ void bilinear_acc4(float4& acc, const float* __restrict__ base, int fH, int fW,
                                              int pix_stride, float h_im, float w_im, float aw) {
  if (!(h_im > -1.f && w_im > -1.f && h_im < (float)fH && w_im < (float)fW)) return;
  const float hf = floorf(h_im), wf = floorf(w_im);
  const int h0 = (int)hf, w0 = (int)wf;
  const float lh = h_im - hf, lw = w_im - wf;
  const float hh = 1.f - lh, hw = 1.f - lw;
  const bool top = h0 >= 0, bot = h0 + 1 <= fH - 1, left = w0 >= 0, right = w0 + 1 <= fW - 1;
  const float* p00 = base + ((int64_t)h0 * fW + w0) * pix_stride;
  float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
  if (top && left) v1 = ldg4(p00);
  if (top && right) v2 = ldg4(p00 + pix_stride);
  if (bot && left) v3 = ldg4(p00 + (int64_t)fW * pix_stride);
  if (bot && right) v4 = ldg4(p00 + (int64_t)fW * pix_stride + pix_stride);
  fma4(acc, aw * (hh * hw), v1);
  fma4(acc, aw * (hh * lw), v2);
  fma4(acc, aw * (lh * hw), v3);
  fma4(acc, aw * (lh * lw), v4);
}

}  // namespace ub
