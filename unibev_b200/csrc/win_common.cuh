// Pieces shared by the window-staged sampling kernels (win_sample.cu: fp16-staged value maps; win_sample32.cu: fp32).
#pragma once
#include "ub_tma.cuh"

namespace ub {

constexpr int kWorkerWarps = 16;                           // one query row / 16 hits of the unit each
constexpr int kWarpItems = 16;                             // items (query, head) per worker warp and unit
constexpr int kBevThreads = (kWorkerWarps + 1) * 32;       // + the scheduler warp
constexpr int kImgThreads = kWorkerWarps * 32;
constexpr int kTQ = 16;                                    // BEV tile: 16 x 16 queries
constexpr int kUnitItems = kWorkerWarps * kWarpItems;      // items per unit

constexpr size_t kSmemBudget = 232448 - 1024 - 64;  // 227 KB per CTA minus the static part


__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// P1 works with two lanes per item, each owning PPL = P / 2 consecutive sampling points.
// Softmax over the item's P logits: in-lane over the PPL own ones, one shuffle with the partner lane.
template <int PPL>
__device__ __forceinline__ void softmax_pair(const float (&lg)[PPL], bool ok, float scale, float (&aw)[PPL]) {
  float mx = lg[0];
#pragma unroll
  for (int i = 1; i < PPL; ++i) mx = fmaxf(mx, lg[i]);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    aw[i] = ex2_approx((lg[i] - mx) * 1.4426950408889634f);
    sum += aw[i];
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = ok ? rcp_approx(sum) * scale : 0.f;
#pragma unroll
  for (int i = 0; i < PPL; ++i) aw[i] *= inv;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}


__device__ __forceinline__ float round_tf32(float x) {   // round to nearest (ties away) at 10 mantissa bits
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  return __uint_as_float(t);
}
__device__ __forceinline__ void red_add4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}


struct __align__(16) UnitInfo {
  int u, b, h, tx0, ty0, wx0, wy0, pad;
};

}  // namespace ub
