// Generic multi-scale deformable attention, forward and backward -- the op-level drop-in for mmcv's
// MultiScaleDeformableAttnFunction (reference call sites: spatial_cross_attention_img.py:432-435,
// spatial_cross_attention_pts.py:439-442, decoder.py:324-327).
//
// Work item = one (b, q, h).  A group of LPG = Dh/4 adjacent lanes owns it; each lane carries four
// channels, so every corner fetch of the group is one fully used 16*LPG-byte segment (128 B at Dh=32).
// Sampling locations / weights are read once per group through the no-allocate path and broadcast by the
// LSU; the value map goes through L1 (it is the only operand with reuse).
#include "ub_common.cuh"

namespace ub {

template <int LPG>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const float* __restrict__ value,
                                                       const int64_t* __restrict__ shapes,
                                                       const int64_t* __restrict__ lvl_start,
                                                       const float* __restrict__ loc,
                                                       const float* __restrict__ aw, float* __restrict__ out,
                                                       int B, int Nv, int H, int Nq, int L, int P) {
  constexpr int Dh = LPG * 4;
  const int lane = threadIdx.x % LPG;
  const int64_t n_items = (int64_t)B * Nq * H;
  const int64_t groups_per_grid = (int64_t)gridDim.x * (blockDim.x / LPG);
  const int row = H * Dh;  // floats between neighbouring value tokens
  for (int64_t item = (int64_t)blockIdx.x * (blockDim.x / LPG) + threadIdx.x / LPG; item < n_items;
       item += groups_per_grid) {
    const int h = (int)(item % H);
    const int64_t bq = item / H;
    const int b = (int)(bq / Nq);
    const float* lp = loc + item * (int64_t)L * P * 2;
    const float* wp = aw + item * (int64_t)L * P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < L; ++l) {
      const int fH = (int)shapes[2 * l], fW = (int)shapes[2 * l + 1];
      const float* base = value + ((int64_t)b * Nv + lvl_start[l]) * row + h * Dh + lane * 4;
      for (int p = 0; p < P; ++p) {
        const float2 xy = ld_stream2(lp + (l * P + p) * 2);
        const float a = ld_stream1(wp + l * P + p);
        bilinear_acc4(acc, base, fH, fW, row, xy.y * fH - 0.5f, xy.x * fW - 0.5f, a);
      }
    }
    st_stream4(out + bq * row + h * Dh + lane * 4, acc);
  }
}

// Any Dh: one thread per (b, q, h), scalar channel loop.  Correctness path for odd head dims.
__global__ void __launch_bounds__(256) msda_fwd_scalar_kernel(const float* __restrict__ value,
                                                              const int64_t* __restrict__ shapes,
                                                              const int64_t* __restrict__ lvl_start,
                                                              const float* __restrict__ loc,
                                                              const float* __restrict__ aw, float* __restrict__ out,
                                                              int B, int Nv, int H, int Dh, int Nq, int L, int P) {
  const int64_t n_items = (int64_t)B * Nq * H * Dh;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_items;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Dh);
    const int64_t item = idx / Dh;
    const int h = (int)(item % H);
    const int b = (int)(item / H / Nq);
    const int row = H * Dh;
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
      const int fH = (int)shapes[2 * l], fW = (int)shapes[2 * l + 1];
      const float* base = value + ((int64_t)b * Nv + lvl_start[l]) * row + h * Dh + c;
      for (int p = 0; p < P; ++p) {
        const float x = loc[(item * L * P + l * P + p) * 2], y = loc[(item * L * P + l * P + p) * 2 + 1];
        const float a = aw[item * L * P + l * P + p];
        const float h_im = y * fH - 0.5f, w_im = x * fW - 0.5f;
        if (!(h_im > -1.f && w_im > -1.f && h_im < (float)fH && w_im < (float)fW)) continue;
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h0 = (int)hf, w0 = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        float v = 0.f;
        if (h0 >= 0 && w0 >= 0) v += hh * hw * base[((int64_t)h0 * fW + w0) * row];
        if (h0 >= 0 && w0 + 1 < fW) v += hh * lw * base[((int64_t)h0 * fW + w0 + 1) * row];
        if (h0 + 1 < fH && w0 >= 0) v += lh * hw * base[((int64_t)(h0 + 1) * fW + w0) * row];
        if (h0 + 1 < fH && w0 + 1 < fW) v += lh * lw * base[((int64_t)(h0 + 1) * fW + w0 + 1) * row];
        acc += a * v;
      }
    }
    out[idx] = acc;
  }
}

// ---- backward ----------------------------------------------------------------------------------------
// Same decomposition.  grad_value: one vectorised reduction (red.global.add.v4.f32, sm_90+) per corner per
// lane; grad_loc / grad_w: per-lane partial sums over its four channels, folded across the LPG lanes with
// shuffles, written once by lane 0 (each (b,q,h,l,p) is owned by exactly one group -> no atomics).
__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float4 scale4(float s, const float4& v) {
  return make_float4(s * v.x, s * v.y, s * v.z, s * v.w);
}

template <int LPG>
__global__ void __launch_bounds__(256) msda_bwd_kernel(const float* __restrict__ value,
                                                       const int64_t* __restrict__ shapes,
                                                       const int64_t* __restrict__ lvl_start,
                                                       const float* __restrict__ loc,
                                                       const float* __restrict__ aw,
                                                       const float* __restrict__ grad_out, float* grad_value,
                                                       float* __restrict__ grad_loc, float* __restrict__ grad_w,
                                                       int B, int Nv, int H, int Nq, int L, int P) {
  constexpr int Dh = LPG * 4;
  const int lane = threadIdx.x % LPG;
  const int64_t n_items = (int64_t)B * Nq * H;
  const int64_t groups_per_grid = (int64_t)gridDim.x * (blockDim.x / LPG);
  const int row = H * Dh;
  // all 32 lanes of a warp must take part in the shuffles: iterate in lock-step, predicate the work
  const int64_t n_iter = (n_items + groups_per_grid - 1) / groups_per_grid;
  int64_t item = (int64_t)blockIdx.x * (blockDim.x / LPG) + threadIdx.x / LPG;
  for (int64_t it = 0; it < n_iter; ++it, item += groups_per_grid) {
    const bool live = item < n_items;
    const int64_t item_c = live ? item : 0;
    const int h = (int)(item_c % H);
    const int64_t bq = item_c / H;
    const int b = (int)(bq / Nq);
    const float4 go = live ? ldg4(grad_out + bq * row + h * Dh + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < L; ++l) {
      const int fH = (int)shapes[2 * l], fW = (int)shapes[2 * l + 1];
      const int64_t tok0 = ((int64_t)b * Nv + lvl_start[l]) * row + h * Dh + lane * 4;
      for (int p = 0; p < P; ++p) {
        const int64_t sp = item_c * L * P + l * P + p;
        const float x = loc[sp * 2], y = loc[sp * 2 + 1], a = aw[sp];
        const float h_im = y * fH - 0.5f, w_im = x * fW - 0.5f;
        float g_a = 0.f, g_x = 0.f, g_y = 0.f;
        if (live && h_im > -1.f && w_im > -1.f && h_im < (float)fH && w_im < (float)fW) {
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int h0 = (int)hf, w0 = (int)wf;
          const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
          const bool top = h0 >= 0, bot = h0 + 1 <= fH - 1, left = w0 >= 0, right = w0 + 1 <= fW - 1;
          const int64_t o00 = tok0 + ((int64_t)h0 * fW + w0) * row;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 v1 = (top && left) ? ldg4(value + o00) : z4;
          const float4 v2 = (top && right) ? ldg4(value + o00 + row) : z4;
          const float4 v3 = (bot && left) ? ldg4(value + o00 + (int64_t)fW * row) : z4;
          const float4 v4 = (bot && right) ? ldg4(value + o00 + (int64_t)fW * row + row) : z4;
          const float4 tg = scale4(a, go);  // top_grad * attn_weight
          if (top && left) red_add4(grad_value + o00, scale4(hh * hw, tg));
          if (top && right) red_add4(grad_value + o00 + row, scale4(hh * lw, tg));
          if (bot && left) red_add4(grad_value + o00 + (int64_t)fW * row, scale4(lh * hw, tg));
          if (bot && right) red_add4(grad_value + o00 + (int64_t)fW * row + row, scale4(lh * lw, tg));
          const float d1 = dot4(tg, v1), d2 = dot4(tg, v2), d3 = dot4(tg, v3), d4 = dot4(tg, v4);
          g_y = (float)fH * (-hw * d1 - lw * d2 + hw * d3 + lw * d4);
          g_x = (float)fW * (-hh * d1 + hh * d2 - lh * d3 + lh * d4);
          g_a = hh * hw * dot4(go, v1) + hh * lw * dot4(go, v2) + lh * hw * dot4(go, v3) + lh * lw * dot4(go, v4);
        }
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) {
          g_a += __shfl_xor_sync(0xffffffffu, g_a, o);
          g_x += __shfl_xor_sync(0xffffffffu, g_x, o);
          g_y += __shfl_xor_sync(0xffffffffu, g_y, o);
        }
        if (live && lane == 0) {
          grad_w[sp] = g_a;
          grad_loc[sp * 2] = g_x;
          grad_loc[sp * 2 + 1] = g_y;
        }
      }
    }
  }
}

// Any Dh: one thread per (b, q, h, l, p); channel loop; scalar atomics on grad_value.
__global__ void __launch_bounds__(256) msda_bwd_scalar_kernel(const float* __restrict__ value,
                                                              const int64_t* __restrict__ shapes,
                                                              const int64_t* __restrict__ lvl_start,
                                                              const float* __restrict__ loc,
                                                              const float* __restrict__ aw,
                                                              const float* __restrict__ grad_out, float* grad_value,
                                                              float* __restrict__ grad_loc,
                                                              float* __restrict__ grad_w, int B, int Nv, int H, int Dh,
                                                              int Nq, int L, int P) {
  const int64_t n = (int64_t)B * Nq * H * L * P;
  const int row = H * Dh;
  for (int64_t sp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sp < n; sp += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)((sp / P) % L);
    const int64_t item = sp / ((int64_t)L * P);
    const int h = (int)(item % H);
    const int64_t bq = item / H;
    const int b = (int)(bq / Nq);
    const int fH = (int)shapes[2 * l], fW = (int)shapes[2 * l + 1];
    const float x = loc[sp * 2], y = loc[sp * 2 + 1], a = aw[sp];
    const float h_im = y * fH - 0.5f, w_im = x * fW - 0.5f;
    float g_a = 0.f, g_x = 0.f, g_y = 0.f;
    if (h_im > -1.f && w_im > -1.f && h_im < (float)fH && w_im < (float)fW) {
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int h0 = (int)hf, w0 = (int)wf;
      const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
      const bool top = h0 >= 0, bot = h0 + 1 <= fH - 1, left = w0 >= 0, right = w0 + 1 <= fW - 1;
      const int64_t o00 = ((int64_t)b * Nv + lvl_start[l] + (int64_t)h0 * fW + w0) * row + h * Dh;
      for (int c = 0; c < Dh; ++c) {
        const float go = grad_out[bq * row + h * Dh + c], tg = a * go;
        const float v1 = (top && left) ? value[o00 + c] : 0.f;
        const float v2 = (top && right) ? value[o00 + row + c] : 0.f;
        const float v3 = (bot && left) ? value[o00 + (int64_t)fW * row + c] : 0.f;
        const float v4 = (bot && right) ? value[o00 + (int64_t)fW * row + row + c] : 0.f;
        if (top && left) atomicAdd(grad_value + o00 + c, hh * hw * tg);
        if (top && right) atomicAdd(grad_value + o00 + row + c, hh * lw * tg);
        if (bot && left) atomicAdd(grad_value + o00 + (int64_t)fW * row + c, lh * hw * tg);
        if (bot && right) atomicAdd(grad_value + o00 + (int64_t)fW * row + row + c, lh * lw * tg);
        g_y += (float)fH * tg * (-hw * v1 - lw * v2 + hw * v3 + lw * v4);
        g_x += (float)fW * tg * (-hh * v1 + hh * v2 - lh * v3 + lh * v4);
        g_a += go * (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4);
      }
    }
    grad_w[sp] = g_a;
    grad_loc[sp * 2] = g_x;
    grad_loc[sp * 2 + 1] = g_y;
  }
}

static int grid_for(int64_t n_groups, int groups_per_block) {
  int64_t blocks = (n_groups + groups_per_block - 1) / groups_per_block;
  const int64_t cap = (int64_t)sm_count() * 32;  // grid-stride beyond 32 resident waves' worth of CTAs
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace ub

using namespace ub;

static int check_msda_args(const char* fn, int B, int Nv, int H, int D, int Nq, int L, int P) {
  UB_REQUIRE(B > 0 && Nv > 0 && H > 0 && D > 0 && Nq > 0 && L > 0 && P > 0,
             "%s: all of B,Nv,H,D,Nq,L,P must be positive (got %d,%d,%d,%d,%d,%d,%d)", fn, B, Nv, H, D, Nq, L, P);
  return UB_OK;
}

extern "C" int ub_msda_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                           const float* sampling_loc, const float* attn_weight, float* out, int B, int Nv, int H,
                           int D, int Nq, int L, int P, ub_stream_t stream) {
  if (int rc = check_msda_args("ub_msda_fwd", B, Nv, H, D, Nq, L, P)) return rc;
  UB_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
             "ub_msda_fwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t items = (int64_t)B * Nq * H;
  const bool vec_ok = (D % 4 == 0) && ((D / 4) & (D / 4 - 1)) == 0 && D <= 128 &&
                      (reinterpret_cast<uintptr_t>(value) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0;
  if (vec_ok) {
#define UB_FWD(LPG)                                                                                              \
  msda_fwd_kernel<LPG><<<grid_for(items, 256 / LPG), 256, 0, s>>>(value, spatial_shapes, level_start_index,      \
                                                                  sampling_loc, attn_weight, out, B, Nv, H, Nq, L, P)
    switch (D / 4) {
      case 1: UB_FWD(1); break;
      case 2: UB_FWD(2); break;
      case 4: UB_FWD(4); break;
      case 8: UB_FWD(8); break;
      case 16: UB_FWD(16); break;
      default: UB_FWD(32); break;
    }
#undef UB_FWD
  } else {
    msda_fwd_scalar_kernel<<<grid_for(items * D, 256), 256, 0, s>>>(value, spatial_shapes, level_start_index,
                                                                    sampling_loc, attn_weight, out, B, Nv, H, D, Nq, L,
                                                                    P);
  }
  return check_launch("ub_msda_fwd");
}

extern "C" int ub_msda_bwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                           const float* sampling_loc, const float* attn_weight, const float* grad_out,
                           float* grad_value, float* grad_loc, float* grad_w, int B, int Nv, int H, int D, int Nq,
                           int L, int P, ub_stream_t stream) {
  if (int rc = check_msda_args("ub_msda_bwd", B, Nv, H, D, Nq, L, P)) return rc;
  UB_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && grad_out && grad_value &&
                 grad_loc && grad_w,
             "ub_msda_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t items = (int64_t)B * Nq * H;
  const bool vec_ok = (D % 4 == 0) && ((D / 4) & (D / 4 - 1)) == 0 && D <= 128 &&
                      (reinterpret_cast<uintptr_t>(value) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(grad_value) & 15u) == 0;
  if (vec_ok) {
#define UB_BWD(LPG)                                                                                              \
  msda_bwd_kernel<LPG><<<grid_for(items, 256 / LPG), 256, 0, s>>>(value, spatial_shapes, level_start_index,      \
                                                                  sampling_loc, attn_weight, grad_out, grad_value, \
                                                                  grad_loc, grad_w, B, Nv, H, Nq, L, P)
    switch (D / 4) {
      case 1: UB_BWD(1); break;
      case 2: UB_BWD(2); break;
      case 4: UB_BWD(4); break;
      case 8: UB_BWD(8); break;
      case 16: UB_BWD(16); break;
      default: UB_BWD(32); break;
    }
#undef UB_BWD
  } else {
    msda_bwd_scalar_kernel<<<grid_for(items * L * P, 256), 256, 0, s>>>(value, spatial_shapes, level_start_index,
                                                                        sampling_loc, attn_weight, grad_out,
                                                                        grad_value, grad_loc, grad_w, B, Nv, H, D, Nq,
                                                                        L, P);
  }
  return check_launch("ub_msda_bwd");
}
