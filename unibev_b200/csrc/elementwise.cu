// Bandwidth-bound glue of the BEV encoder: residual+LayerNorm [R5], CNW fusion [R6], feature flatten [R7].
// Each is one pass: 128-bit accesses along the channel dim, no intermediate tensors.
#include <cuda_fp16.h>

#include "ub_common.cuh"

namespace ub {

// One warp per row; lane i owns float4 chunks i, i+32, ...  (NV chunks per lane, C <= 128*NV).
template <int NV>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                                            const float* __restrict__ res,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ out,
                                                            uint2* __restrict__ out16, int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)C;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) {
        v[k] = ld_stream4(x + r * C + c);
        if (bias) {
          const float4 t = ldg4(bias + c);
          v[k].x += t.x, v[k].y += t.y, v[k].z += t.z, v[k].w += t.w;
        }
        if (res) {
          const float4 t = ld_stream4(res + r * C + c);
          v[k].x += t.x, v[k].y += t.y, v[k].z += t.z, v[k].w += t.w;
        }
        sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < C) {
        const float a = v[k].x - mean, b2 = v[k].y - mean, c2 = v[k].z - mean, d = v[k].w - mean;
        sq += (a * a + b2 * b2) + (c2 * c2 + d * d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < C) {
        const float4 g = ldg4(gamma + c), bt = ldg4(beta + c);
        float4 o;
        o.x = (v[k].x - mean) * rstd * g.x + bt.x;
        o.y = (v[k].y - mean) * rstd * g.y + bt.y;
        o.z = (v[k].z - mean) * rstd * g.z + bt.z;
        o.w = (v[k].w - mean) * rstd * g.w + bt.w;
        st_stream4(out + r * C + c, o);
        if (out16) {   // fp16 copy of the row: the A operand of the projections that read it next
          const float lim = 65504.f;
          const __half2 h0 = __floats2half2_rn(fminf(fmaxf(o.x, -lim), lim), fminf(fmaxf(o.y, -lim), lim));
          const __half2 h1 = __floats2half2_rn(fminf(fmaxf(o.z, -lim), lim), fminf(fmaxf(o.w, -lim), lim));
          uint2 p;
          p.x = *reinterpret_cast<const uint32_t*>(&h0), p.y = *reinterpret_cast<const uint32_t*>(&h1);
          out16[(r * C + c) >> 2] = p;
        }
      }
    }
  }
}

struct FuseParams {
  int mode, c_flag, l_flag, rows_per_item, C;
};

// thread = one float4 of channels of one row.  Per-channel CNW weights are recomputed per thread (2 exps).
__global__ void __launch_bounds__(1024) cnw_fuse_kernel(const float* __restrict__ img, const float* __restrict__ pts,
                                                       const float* __restrict__ w_img, const float* __restrict__ w_pts,
                                                       const float* __restrict__ s_img, const float* __restrict__ s_pts,
                                                       const float* __restrict__ modal, float* __restrict__ out,
                                                       int64_t rows, FuseParams fp) {
  const int C4 = fp.C / 4;
  const float cf = (float)fp.c_flag, lf = (float)fp.l_flag;
  // blockDim is a multiple of C4 (host): a thread keeps its channel quad for the whole launch, so the per-channel CNW
  // softmax is evaluated once per thread instead of once per element, and no 64-bit div / mod runs in the loop
  const int c = ((int)threadIdx.x % C4) * 4;
  const int rows_per_block = (int)blockDim.x / C4;
  float wi[4] = {1.f, 1.f, 1.f, 1.f}, wp[4] = {1.f, 1.f, 1.f, 1.f};
  if (w_img && fp.c_flag == 1 && fp.l_flag == 1) {  // feature_norm == 'ChannelNormWeights'; a single row softmaxes to 1
    const float4 a = ldg4(w_img + c), b = ldg4(w_pts + c);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float m = fmaxf(av[k], bv[k]);
      const float ea = expf(av[k] - m), eb = expf(bv[k] - m), s = ea + eb;
      wi[k] = ea / s, wp[k] = eb / s;
    }
  }
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + (int)threadIdx.x / C4; r < rows;
       r += (int64_t)gridDim.x * rows_per_block) {
    float si = 1.f, sp = 1.f;
    if (s_img) {  // spatial_norm == 'SpatialNormWeights'
      const int q = (int)(r % fp.rows_per_item);
      if (fp.c_flag == 1 && fp.l_flag == 1) {
        const float a = __ldg(s_img + q), b = __ldg(s_pts + q), m = fmaxf(a, b);
        const float ea = expf(a - m), eb = expf(b - m), s = ea + eb;
        si = ea / s, sp = eb / s;
      }
    }
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a4;
    if (img) a4 = ld_stream4(img + r * fp.C + c);
    if (pts) b4 = ld_stream4(pts + r * fp.C + c);
    float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w}, o[4], o2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x = __fmul_rn(a[k], wi[k]), y = __fmul_rn(b[k], wp[k]);
      if (s_img) x = __fmul_rn(x, si), y = __fmul_rn(y, sp);
      if (fp.mode == UB_FUSE_LINEAR) {
        o[k] = __fadd_rn(__fmul_rn(cf, x), __fmul_rn(lf, y));
      } else if (fp.mode == UB_FUSE_AVG) {
        const float den = cf + lf;
        o[k] = __fadd_rn(__fdiv_rn(__fmul_rn(x, cf), den), __fdiv_rn(__fmul_rn(y, lf), den));
      } else {
        o[k] = __fmul_rn(x, cf), o2[k] = __fmul_rn(y, lf);
      }
    }
    if (fp.mode == UB_FUSE_CAT) {
      if (modal) {
        const float4 m1 = ldg4(modal + c), m2 = ldg4(modal + fp.C + c);
        o[0] += m1.x, o[1] += m1.y, o[2] += m1.z, o[3] += m1.w;
        o2[0] += m2.x, o2[1] += m2.y, o2[2] += m2.z, o2[3] += m2.w;
      }
      st_stream4(out + r * 2 * fp.C + c, make_float4(o[0], o[1], o[2], o[3]));
      st_stream4(out + r * 2 * fp.C + fp.C + c, make_float4(o2[0], o2[1], o2[2], o2[3]));
    } else {
      if (modal) {
        const float4 m1 = ldg4(modal + c);
        o[0] += m1.x, o[1] += m1.y, o[2] += m1.z, o[3] += m1.w;
      }
      st_stream4(out + r * fp.C + c, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

// (G, C, HW) -> (G, HW, C) through a 32x33 shared tile; coalesced on both sides.
__global__ void __launch_bounds__(256) flatten_feats_kernel(const float* __restrict__ in,
                                                            const float* __restrict__ embed_a, int n_a,
                                                            const float* __restrict__ embed_b, float* __restrict__ out,
                                                            __half* __restrict__ out16, int C, int HW,
                                                            float* __restrict__ absmax) {
  __shared__ float tile[32][33];
  pdl_trigger();
  const int g = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = in + (int64_t)g * C * HW;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, p = p0 + tx;
    tile[ty + j][tx] = (c < C && p < HW) ? __ldg(src + (int64_t)c * HW + p) : 0.f;
  }
  __syncthreads();
  const int64_t dst0 = (int64_t)g * HW * C;
  const int c = c0 + tx;
  float e1 = 0.f, e2 = 0.f;
  if (c < C) {
    if (embed_a) e1 = __ldg(embed_a + (int64_t)(g % n_a) * C + c);
    if (embed_b) e2 = __ldg(embed_b + c);
  }
  float mx = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int p = p0 + ty + j;
    if (c < C && p < HW) {
      float v = tile[tx][ty + j];
      if (embed_a) v = __fadd_rn(v, e1);
      if (embed_b) v = __fadd_rn(v, e2);
      mx = fmaxf(mx, fabsf(v));
      if (out) out[dst0 + (int64_t)p * C + c] = v;
      if (out16) out16[dst0 + (int64_t)p * C + c] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    }
  }
  if (absmax) {   // largest magnitude written (non-negative floats order like their bit patterns)
    // One atomic per BLOCK at most, and only when it would raise the maximum: every atomic of the launch targets the same
    // address and they serialise in its L2 slice (one per warp made this kernel 4x slower than its copy).
    __shared__ float s_mx[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (tx == 0) s_mx[ty] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_mx[i]);
      if (mx > __ldcg(absmax)) atomicMax(reinterpret_cast<int*>(absmax), __float_as_int(mx));
    }
  }
}

// fp16-only output (the fused fp16 pipeline): 64 channels x 32 pixels per block, a lane converts two neighbouring
// channels, so every store instruction writes full 128-byte row segments.
__global__ void __launch_bounds__(256) flatten_feats_h_kernel(const float* __restrict__ in,
                                                              const float* __restrict__ embed_a, int n_a,
                                                              const float* __restrict__ embed_b,
                                                              __half* __restrict__ out16, int C, int HW) {
  __shared__ float tile[64][33];
  pdl_trigger();
  const int g = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = in + (int64_t)g * C * HW;
#pragma unroll
  for (int j = 0; j < 64; j += 8) {
    const int c = c0 + ty + j, p = p0 + tx;
    tile[ty + j][tx] = (c < C && p < HW) ? __ldg(src + (int64_t)c * HW + p) : 0.f;
  }
  __syncthreads();
  const int c = c0 + 2 * tx;     // this lane's channel pair (C is even)
  if (c >= C) return;
  float ea[2] = {0.f, 0.f}, eb[2] = {0.f, 0.f};
  if (embed_a) ea[0] = __ldg(embed_a + (int64_t)(g % n_a) * C + c), ea[1] = __ldg(embed_a + (int64_t)(g % n_a) * C + c + 1);
  if (embed_b) eb[0] = __ldg(embed_b + c), eb[1] = __ldg(embed_b + c + 1);
  __half* dst = out16 + (int64_t)g * HW * C + c;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int p = p0 + ty + j;
    if (p < HW) {
      float v0 = tile[2 * tx][ty + j], v1 = tile[2 * tx + 1][ty + j];
      if (embed_a) v0 = __fadd_rn(v0, ea[0]), v1 = __fadd_rn(v1, ea[1]);
      if (embed_b) v0 = __fadd_rn(v0, eb[0]), v1 = __fadd_rn(v1, eb[1]);
      const float lim = 65504.f;
      *reinterpret_cast<__half2*>(dst + (int64_t)p * C) =
          __floats2half2_rn(fminf(fmaxf(v0, -lim), lim), fminf(fmaxf(v1, -lim), lim));
    }
  }
}

// src (rows, C) -> out32 / out16 (B, rows, C): the BEV query table repeated for every sample of the batch
// (transformer_fusion.py:493-498 `bev_queries.unsqueeze(1).repeat(1, bs, 1)`), with the fp16 copy the first
// projections read.  Thread = 8 channels of one source row.
__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float* __restrict__ src, int64_t n8, int B,
                                                             float* __restrict__ out32, __half* __restrict__ out16) {
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 h;
    const float lim = 65504.f;
    auto pk = [lim](float x, float y) {
      const __half2 t = __floats2half2_rn(fminf(fmaxf(x, -lim), lim), fminf(fmaxf(y, -lim), lim));
      return *reinterpret_cast<const uint32_t*>(&t);
    };
    h.x = pk(a.x, a.y), h.y = pk(a.z, a.w), h.z = pk(b.x, b.y), h.w = pk(b.z, b.w);
    for (int r = 0; r < B; ++r) {
      if (out32)   // 32 bytes in one store (STG.256)
        asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out32 + (r * n8 + i) * 8),
                     "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                     : "memory");
      if (out16) reinterpret_cast<uint4*>(out16)[r * n8 + i] = h;
    }
  }
}

}  // namespace ub

using namespace ub;

extern "C" int ub_add_layernorm(const float* x, const float* bias, const float* residual, const float* gamma,
                                const float* beta, float* out, int64_t rows, int C, float eps, ub_stream_t stream) {
  return ub_add_layernorm16(x, bias, residual, gamma, beta, out, nullptr, rows, C, eps, stream);
}

extern "C" int ub_add_layernorm16(const float* x, const float* bias, const float* residual, const float* gamma,
                                  const float* beta, float* out, void* out16, int64_t rows, int C, float eps,
                                  ub_stream_t stream) {
  UB_REQUIRE(x && gamma && beta && out, "ub_add_layernorm: null pointer");
  UB_REQUIRE((reinterpret_cast<uintptr_t>(out16) & 7u) == 0, "ub_add_layernorm: out16 not 8-byte aligned");
  UB_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= 1024, "ub_add_layernorm: need rows>0, C%%4==0, C<=1024 (C=%d)", C);
  UB_REQUIRE_ALIGNED16(x);
  UB_REQUIRE_ALIGNED16(out);
  UB_REQUIRE_ALIGNED16(gamma);
  UB_REQUIRE_ALIGNED16(beta);
  if (bias) UB_REQUIRE_ALIGNED16(bias);
  if (residual) UB_REQUIRE_ALIGNED16(residual);
  int64_t blocks = (rows + 7) / 8;
  if (blocks > sm_count() * 32) blocks = sm_count() * 32;
  cudaStream_t s = (cudaStream_t)stream;
  const int nv = (C + 127) / 128;
#define UB_LN(NV) add_layernorm_kernel<NV><<<(int)blocks, 256, 0, s>>>(x, bias, residual, gamma, beta, out, reinterpret_cast<uint2*>(out16), rows, C, eps)
  if (nv <= 1) UB_LN(1);
  else if (nv <= 2) UB_LN(2);
  else if (nv <= 4) UB_LN(4);
  else UB_LN(8);
#undef UB_LN
  return check_launch("ub_add_layernorm");
}

extern "C" int ub_cnw_fuse(const float* img, const float* pts, const float* w_img, const float* w_pts,
                           const float* s_img, const float* s_pts, const float* modal_embed, float* out, int64_t rows,
                           int rows_per_item, int C, int mode, int c_flag, int l_flag, ub_stream_t stream) {
  UB_REQUIRE(out && (img || pts), "ub_cnw_fuse: need an output and at least one modality");
  UB_REQUIRE(rows > 0 && rows_per_item > 0 && C > 0 && C % 4 == 0, "ub_cnw_fuse: bad shape (rows=%lld C=%d)",
             (long long)rows, C);
  UB_REQUIRE(mode == UB_FUSE_LINEAR || mode == UB_FUSE_AVG || mode == UB_FUSE_CAT, "ub_cnw_fuse: unknown mode %d", mode);
  UB_REQUIRE((w_img == nullptr) == (w_pts == nullptr), "ub_cnw_fuse: w_img and w_pts must both be set or both be NULL");
  UB_REQUIRE((s_img == nullptr) == (s_pts == nullptr), "ub_cnw_fuse: s_img and s_pts must both be set or both be NULL");
  UB_REQUIRE((c_flag == 0 || c_flag == 1) && (l_flag == 0 || l_flag == 1), "ub_cnw_fuse: flags must be 0/1");
  UB_REQUIRE(!(mode == UB_FUSE_AVG && c_flag + l_flag == 0), "ub_cnw_fuse: avg fusion with both flags 0");
  if (img) UB_REQUIRE_ALIGNED16(img);
  if (pts) UB_REQUIRE_ALIGNED16(pts);
  UB_REQUIRE_ALIGNED16(out);
  if (w_img) { UB_REQUIRE_ALIGNED16(w_img); UB_REQUIRE_ALIGNED16(w_pts); }
  if (modal_embed) UB_REQUIRE_ALIGNED16(modal_embed);
  FuseParams fp{mode, c_flag, l_flag, rows_per_item, C};
  // threads per block: the largest multiple of C / 4 that fits 256 (a thread keeps one channel quad), C / 4 <= 1024
  const int C4 = C / 4;
  UB_REQUIRE(C4 <= 1024, "ub_cnw_fuse: C = %d too large (max 4096)", C);
  const int threads = C4 >= 256 ? C4 : (256 / C4) * C4;
  const int rows_per_block = threads / C4;
  int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  if (blocks > sm_count() * 32) blocks = sm_count() * 32;
  cnw_fuse_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(img, pts, w_img, w_pts, s_img, s_pts, modal_embed, out,
                                                                 rows, fp);
  return check_launch("ub_cnw_fuse");
}

extern "C" int ub_flatten_feats16(const float* in, const float* embed_a, int n_a, const float* embed_b, float* out,
                                  void* out16, int G, int C, int HW, ub_stream_t stream) {
  UB_REQUIRE(in && (out || out16), "ub_flatten_feats: null pointer");
  UB_REQUIRE(G > 0 && G <= 65535 && C > 0 && HW > 0, "ub_flatten_feats: bad shape (G=%d C=%d HW=%d)", G, C, HW);
  UB_REQUIRE(embed_a == nullptr || n_a > 0, "ub_flatten_feats: embed_a given with n_a=%d", n_a);
  if (!out && C % 2 == 0 && (reinterpret_cast<uintptr_t>(out16) & 3u) == 0) {
    dim3 grid((HW + 31) / 32, (C + 63) / 64, G);
    flatten_feats_h_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, embed_a, n_a > 0 ? n_a : 1, embed_b,
                                                                   reinterpret_cast<__half*>(out16), C, HW);
    return check_launch("ub_flatten_feats");
  }
  dim3 grid((HW + 31) / 32, (C + 31) / 32, G);
  flatten_feats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, embed_a, n_a > 0 ? n_a : 1, embed_b, out,
                                                               reinterpret_cast<__half*>(out16), C, HW, nullptr);
  return check_launch("ub_flatten_feats");
}

// ub_flatten_feats that also raises *absmax (a device float the caller zeroed) to the largest magnitude it wrote: the
// device-side operand bound of ub_linear_f16x3 for rows that come straight from an input tensor.
extern "C" int ub_flatten_feats_max(const float* in, const float* embed_a, int n_a, const float* embed_b, float* out,
                                    float* absmax, int G, int C, int HW, ub_stream_t stream) {
  UB_REQUIRE(in && out && absmax, "ub_flatten_feats_max: null pointer");
  UB_REQUIRE(G > 0 && G <= 65535 && C > 0 && HW > 0, "ub_flatten_feats_max: bad shape (G=%d C=%d HW=%d)", G, C, HW);
  UB_REQUIRE(embed_a == nullptr || n_a > 0, "ub_flatten_feats_max: embed_a given with n_a=%d", n_a);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, G);
  flatten_feats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, embed_a, n_a > 0 ? n_a : 1, embed_b, out, nullptr, C, HW,
                                                               absmax);
  return check_launch("ub_flatten_feats_max");
}

extern "C" int ub_flatten_feats(const float* in, const float* embed_a, int n_a, const float* embed_b, float* out, int G,
                                int C, int HW, ub_stream_t stream) {
  UB_REQUIRE(out, "ub_flatten_feats: null pointer");
  return ub_flatten_feats16(in, embed_a, n_a, embed_b, out, nullptr, G, C, HW, stream);
}

extern "C" int ub_broadcast_rows(const float* src, int64_t rows, int C, int B, float* out32, void* out16,
                                 ub_stream_t stream) {
  UB_REQUIRE(src && (out32 || out16), "ub_broadcast_rows: null pointer");
  UB_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && B > 0, "ub_broadcast_rows: need rows > 0, C %% 8 == 0, B > 0");
  UB_REQUIRE_ALIGNED16(src);
  UB_REQUIRE(!out32 || (reinterpret_cast<uintptr_t>(out32) & 31u) == 0, "ub_broadcast_rows: out32 must be 32-byte aligned");
  if (out16) UB_REQUIRE_ALIGNED16(out16);
  const int64_t n8 = rows * C / 8;
  int blocks = (int)((n8 + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  broadcast_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, n8, B, out32, reinterpret_cast<__half*>(out16));
  return check_launch("ub_broadcast_rows");
}
