// Streaming backward helpers of the training step (BASELINE configs[4]): the row-reductions autograd needs around the
// projections and the 'norm' steps of the encoder layers (encoder_unibev_detr_img.py:434-436,476-479).
//
//   ub_colsum          out[n] += sum_m x[m, n]                      bias gradient of a projection (M = B * 40 000 rows)
//   ub_layernorm_bwd   dx, dgamma += , dbeta +=                     backward of y = LayerNorm(x [+ residual]) * gamma + beta
//   ub_dropout_add_layernorm_fwd / _bwd                             y = LayerNorm(dropout(x) + residual) * gamma + beta: the
//                      `dropout(out) + identity` of an attention / FFN and the 'norm' step after it in ONE pass per direction
//                      (three kernels and two extra round trips of the activation otherwise); the keep mask is drawn in
//                      the kernel from a counter-based generator and kept as 4 bits per 16 bytes for the backward pass
//
// Both are one pass over their inputs with 128-bit accesses (HBM-bound: 4 bytes per element read, LayerNorm 12 bytes per
// element moved); per-block partial column sums are folded into the (pre-zeroed) outputs with red.global.add.
// torch's generic column reduction / LayerNorm kernels take 47-116 us for these (80 000, 128) tensors; a full pass at the
// measured HBM rate is 7-20 us.
#include "ub_common.cuh"

namespace ub {

// thread = one float4 column group x one row lane; block = RL row lanes x CG column groups (CG = N / 4 <= 256)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t M, int N) {
  __shared__ float4 s_part[256];
  const int CG = N >> 2, RL = (int)blockDim.x / CG;
  const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < RL) {
    const int64_t step = (int64_t)gridDim.x * RL;
    int64_t r = (int64_t)blockIdx.x * RL + rl;
    // two rows in flight per thread
    for (; r + step < M; r += 2 * step) {
      const float4 a = ld_stream4(x + r * N + cg * 4), b = ld_stream4(x + (r + step) * N + cg * 4);
      acc.x += a.x + b.x, acc.y += a.y + b.y, acc.z += a.z + b.z, acc.w += a.w + b.w;
    }
    if (r < M) {
      const float4 a = ld_stream4(x + r * N + cg * 4);
      acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
    }
  }
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < RL; ++k) {
      const float4 t = s_part[k * CG + cg];
      acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + cg * 4), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w)
                 : "memory");
  }
}

// Counter-based keep mask: one 64-bit mix (splitmix64 finaliser) of (key, float4 index) yields four 16-bit uniform
// fields, one per element; element j is KEPT iff field j >= thr (thr = round(p 65536)).  key = mix(seed, step, call site):
// the caller advances `step` once per training step and numbers the call sites, so no (step, site, element) repeats.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t dropout_key(const int64_t* __restrict__ rng, int salt) {
  const uint64_t seed = (uint64_t)__ldg(rng), step = (uint64_t)__ldg(rng + 1);
  return mix64(seed ^ mix64(step * 0x9e3779b97f4a7c15ull + (uint64_t)(uint32_t)salt));
}
__device__ __forceinline__ unsigned keep_bits(uint64_t key, uint64_t idx4, unsigned thr) {
  const uint64_t r = mix64(key + idx4 * 0x9e3779b97f4a7c15ull);
  return ((unsigned)(r & 0xffffu) >= thr ? 1u : 0u) | ((unsigned)((r >> 16) & 0xffffu) >= thr ? 2u : 0u) |
         ((unsigned)((r >> 32) & 0xffffu) >= thr ? 4u : 0u) | ((unsigned)(r >> 48) >= thr ? 8u : 0u);
}

// y = LayerNorm(keep * x * scale + res) * gamma + beta; mask[(r C + c) / 4] = keep bits of the float4 at (r, c)
template <int NV>
__global__ void __launch_bounds__(256) dropout_add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float* __restrict__ out,
                                                                    uint8_t* __restrict__ mask, int64_t rows, int C, float eps,
                                                                    unsigned thr, float scale, const int64_t* __restrict__ rng,
                                                                    int salt) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)C;
  const uint64_t key = dropout_key(rng, salt);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) {
        const int64_t i4 = (r * C + c) >> 2;
        const unsigned kb = keep_bits(key, (uint64_t)i4, thr);
        mask[i4] = (uint8_t)kb;
        const float4 t = ld_stream4(x + r * C + c), q = ld_stream4(res + r * C + c);
        v[k].x = ((kb & 1u) ? t.x * scale : 0.f) + q.x, v[k].y = ((kb & 2u) ? t.y * scale : 0.f) + q.y;
        v[k].z = ((kb & 4u) ? t.z * scale : 0.f) + q.z, v[k].w = ((kb & 8u) ? t.w * scale : 0.f) + q.w;
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
      }
    }
  }
}

// One warp per row; lane i owns float4 chunks i, i + 32, ... (NV chunks per lane, C <= 128 NV).  Row statistics are
// recomputed from x (a row lives in registers), so the forward pass saves nothing but its input.
//   xhat = (x - mean) rstd;  g = gamma dy;  dx = rstd (g - mean(g) - xhat mean(g xhat));  dgamma += dy xhat;  dbeta += dy
template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                            const float* __restrict__ dy,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                                                            int C, float eps, const uint8_t* __restrict__ mask, float scale,
                                                            float* __restrict__ dx_drop) {
  __shared__ float4 s_red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)C;
  float4 gm[NV], ag[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    gm[k] = c < C ? ldg4(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[k] = ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; r < rows; r += warps) {
    float4 v[NV], d[NV];
    unsigned kb[NV];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      v[k] = d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      kb[k] = 15u;
      if (c < C) {
        v[k] = ld_stream4(x + r * C + c);
        if (mask) {   // the normalised row was dropout(x) + residual with the saved keep bits
          kb[k] = mask[(r * C + c) >> 2];
          v[k].x = (kb[k] & 1u) ? v[k].x * scale : 0.f, v[k].y = (kb[k] & 2u) ? v[k].y * scale : 0.f;
          v[k].z = (kb[k] & 4u) ? v[k].z * scale : 0.f, v[k].w = (kb[k] & 8u) ? v[k].w * scale : 0.f;
        }
        if (res) {   // the normalised row was x + residual (ub_add_layernorm): the same sum, in the same order
          const float4 t = ld_stream4(res + r * C + c);
          v[k].x += t.x, v[k].y += t.y, v[k].z += t.z, v[k].w += t.w;
        }
        d[k] = ld_stream4(dy + r * C + c);
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
        v[k].x -= mean, v[k].y -= mean, v[k].z -= mean, v[k].w -= mean;
        sq += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      // v <- xhat; accumulate dgamma / dbeta; d <- g = gamma dy
      v[k].x *= rstd, v[k].y *= rstd, v[k].z *= rstd, v[k].w *= rstd;
      ag[k].x = fmaf(d[k].x, v[k].x, ag[k].x), ag[k].y = fmaf(d[k].y, v[k].y, ag[k].y);
      ag[k].z = fmaf(d[k].z, v[k].z, ag[k].z), ag[k].w = fmaf(d[k].w, v[k].w, ag[k].w);
      ab[k].x += d[k].x, ab[k].y += d[k].y, ab[k].z += d[k].z, ab[k].w += d[k].w;
      d[k].x *= gm[k].x, d[k].y *= gm[k].y, d[k].z *= gm[k].z, d[k].w *= gm[k].w;
      s1 += (d[k].x + d[k].y) + (d[k].z + d[k].w);
      s2 += (d[k].x * v[k].x + d[k].y * v[k].y) + (d[k].z * v[k].z + d[k].w * v[k].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * inv_c, m2 = s2 * inv_c;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < C) {
        float4 o;
        o.x = rstd * (d[k].x - m1 - v[k].x * m2);
        o.y = rstd * (d[k].y - m1 - v[k].y * m2);
        o.z = rstd * (d[k].z - m1 - v[k].z * m2);
        o.w = rstd * (d[k].w - m1 - v[k].w * m2);
        st_stream4(dx + r * C + c, o);
        if (dx_drop)   // gradient of the un-dropped x: the gradient of the sum through the keep mask
          st_stream4(dx_drop + r * C + c, make_float4((kb[k] & 1u) ? o.x * scale : 0.f, (kb[k] & 2u) ? o.y * scale : 0.f,
                                                       (kb[k] & 4u) ? o.z * scale : 0.f, (kb[k] & 8u) ? o.w * scale : 0.f));
      }
    }
  }
  // fold the block's eight warps, then one vector reduction per column quad and block
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      __syncthreads();
      s_red[warp][lane] = pass ? ab[k] : ag[k];
      __syncthreads();
      const int c = (k * 32 + lane) * 4;
      if (warp == 0 && c < C) {
        float4 t = s_red[0][lane];
        for (int w = 1; w < 8; ++w) {
          const float4 u = s_red[w][lane];
          t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
        }
        float* dst = (pass ? dbeta : dgamma) + c;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
      }
    }
  }
}

}  // namespace ub

using namespace ub;

extern "C" int ub_colsum(const float* x, float* out, int64_t M, int N, ub_stream_t stream) {
  UB_REQUIRE(x && out && M > 0 && N > 0, "ub_colsum: bad argument");
  UB_REQUIRE_ALIGNED16(x);
  UB_REQUIRE_ALIGNED16(out);
  if (N % 4 != 0 || N > 1024) {
    set_error("ub_colsum: N = %d not covered (need N %% 4 == 0, N <= 1024)", N);
    return ub::unsupported();
  }
  const int CG = N / 4, RL = 256 / CG;
  int64_t blocks = (M + 4 * RL - 1) / (4 * RL);      // >= 4 rows per thread
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  colsum_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, out, M, N);
  return check_launch("ub_colsum");
}

extern "C" int ub_layernorm_bwd(const float* x, const float* residual, const float* dy, const float* gamma, float* dx,
                                float* dgamma, float* dbeta, int64_t rows, int C, float eps, ub_stream_t stream) {
  UB_REQUIRE(x && dy && gamma && dx && dgamma && dbeta && rows > 0 && C > 0, "ub_layernorm_bwd: bad argument");
  UB_REQUIRE_ALIGNED16(x);
  if (residual) UB_REQUIRE_ALIGNED16(residual);
  UB_REQUIRE_ALIGNED16(dy);
  UB_REQUIRE_ALIGNED16(dx);
  UB_REQUIRE_ALIGNED16(gamma);
  UB_REQUIRE_ALIGNED16(dgamma);
  UB_REQUIRE_ALIGNED16(dbeta);
  if (C % 4 != 0 || C > 1024) {
    set_error("ub_layernorm_bwd: C = %d not covered (need C %% 4 == 0, C <= 1024)", C);
    return ub::unsupported();
  }
  int64_t blocks = (rows + 31) / 32;      // >= 4 rows per warp
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int nv = (C + 127) / 128;
  cudaStream_t s = (cudaStream_t)stream;
#define UB_LNB(NV) layernorm_bwd_kernel<NV><<<(int)blocks, 256, 0, s>>>(x, residual, dy, gamma, dx, dgamma, dbeta, rows, C, eps, nullptr, 1.f, nullptr)
  if (nv <= 1) UB_LNB(1);
  else if (nv <= 2) UB_LNB(2);
  else if (nv <= 4) UB_LNB(4);
  else UB_LNB(8);
#undef UB_LNB
  return check_launch("ub_layernorm_bwd");
}

static int check_dropout_ln(const char* fn, const void* x, const void* residual, int64_t rows, int C, float p) {
  UB_REQUIRE(x && residual && rows > 0 && C > 0, "%s: bad argument", fn);
  UB_REQUIRE(p >= 0.f && p < 1.f, "%s: drop probability %g outside [0, 1)", fn, p);
  UB_REQUIRE_ALIGNED16(x);
  UB_REQUIRE_ALIGNED16(residual);
  if (C % 4 != 0 || C > 1024) {
    set_error("%s: C = %d not covered (need C %% 4 == 0, C <= 1024)", fn, C);
    return ub::unsupported();
  }
  return UB_OK;
}

extern "C" int ub_dropout_add_layernorm_fwd(const float* x, const float* residual, const float* gamma, const float* beta,
                                            float* out, uint8_t* mask, int64_t rows, int C, float eps, float p,
                                            const int64_t* rng_state, int call_site, ub_stream_t stream) {
  const char* fn = "ub_dropout_add_layernorm_fwd";
  if (int rc = check_dropout_ln(fn, x, residual, rows, C, p)) return rc;
  UB_REQUIRE(gamma && beta && out && mask && rng_state, "%s: null pointer", fn);
  UB_REQUIRE_ALIGNED16(out);
  int64_t blocks = (rows + 7) / 8;
  if (blocks > (int64_t)sm_count() * 32) blocks = (int64_t)sm_count() * 32;
  const unsigned thr = (unsigned)lrintf(p * 65536.f);
  const float scale = 1.f / (1.f - (float)thr / 65536.f);      // of the probability actually applied
  const int nv = (C + 127) / 128;
  cudaStream_t s = (cudaStream_t)stream;
#define UB_DLN(NV) \
  dropout_add_layernorm_kernel<NV><<<(int)blocks, 256, 0, s>>>(x, residual, gamma, beta, out, mask, rows, C, eps, thr, scale, rng_state, call_site)
  if (nv <= 1) UB_DLN(1);
  else if (nv <= 2) UB_DLN(2);
  else if (nv <= 4) UB_DLN(4);
  else UB_DLN(8);
#undef UB_DLN
  return check_launch(fn);
}

extern "C" int ub_dropout_add_layernorm_bwd(const float* x, const float* residual, const uint8_t* mask, const float* dy,
                                            const float* gamma, float* dx, float* dresidual, float* dgamma, float* dbeta,
                                            int64_t rows, int C, float eps, float p, ub_stream_t stream) {
  const char* fn = "ub_dropout_add_layernorm_bwd";
  if (int rc = check_dropout_ln(fn, x, residual, rows, C, p)) return rc;
  UB_REQUIRE(mask && dy && gamma && dx && dresidual && dgamma && dbeta, "%s: null pointer", fn);
  UB_REQUIRE_ALIGNED16(dy);
  UB_REQUIRE_ALIGNED16(dx);
  UB_REQUIRE_ALIGNED16(dresidual);
  UB_REQUIRE_ALIGNED16(gamma);
  UB_REQUIRE_ALIGNED16(dgamma);
  UB_REQUIRE_ALIGNED16(dbeta);
  int64_t blocks = (rows + 31) / 32;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const unsigned thr = (unsigned)lrintf(p * 65536.f);
  const float scale = 1.f / (1.f - (float)thr / 65536.f);
  const int nv = (C + 127) / 128;
  cudaStream_t s = (cudaStream_t)stream;
#define UB_DLNB(NV) \
  layernorm_bwd_kernel<NV><<<(int)blocks, 256, 0, s>>>(x, residual, dy, gamma, dresidual, dgamma, dbeta, rows, C, eps, mask, scale, dx)
  if (nv <= 1) UB_DLNB(1);
  else if (nv <= 2) UB_DLNB(2);
  else if (nv <= 4) UB_DLNB(4);
  else UB_DLNB(8);
#undef UB_DLNB
  return check_launch(fn);
}
