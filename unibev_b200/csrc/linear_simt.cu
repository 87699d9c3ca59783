// Generic fp32 projection for the shapes the tensor-core GEMM (gemm_tc.cu) does not cover (K or N not a multiple of 32,
// e.g. the 4-head / 2-point toy configurations of the golden fixtures): plain FFMA arithmetic, exactly the class of the
// reference's fp32 nn.Linear.  64 x 64 output tile per CTA, 16-wide k-steps through shared memory, 4 x 4 outputs per
// thread.  Not a hot kernel: every shipped UniBEV configuration runs on ub_linear_tf32x3 / ub_linear_f16.
#include "ub_common.cuh"

namespace ub {

constexpr int kSM = 64, kSN = 64, kSK = 16;

__global__ void __launch_bounds__(256) linear_simt_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                          const float* __restrict__ bias, const float* __restrict__ res,
                                                          int ldr, float* __restrict__ out, int ldc, int M, int N, int K,
                                                          int relu) {
  __shared__ float sA[kSK][kSM + 4], sW[kSK][kSN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kSM, n0 = blockIdx.x * kSN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kSK) {
    for (int e = tid; e < kSM * kSK; e += 256) {
      const int r = e / kSK, k = e % kSK;
      sA[k][r] = (m0 + r < M && k0 + k < K) ? A[(int64_t)(m0 + r) * K + k0 + k] : 0.f;
      sW[k][r] = (n0 + r < N && k0 + k < K) ? W[(int64_t)(n0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i], w[i] = sW[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (res) v += res[(int64_t)m * ldr + n];
      if (relu) v = fmaxf(v, 0.f);
      out[(int64_t)m * ldc + n] = v;
    }
  }
}

}  // namespace ub

// out (M, N; row stride ldc) = [relu](A (M, K) @ W (N, K)^T + bias + residual (row stride ldr)); any M, N, K > 0.
extern "C" int ub_linear_simt(const float* A, const float* W, const float* bias, const float* residual, int ldr, float* out,
                              int ldc, int M, int N, int K, int relu, ub_stream_t stream) {
  UB_REQUIRE(A && W && out, "ub_linear_simt: null pointer");
  UB_REQUIRE(M > 0 && N > 0 && K > 0 && ldc >= N && (!residual || ldr >= N), "ub_linear_simt: bad dimension");
  const dim3 grid((N + ub::kSN - 1) / ub::kSN, (M + ub::kSM - 1) / ub::kSM);
  UB_REQUIRE(grid.y <= 65535, "ub_linear_simt: M too large for this generic kernel");
  ub::linear_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, W, bias, residual, ldr, out, ldc, M, N, K, relu);
  return ub::check_launch("ub_linear_simt");
}
