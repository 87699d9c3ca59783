// Dense projections of the BEV encoder on the 5th-generation tensor cores (tcgen05, TF32 inputs, fp32
// accumulation in tensor memory) with the epilogues the encoder needs fused in:                         [R5]
//
//   out = epilogue(A (M, K) @ W (N, K)^T)       A, W row-major fp32 (both K-major), M = tokens / BEV queries
//
//   plain      out = [relu](acc + bias [+ residual])                        fp32 rows (ldc)
//   layernorm  out = LN(acc + bias + residual) * gamma + beta               full rows (N <= 256) in one tile
//   planes     out = fp16(acc + bias) written head-major (G, H, Nv, 32): the value-map layout of the
//              window-staged sampling kernels (win_sample.cu) -- no separate conversion pass
//
// Warp-specialised persistent kernel, one CTA per SM, tiles of 128 rows x BN (<= 256) columns:
//   warp 0   TMA producer: A / W k-blocks of 32 floats (128-byte swizzled rows) into a 3-stage ring
//   warp 1   one lane issues tcgen05.mma (kind::tf32, M = 128, N = BN, K = 8 per instruction) into one of two
//            TMEM accumulators, tcgen05.commit releases ring slots / publishes the accumulator
//   warp 2   allocates / frees tensor memory
//   warps 4-7 epilogue: thread = one row of the tile (its TMEM lane).  Rows are exchanged with global memory in
//            32-column chunks through per-warp 128-byte-swizzled shared-memory buffers moved by TMA (residual in,
//            result out), so global traffic is coalesced and the warps never synchronise with each other.  The
//            LayerNorm statistics are thread-private (one thread owns the whole row); the pre-norm row is parked
//            back in TMEM between the two passes.
// The GEMMs here are bound by their activation traffic, not by math: the point of the kernel is to touch each
// activation row once (no separate bias / residual / LayerNorm / fp16-conversion passes).
#include <cuda_fp16.h>

#include "ub_tma.cuh"

namespace ub {

constexpr int kGemmThreads = 256;
constexpr int kBM = 128, kBK = 32, kStages = 3;
constexpr int kEpiWarps = 4, kChunk = 32;          // epilogue column chunk
constexpr int kStageBuf = kChunk * 32 * 4;         // one warp's 32 rows x 32 columns staging buffer (4 KB)

struct GemmArgs {
  const float* bias;    // (N) or null
  const float* gamma;   // (N) layernorm
  const float* beta;
  __half* planes;       // fp16 head-major output or null
  int Nv, H;            // planes: rows per plane group, heads (= N / 32)
  int M, N, K, BN, n_tiles_m, n_tiles_n;
  float eps;
  int relu, ln, has_res;
};

// ---- tcgen05 wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in shared memory, rows of 128 bytes with the 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// D (fp32) += A (tf32, K-major) * B (tf32, K-major), M = 128
__device__ __forceinline__ uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31};" ::"r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
      "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
      "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 16-byte chunk j of row `row` in a 128-byte-swizzled buffer whose rows are 128 bytes
__device__ __forceinline__ uint32_t swz(uint32_t base, int row, int j) {
  return base + (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_tf32_kernel(const GemmArgs a, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages], s_tfull[2], s_tempty[2], s_res[kEpiWarps][2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_bytes = (uint32_t)(kBM * 128 + a.BN * 128);
  const uint32_t sm_base = smem_u32(smem);
  const uint32_t sm_stage_buf = sm_base + kStages * stage_bytes;                 // [4 warps][2] x 4 KB
  float* s_par = reinterpret_cast<float*>(smem + kStages * stage_bytes + kEpiWarps * 2 * kStageBuf);  // bias|gamma|beta
  const int n_tiles = a.n_tiles_m * a.n_tiles_n;
  const int k_blocks = a.K / kBK;

  for (int i = tid; i < a.N; i += kGemmThreads) {
    s_par[i] = a.bias ? a.bias[i] : 0.f;
    if (a.ln) s_par[a.N + i] = a.gamma[i], s_par[2 * a.N + i] = a.beta[i];
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(smem_u32(&s_full[i]), 1), mbar_init(smem_u32(&s_empty[i]), 1);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&s_tfull[i]), 1), mbar_init(smem_u32(&s_tempty[i]), kEpiWarps);
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(smem_u32(&s_res[i][0]), 1), mbar_init(smem_u32(&s_res[i][1]), 1);
    mbar_init_fence();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_w);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / a.n_tiles_n) * kBM, n0 = (t % a.n_tiles_n) * a.BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(smem_u32(&s_empty[stage]), phase ^ 1u);
          const uint32_t bar = smem_u32(&s_full[stage]);
          const uint32_t dst = sm_base + (uint32_t)stage * stage_bytes;
          mbar_arrive_expect_tx(bar, stage_bytes);
          tma_load_2d(dst, &map_a, bar, kb * kBK, m0);
          tma_load_2d(dst + kBM * 128, &map_w, bar, kb * kBK, n0);
          if (++stage == kStages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(a.BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(smem_u32(&s_tempty[acc]), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(smem_u32(&s_full[stage]), phase);
          tc_fence_after();
          const uint32_t sa = sm_base + (uint32_t)stage * stage_bytes;
          const uint64_t adesc = smem_desc_k128(sa), bdesc = smem_desc_k128(sa + kBM * 128);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k)   // 8 tf32 = 32 bytes per MMA along K: advance the start address
            mma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          mma_commit(smem_u32(&s_empty[stage]));
          if (++stage == kStages) stage = 0, phase ^= 1u;
        }
        mma_commit(smem_u32(&s_tfull[acc]));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const uint32_t buf0 = sm_stage_buf + (uint32_t)ew * 2 * kStageBuf;
    const uint32_t bar_res0 = smem_u32(&s_res[ew][0]);
    const int n_chunks = a.BN / kChunk;
    uint32_t res_uses = 0;   // completed waits on the residual barriers: parity of buffer b = (uses of b) & 1
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int m0 = (t / a.n_tiles_n) * kBM, n0 = (t % a.n_tiles_n) * a.BN;
      const int row0 = m0 + ew * 32;                        // this warp's 32 rows
      const uint32_t tbase = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(ew * 32) << 16);
      const bool rows_live = row0 < a.M;                    // warp-uniform: rows beyond M are never loaded / stored
      // residual chunks 0 / 1 stream in while the accumulator is still being computed
      if (a.has_res && lane == 0 && rows_live) {
        bulk_wait_read<0>();                                // the previous tile's stores have left the buffers
        for (int c = 0; c < 2 && c < n_chunks; ++c) {
          mbar_arrive_expect_tx(bar_res0 + 8u * c, kStageBuf);
          tma_load_2d(buf0 + (uint32_t)c * kStageBuf, &map_r, bar_res0 + 8u * c, n0 + c * kChunk, row0);
        }
      }
      mbar_wait(smem_u32(&s_tfull[acc]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();

      float sum = 0.f, sumsq = 0.f;
      // ---- pass A: acc + bias (+ residual); LayerNorm: statistics, row parked back in TMEM
      //              otherwise: activation and store
      for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        const uint32_t buf = buf0 + (uint32_t)b * kStageBuf;
        uint32_t v[32];
        tmem_ld32(tbase + (uint32_t)(c * kChunk), v);
        if (!rows_live) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_par[n0 + c * kChunk + j];
        if (a.has_res) {
          mbar_wait(bar_res0 + 8u * b, (res_uses >> b) & 1u);
          res_uses ^= 1u << b;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(swz(buf, lane, j)));
            f[4 * j] += r.x, f[4 * j + 1] += r.y, f[4 * j + 2] += r.z, f[4 * j + 3] += r.w;
          }
        }
        if (a.ln) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            sum += f[j];
            sumsq = fmaf(f[j], f[j], sumsq);
            v[j] = __float_as_uint(f[j]);
          }
          tmem_st32(tbase + (uint32_t)(c * kChunk), v);
          __syncwarp();
          if (a.has_res && lane == 0 && c + 2 < n_chunks) {   // the buffer is free: next residual chunk
            mbar_arrive_expect_tx(bar_res0 + 8u * b, kStageBuf);
            tma_load_2d(buf, &map_r, bar_res0 + 8u * b, n0 + (c + 2) * kChunk, row0);
          }
        } else if (a.planes) {
          // fp16 head-major planes: chunk c of the row is head (n0 / 32 + c) of token (row % Nv) in group row / Nv
          const int row = row0 + lane;
          if (row < a.M) {
            const int g = row / a.Nv, tok = row - g * a.Nv;
            uint4* dst = reinterpret_cast<uint4*>(a.planes + (((int64_t)g * a.H + (n0 / kChunk + c)) * a.Nv + tok) * 32);
            const float lim = 65504.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              uint32_t* ow = &o.x;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h = __floats2half2_rn(fminf(fmaxf(f[8 * j + 2 * e], -lim), lim),
                                                    fminf(fmaxf(f[8 * j + 2 * e + 1], -lim), lim));
                ow[e] = *reinterpret_cast<const uint32_t*>(&h);
              }
              dst[j] = o;
            }
          }
        } else {
          if (!a.has_res) {                                   // the buffer may still feed an earlier store
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            if (a.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(swz(buf, lane, j)), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_c, buf, n0 + c * kChunk, row0);
            bulk_commit();
            if (a.has_res && c + 2 < n_chunks) {              // reuse the buffer once the store has read it
              bulk_wait_read<0>();
              mbar_arrive_expect_tx(bar_res0 + 8u * b, kStageBuf);
              tma_load_2d(buf, &map_r, bar_res0 + 8u * b, n0 + (c + 2) * kChunk, row0);
            }
          }
        }
      }
      // ---- pass B (LayerNorm): normalise the parked row, store
      if (a.ln && rows_live) {
        const float inv_n = 1.f / (float)a.BN;
        const float mean = sum * inv_n;
        const float rstd = rsqrtf(fmaxf(sumsq * inv_n - mean * mean, 0.f) + a.eps);
        const float* gam = s_par + a.N + n0;
        const float* bet = s_par + 2 * a.N + n0;
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t buf = buf0 + (uint32_t)(c & 1) * kStageBuf;
          uint32_t v[32];
          tmem_ld32(tbase + (uint32_t)(c * kChunk), v);
          if (lane == 0) bulk_wait_read<1>();                 // the store issued two chunks ago has read this buffer
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = (__uint_as_float(v[4 * j]) - mean) * rstd * gam[c * kChunk + 4 * j] + bet[c * kChunk + 4 * j];
            o.y = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * gam[c * kChunk + 4 * j + 1] + bet[c * kChunk + 4 * j + 1];
            o.z = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * gam[c * kChunk + 4 * j + 2] + bet[c * kChunk + 4 * j + 2];
            o.w = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * gam[c * kChunk + 4 * j + 3] + bet[c * kChunk + 4 * j + 3];
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(swz(buf, lane, j)), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_c, buf, n0 + c * kChunk, row0);
            bulk_commit();
          }
        }
      }
      // the accumulator is drained
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_tempty[acc]));
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace ub

using namespace ub;

// out = epilogue(A (M, K) @ W (N, K)^T).  flags: bit 0 relu, bit 1 layernorm (needs residual-or-not, gamma, beta,
// N <= 256).  planes != NULL: fp16 head-major output (G = M / Nv groups, H = N / 32 heads), `out` ignored.
extern "C" int ub_linear_tf32(const float* A, const float* W, const float* bias, const float* residual, int ldr,
                              const float* gamma, const float* beta, float eps, float* out, int ldc, void* planes,
                              int Nv, int M, int N, int K, int flags, ub_stream_t stream) {
  const char* fn = "ub_linear_tf32";
  const int relu = flags & 1, ln = (flags >> 1) & 1;
  UB_REQUIRE(A && W && (out || planes), "%s: null pointer", fn);
  UB_REQUIRE(M > 0 && N > 0 && K > 0, "%s: non-positive dimension", fn);
  UB_REQUIRE(!ln || (gamma && beta), "%s: layernorm needs gamma and beta", fn);
  UB_REQUIRE_ALIGNED16(A);
  UB_REQUIRE_ALIGNED16(W);
  if (K % kBK != 0 || N % 32 != 0 || (N > 256 && N % 256 != 0) || (ln && N > 256) || (planes && (ln || relu || residual)) ||
      (planes && (Nv <= 0 || M % Nv != 0)) || (out && (ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(out) & 15u))) ||
      (residual && (ldr % 4 != 0 || (reinterpret_cast<uintptr_t>(residual) & 15u))) || N > 1024) {
    set_error("%s: shape not covered (M=%d N=%d K=%d flags=%d)", fn, M, N, K, flags);
    return UB_EUNSUPPORTED;
  }
  GemmArgs a;
  a.bias = bias, a.gamma = gamma, a.beta = beta, a.planes = reinterpret_cast<__half*>(planes);
  a.Nv = Nv, a.H = N / 32;
  a.M = M, a.N = N, a.K = K, a.BN = N > 256 ? 256 : N;
  a.n_tiles_m = (M + kBM - 1) / kBM, a.n_tiles_n = N / a.BN;
  a.eps = eps, a.relu = relu, a.ln = ln, a.has_res = residual != nullptr;
  CUtensorMap ma, mw, mc, mr;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)K * 4};
    const uint32_t box[2] = {kBK, kBM};
    if (int rc = make_tensor_map(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K * 4};
    const uint32_t box[2] = {kBK, (uint32_t)a.BN};
    if (int rc = make_tensor_map(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  const uint32_t cbox[2] = {kChunk, 32};
  if (out) {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ldc * 4};
    if (int rc = make_tensor_map(&mc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, str, cbox, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  } else {
    mc = ma;
  }
  if (residual) {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ldr * 4};
    if (int rc = make_tensor_map(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, residual, dims, str, cbox,
                                 CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  } else {
    mr = ma;
  }
  const size_t smem = (size_t)kStages * (kBM * 128 + a.BN * 128) + kEpiWarps * 2 * kStageBuf + (size_t)3 * N * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("%s: cannot reserve %zu bytes of shared memory", fn, smem);
      cudaGetLastError();
      return UB_ECUDA;
    }
    configured = smem;
  }
  const int n_tiles = a.n_tiles_m * a.n_tiles_n;
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  gemm_tf32_kernel<<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(a, ma, mw, mc, mr);
  return check_launch(fn);
}
