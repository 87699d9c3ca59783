// 3xTF32 projection on CTA PAIRS: tcgen05.mma.cta_group::2 (M = 256 = 2 x 128 rows, one 128-row tile per CTA of the pair).
//
// The single-CTA 3xTF32 kernel (gemm_tc.cu, SPLIT) is bound by the shared-memory data pipe, not by the tensor pipe or HBM
// (profiles/r2_gemm_x3_ncu.txt: tensor-core operand wavefronts 39 % + LSU wavefronts 49 % of the pipe's peak, tensor pipe
// 23 % active): per 32-float k-block every CTA writes a whole W_hi / W_lo k-block pair into shared memory by TMA (64 KB at
// N = 256) and its three MMA groups read 96 KB of B operand from it.  A CTA pair halves both: each CTA holds HALF of the
// W rows (N / 2), the pair's MMA reads each half once for both 128-row tiles.
//
//   both CTAs   warp 0  A producer: residual boxes + A k-blocks of ITS 128 rows -> its own ring (local full barrier)
//               warp 3  W producer: its half of the W_hi / W_lo k-blocks -> its own ring; the TMA (.cta_group::2) completes on
//                       the LEADER's full barrier, which expects both halves
//               warps 2, 20, 21  converters: A slot -> (a_hi in place, a_lo in the LO slot of the same index); arrive on the
//                       LEADER's ready barrier (remote arrive, cluster scope) and on the local slot-passed barrier
//               warps 4-19  epilogue of the CTA's own 128 x N accumulator (its own TMEM lanes); release the accumulator on
//                       the LEADER's barrier
//   leader only warp 1  issues tcgen05.cp (residual -> accumulator, both CTAs) and tcgen05.mma for the pair; every
//                       tcgen05.commit is multicast to the same barrier in both CTAs
// Results leave straight from registers (64 contiguous bytes per thread and chunk): no staging buffers, which is where the
// third W ring slot pair comes from.
#include "gemm_common.cuh"

namespace ub {

constexpr int kPairThreads = 640 + 32 * kConvExtraWarps;   // 4 control warps + 16 epilogue warps + 2 extra converter warps
constexpr int kPairSA = 3;                                 // A / LO ring depth (slots of 128 rows x 32 floats)
constexpr int kPairMaxSW = 8;

struct PairArgs {
  const float* bias;
  const float* gamma;
  const float* beta;
  float* out;           // (M, N) row stride ldc, or null with planes32
  float* planes32;      // fp32 half-head planes (G, N / 16, Nv, 16), or null
  int ldc, Nv;
  int M, N, K, BN, n_tiles_m, n_tiles_n;
  float eps;
  int relu, ln;
  int SW;               // W ring slots (each: BN / 2 rows x 32 floats), hi / lo k-blocks alternate
  int res_chunks;       // residual boxes (32 fp32 columns x 128 rows) per tile preloaded into the accumulator, 0 = none
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the arrivals come from the peer CTA as well
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA tile load into THIS CTA's shared memory whose completion is posted on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b_pair(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs / copies have completed
__device__ __forceinline__ void commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// D (fp32) += A (tf32, K-major) * B (tf32, K-major), M = 256 over the pair
__device__ __forceinline__ uint32_t idesc_tf32_pair(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__global__ void __launch_bounds__(kPairThreads, 1)
    gemm_x3_pair_kernel(const PairArgs a, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                        const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_wl) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_fa[kPairSA], s_ea[kPairSA], s_rdy[kPairSA], s_fw[kPairMaxSW], s_ew[kPairMaxSW], s_tfull[2],
      s_tempty[2];
  __shared__ uint32_t s_tmem;
  __shared__ float2 s_stat[4][kBM];   // LayerNorm partial (sum, sum of squares) per epilogue warp of a lane quarter

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int SA = kPairSA;
  const int SW = a.SW;
  const uint32_t a_bytes = kBM * 128, w_bytes = (uint32_t)(a.BN / 2) * 128;
  const uint32_t sm_a = smem_u32(smem), sm_l = sm_a + (uint32_t)SA * a_bytes, sm_w = sm_l + (uint32_t)SA * a_bytes;
  float* s_par = reinterpret_cast<float*>(smem + (size_t)2 * SA * a_bytes + (size_t)SW * w_bytes);
  const int k_blocks = a.K / 32;
  // work = pairs of consecutive row tiles of one column tile; CTA pair p takes items p, p + n_pairs, ...
  const int pair_id = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  const int groups_m = (a.n_tiles_m + 1) >> 1, n_groups = groups_m * a.n_tiles_n;
  const int n_iter = pair_id < n_groups ? (n_groups - pair_id + n_pairs - 1) / n_pairs : 0;
  auto tile_m0 = [&](int i) { return (((pair_id + i * n_pairs) % groups_m) * 2 + (int)rank) * kBM; };
  auto tile_n0 = [&](int i) { return ((pair_id + i * n_pairs) / groups_m) * a.BN; };

  for (int i = tid; i < a.N; i += kPairThreads) {
    s_par[i] = a.bias ? a.bias[i] : 0.f;
    if (a.ln) s_par[a.N + i] = a.gamma[i], s_par[2 * a.N + i] = a.beta[i];
  }
  if (tid == 0) {
    for (int i = 0; i < SA; ++i) {
      mbar_init(smem_u32(&s_fa[i]), 1);
      mbar_init(smem_u32(&s_ea[i]), 1 + kConvWarps);      // the pair's MMA commit + this CTA's converter warps
      mbar_init(smem_u32(&s_rdy[i]), 2 * kConvWarps);     // (leader) the converter warps of both CTAs
    }
    for (int i = 0; i < SW; ++i) mbar_init(smem_u32(&s_fw[i]), 1), mbar_init(smem_u32(&s_ew[i]), 1);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&s_tfull[i]), 1), mbar_init(smem_u32(&s_tempty[i]), 2 * kEpiWarps);
    mbar_init_fence();
  }
  if (warp == 2) {   // one warp of EACH CTA of the pair: the allocation is collective
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers and TMEM exist before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer (this CTA's rows)
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      if (a.res_chunks) tma_prefetch_desc(&map_r);
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_iter; ++i) {
        const int m0 = tile_m0(i), n0 = tile_n0(i);
        for (int rc = 0; rc < a.res_chunks; ++rc) {
          mbar_wait(smem_u32(&s_ea[stage]), phase ^ 1u);
          const uint32_t bar = smem_u32(&s_fa[stage]);
          mbar_arrive_expect_tx(bar, a_bytes);
          tma_load_2d(sm_a + (uint32_t)stage * a_bytes, &map_r, bar, n0 + rc * 32, m0);
          if (++stage == SA) stage = 0, phase ^= 1u;
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(smem_u32(&s_ea[stage]), phase ^ 1u);
          const uint32_t bar = smem_u32(&s_fa[stage]);
          mbar_arrive_expect_tx(bar, a_bytes);
          tma_load_2d(sm_a + (uint32_t)stage * a_bytes, &map_a, bar, kb * 32, m0);
          if (++stage == SA) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ W producer (this CTA's half of the W rows)
    if (lane == 0) {
      tma_prefetch_desc(&map_w);
      tma_prefetch_desc(&map_wl);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_iter; ++i) {
        const int n0 = tile_n0(i) + (int)rank * (a.BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
#pragma unroll
          for (int part = 0; part < 2; ++part) {   // the hi k-block, then the lo k-block
            mbar_wait(smem_u32(&s_ew[stage]), phase ^ 1u);   // the pair's MMAs have consumed the slot (commit is multicast)
            const uint32_t lbar = mapa(smem_u32(&s_fw[stage]), 0);
            if (leader) mbar_arrive_expect_tx(smem_u32(&s_fw[stage]), 2u * w_bytes);   // both halves
            tma_load_2d_pair(sm_w + (uint32_t)stage * w_bytes, part ? &map_wl : &map_w, lbar, kb * 32, n0);
            if (++stage == SW) stage = 0, phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one lane)
    if (leader && lane == 0) {
      const uint32_t idesc = idesc_tf32_pair(a.BN);
      int sa = 0, sw = 0;
      uint32_t pa = 0, pw = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int acc = it & 1;
        mbar_wait_cluster(smem_u32(&s_tempty[acc]), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        for (int rc = 0; rc < a.res_chunks; ++rc) {   // accumulators <- residual tiles (each CTA's own box)
          mbar_wait_cluster(smem_u32(&s_rdy[sa]), pa);
          tc_fence_after();
          const uint64_t rdesc = smem_desc_k128(sm_a + (uint32_t)sa * a_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) tmem_cp_128x256b_pair(tmem_d + (uint32_t)(rc * 32 + k * 8), rdesc + (uint64_t)(k * 2));
          commit_pair(smem_u32(&s_ea[sa]));
          if (++sa == SA) sa = 0, pa ^= 1u;
        }
        const uint32_t acc0 = a.res_chunks ? 1u : 0u;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait_cluster(smem_u32(&s_rdy[sa]), pa);   // both CTAs: A slot masked to a_hi, LO slot filled
          mbar_wait_cluster(smem_u32(&s_fw[sw]), pw);    // W_hi k-block, both halves
          tc_fence_after();
          const uint64_t adesc = smem_desc_k128(sm_a + (uint32_t)sa * a_bytes);
          const uint64_t ldesc = smem_desc_k128(sm_l + (uint32_t)sa * a_bytes);
          uint64_t bdesc = smem_desc_k128(sm_w + (uint32_t)sw * w_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_tf32_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : acc0);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_tf32_pair(tmem_d, ldesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
          commit_pair(smem_u32(&s_ew[sw]));
          if (++sw == SW) sw = 0, pw ^= 1u;
          mbar_wait_cluster(smem_u32(&s_fw[sw]), pw);    // W_lo k-block
          tc_fence_after();
          bdesc = smem_desc_k128(sm_w + (uint32_t)sw * w_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_tf32_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
          commit_pair(smem_u32(&s_ew[sw]));
          if (++sw == SW) sw = 0, pw ^= 1u;
          commit_pair(smem_u32(&s_ea[sa]));
          if (++sa == SA) sa = 0, pa ^= 1u;
        }
        commit_pair(smem_u32(&s_tfull[acc]));
      }
    }
  } else if (warp == 2 || warp >= 4 + kEpiWarps) {
    // ------------------------------------------------------------------ converters: A k-block -> (a_hi in place, a_lo)
    const int ct = (warp == 2 ? 0 : warp - (4 + kEpiWarps) + 1) * 32 + lane;   // 0 .. 32 kConvWarps - 1
    int sa = 0;
    uint32_t pa = 0;
    for (int it = 0; it < n_iter; ++it) {
      for (int rc = 0; rc < a.res_chunks; ++rc) {   // residual boxes pass untouched: report them ready, mark the slot passed
        mbar_wait(smem_u32(&s_fa[sa]), pa);
        if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&s_rdy[sa]), 0)), mbar_arrive(smem_u32(&s_ea[sa]));
        if (++sa == SA) sa = 0, pa ^= 1u;
      }
      for (int kb = 0; kb < k_blocks; ++kb) {
        // (the LO slot of this index is free: its last readers completed before the A slot was refilled)
        mbar_wait(smem_u32(&s_fa[sa]), pa);
        // all of a thread's loads first (plain C++ accesses: the compiler keeps them in flight together), then the splits
        constexpr int kPer = (1024 + 32 * kConvWarps - 1) / (32 * kConvWarps);
        unsigned char* a_slot = smem + (size_t)sa * a_bytes;
        unsigned char* l_slot = smem + (size_t)(SA + sa) * a_bytes;
        uint4 xs[kPer];
#pragma unroll
        for (int m = 0; m < kPer; ++m) {
          const uint32_t j = ((uint32_t)ct + 32u * kConvWarps * m) * 16u;
          xs[m] = j < a_bytes ? *reinterpret_cast<const uint4*>(a_slot + j) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int m = 0; m < kPer; ++m) {
          const uint32_t j = ((uint32_t)ct + 32u * kConvWarps * m) * 16u;
          const uint32_t x[4] = {xs[m].x, xs[m].y, xs[m].z, xs[m].w};
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h[e] = x[e] & 0xffffe000u;
            const float lo = __uint_as_float(x[e]) - __uint_as_float(h[e]);    // exact
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l[e]) : "f"(lo));
          }
          if (j < a_bytes) {
            *reinterpret_cast<uint4*>(a_slot + j) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(l_slot + j) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&s_rdy[sa]), 0)), mbar_arrive(smem_u32(&s_ea[sa]));
        if (++sa == SA) sa = 0, pa ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ------------------------------------------------------------------ epilogue: thread = one row of this CTA's tile
    const int ew = warp - 4, q = ew & 3, part = ew >> 2;
    const int n_chunks = a.BN / kChunk;
    const uint32_t leader_tempty = mapa(smem_u32(&s_tempty[0]), 0);
    for (int it = 0; it < n_iter; ++it) {
      const int acc = it & 1;
      const int m0 = tile_m0(it), n0 = tile_n0(it);
      const int row = m0 + q * 32 + lane;
      const uint32_t tbase = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
      const bool live = row < a.M;

      auto store_chunk = [&](int c, const float (&f)[16]) {   // 64 contiguous bytes of this thread's row
        if (!live) return;
        float* p;
        if (a.planes32) {
          const int gi = row / a.Nv, tok = row - gi * a.Nv;
          p = a.planes32 + (((int64_t)gi * (a.N / 16) + (n0 / 16 + c)) * a.Nv + tok) * 16;
        } else {
          p = a.out + (size_t)row * a.ldc + n0 + c * kChunk;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + 8 * j), "f"(f[8 * j]),
                       "f"(f[8 * j + 1]), "f"(f[8 * j + 2]), "f"(f[8 * j + 3]), "f"(f[8 * j + 4]), "f"(f[8 * j + 5]),
                       "f"(f[8 * j + 6]), "f"(f[8 * j + 7])
                       : "memory");
      };
      auto params4 = [&](const float* p, int j) { return *reinterpret_cast<const float4*>(p + 4 * j); };

      mbar_wait_cluster(smem_u32(&s_tfull[acc]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float sum = 0.f, sumsq = 0.f;
      for (int c = part; c < n_chunks; c += 4) {
        uint32_t v[16];
        tmem_ld16(tbase + (uint32_t)(c * kChunk), v);
        float f[16];
        const float* bias = s_par + n0 + c * kChunk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = params4(bias, j);
          f[4 * j] = __uint_as_float(v[4 * j]) + b4.x, f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z, f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
        }
        if (a.ln) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum += f[j];
            sumsq = fmaf(f[j], f[j], sumsq);
          }
        } else {
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          store_chunk(c, f);
        }
      }
      if (a.ln) {   // the four warps of the lane quarter exchange their partial row statistics, then normalise their chunks
        s_stat[part][q * 32 + lane] = make_float2(sum, sumsq);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
        float tsum = 0.f, tsq = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 o = s_stat[p][q * 32 + lane];
          tsum += o.x, tsq += o.y;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // s_stat is reused by the next tile
        const float inv_n = 1.f / (float)a.BN;
        const float mean = tsum * inv_n;
        const float rstd = rsqrtf(fmaxf(tsq * inv_n - mean * mean, 0.f) + a.eps);
        for (int c = part; c < n_chunks; c += 4) {
          uint32_t v[16];
          tmem_ld16(tbase + (uint32_t)(c * kChunk), v);
          float f[16];
          const float* bias = s_par + n0 + c * kChunk;
          const float* gam = s_par + a.N + n0 + c * kChunk;
          const float* bet = s_par + 2 * a.N + n0 + c * kChunk;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 g4 = params4(gam, j), b4 = params4(bet, j), c4 = params4(bias, j);
            f[4 * j] = (__uint_as_float(v[4 * j]) + c4.x - mean) * rstd * g4.x + b4.x;
            f[4 * j + 1] = (__uint_as_float(v[4 * j + 1]) + c4.y - mean) * rstd * g4.y + b4.y;
            f[4 * j + 2] = (__uint_as_float(v[4 * j + 2]) + c4.z - mean) * rstd * g4.z + b4.z;
            f[4 * j + 3] = (__uint_as_float(v[4 * j + 3]) + c4.w - mean) * rstd * g4.w + b4.w;
          }
          store_chunk(c, f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty + 8u * (uint32_t)acc);   // the accumulator pair is drained (this warp)
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves (or frees tensor memory) while its peer may still signal it or read its shared memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// Host side: called by gemm_tc.cu's launcher for the shapes this kernel covers.  Returns UB_OK, an error, or
// UB_EUNSUPPORTED (the caller then uses the single-CTA kernel) -- without counting it as a fallback.
int launch_x3_pair(const char* fn, const float* A, const float* W_hi, const float* W_lo, const float* bias, const float* residual,
                   int ldr, const float* gamma, const float* beta, float eps, float* out, int ldc, float* planes32, int Nv,
                   int M, int N, int K, int relu, int ln, cudaStream_t stream) {
  PairArgs a;
  a.bias = bias, a.gamma = gamma, a.beta = beta, a.out = out, a.planes32 = planes32, a.ldc = ldc, a.Nv = Nv;
  a.M = M, a.N = N, a.K = K, a.BN = N > 256 ? 256 : N;
  a.n_tiles_m = (M + kBM - 1) / kBM, a.n_tiles_n = N / a.BN;
  a.eps = eps, a.relu = relu, a.ln = ln;
  if (K % 32 != 0 || a.BN % 32 != 0 || N % a.BN != 0 || a.n_tiles_m < 2 || (out && (ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 31u))))
    return UB_EUNSUPPORTED;
  const size_t fixed = (size_t)3 * N * sizeof(float);
  const size_t budget = 232448 - 6144 - 1024;   // minus static shared memory and slack
  const size_t rings = (size_t)2 * kPairSA * kBM * 128;
  const size_t w_slot = (size_t)(a.BN / 2) * 128;
  a.SW = (int)((budget - fixed - rings) / w_slot);
  if (a.SW > kPairMaxSW) a.SW = kPairMaxSW;
  if (a.SW < 4) return UB_EUNSUPPORTED;
  CUtensorMap ma, mw, mwl, mr;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)K * 4};
    const uint32_t box[2] = {32, kBM};
    if (int rc = make_tensor_map(&ma, dt, 2, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K * 4};
    const uint32_t box[2] = {32, (uint32_t)(a.BN / 2)};
    if (int rc = make_tensor_map(&mw, dt, 2, W_hi, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_tensor_map(&mwl, dt, 2, W_lo, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  a.res_chunks = 0;
  mr = ma;
  if (residual) {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ldr * 4};
    const uint32_t box[2] = {32, kBM};
    if (int rc = make_tensor_map(&mr, dt, 2, residual, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    a.res_chunks = a.BN / 32;
  }
  const size_t smem = rings + (size_t)a.SW * w_slot + fixed;
  if (int rc = ensure_smem(gemm_x3_pair_kernel, smem, fn)) return rc;
  const int n_groups = ((a.n_tiles_m + 1) / 2) * a.n_tiles_n;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 2 : 1;
  int pairs = sm_count() / 2;
  if (pairs > n_groups) pairs = n_groups;
  cfg.gridDim = dim3(2 * pairs);
  if (cudaLaunchKernelEx(&cfg, gemm_x3_pair_kernel, a, ma, mw, mr, mwl) != cudaSuccess) {
    set_error("%s: pair-kernel launch failed: %s", fn, cudaGetErrorString(cudaGetLastError()));
    return UB_ECUDA;
  }
  return check_launch(fn);
}

}  // namespace ub
