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
//   warps 4-19 epilogue: thread = one row of the tile (its TMEM lane); the four warps of a lane quarter share the
//            16-column chunks of their rows.  Result rows reach global memory through per-warp swizzled
//            shared-memory buffers (coalesced 16-byte stores).
//   residual  never touches the epilogue: the producer streams the residual tile through the A ring (fp32 boxes of
//            32 columns x 128 rows, the same 128-byte-swizzled layout as a TF32 A k-block) and the MMA lane copies
//            each box straight into the accumulator with tcgen05.cp (smem -> TMEM) BEFORE the tile's MMAs, which
//            then accumulate on top of it.  The residual so rides the deep asynchronous TMA pipeline instead of
//            exposing its DRAM latency to the epilogue threads, and LayerNorm needs no second TMEM write pass.
//   fp16 copies of the result rows / value planes leave each epilogue thread as one 32-byte store (STG.256): the
//            thread-per-row epilogue is bound by LSU line lookups, not by bytes.
//   launch   programmatic dependent launch: set-up (barriers, TMEM, parameters, resident W) overlaps the predecessor's
//            tail; the A producer waits for the predecessor (griddepcontrol.wait) before its first load.
// The GEMMs here are bound by their activation traffic, not by math: the point of the kernel is to touch each
// activation row once (no separate bias / residual / LayerNorm / fp16-conversion passes).
#include <cuda_fp16.h>

#include <atomic>

#include "gemm_common.cuh"

namespace ub {

constexpr int kGemmThreads = 640;                 // 4 control warps + 16 epilogue warps
constexpr int kGemmThreadsSplit = kGemmThreads + 32 * kConvExtraWarps;
constexpr int kConvWarpsF16 = 6;                  // fp16 x3 mode: its converters do more work per chunk (scale, two roundings, swizzle)
constexpr int kGemmThreadsF16S = kGemmThreads + 32 * (kConvWarpsF16 - 1);   // 800 threads x 80 registers still fit the SM
constexpr int kMaxStages = 8;
constexpr int kStageBuf = kChunk * 32 * 4;         // one warp's 32 rows x 16 columns staging buffer (2 KB)

struct GemmArgs {
  const float* bias;    // (N) or null
  const float* gamma;   // (N) layernorm
  const float* beta;
  const float* A;         // (M, K) contiguous (L2 prefetch; the k-blocks themselves come through the tensor map)
  const float* residual;  // (M, N) row stride ldr, or null
  float* out;           // (M, N) row stride ldc (unused with planes)
  int ldr, ldc;
  __half* planes;       // fp16 head-major output or null
  int Nv, H;            // planes: rows per plane group, heads (= N / 32)
  int M, N, K, BN, n_tiles_m, n_tiles_n;
  float eps;
  int relu, ln;
  int SA, SW;           // ring depths (A k-blocks, W k-blocks)
  int f16;              // operands are fp16 (kind::f16, 64-element k-blocks) instead of fp32 / TF32 (32-element k-blocks)
  int kb_elems;         // elements per k-block: one 128-byte swizzled row
  int balanced;         // contiguous, equally sized row range per CTA (see the kernel)
  int n_split;          // balanced mode: CTAs per row range, one column tile each (1 or n_tiles_n)
  int w_res;            // the whole W tile stays resident in the W ring (loaded once per CTA)
  __half* out16;        // optional fp16 copy of the result rows (row stride ldc16), the next GEMM's A operand
  int ldc16;
  int cs;               // CTAs per cluster: they work on `cs` consecutive row tiles and share W by TMA multicast
  int res_chunks;       // residual boxes (32 fp32 columns x 128 rows) per tile preloaded into the accumulator, 0 = none
  unsigned long long* trace;   // optional per-CTA event timestamps (tools/trace_gemm.py), else null
  int SL;               // 3xTF32 mode: depth of the A_lo ring
  float* planes32;      // fp32 half-head planes (G, N / 16, Nv, 16): the value-map layout of the fp32 window kernels, or null
  // row scatter (camera cross-attention: offset|logit rows written in hit-list order): row r = b * sc_rows + q goes to the rows
  // b * sc_dst_rows + scatter[q * sc_r + j], j = 0 .. until the first negative entry, of `out`; null: row r goes to row r
  const int* scatter;
  int sc_r, sc_rows, sc_dst_rows;
  // MODE 2 (fp16 x3): A is multiplied by a_scale (a power of two chosen from the caller's bound) before the split, W row n
  // arrives multiplied by a power of two s_n; cscale[n] = 1 / (a_scale s_n) undoes both in the epilogue (exact).  The residual
  // is added in the epilogue (it cannot ride in the scaled accumulator).
  int no_staging;       // no per-warp staging buffers in shared memory (results leave straight from registers)
  float a_scale;
  const float* cscale;  // (N) or null (= 1 / a_scale)
  // bound_ptr != null: the bound of |A| lives on the device (|A| <= *bound_ptr * bound_mul + bound_add, e.g. the largest input
  // token found by ub_flatten_feats_max, pushed through a projection's row sums); the kernel derives a_scale from it and
  // cscale[n] is 1 / s_n only
  const float* bound_ptr;
  float bound_mul, bound_add;
  int epi_res;
  int x3_inplace;       // 3xTF32: the converters also write a_hi back (result independent of the tensor core's operand rounding)
  int stagger_ns;       // every other cluster starts its first tile this much later: the CTAs' epilogue bursts (output stores)
                        // then interleave with the others' operand loads instead of all hitting DRAM at once
  int direct_store;     // fp32 result rows straight from registers (2 x 32 B per thread and chunk) instead of through the
                        // per-warp staging buffers: fewer shared-memory wavefronts where the data pipe is the bottleneck
};

// trace slots per CTA: [0] start, [1 + 32 r + i]: role r (0 producer, 1 mma, 2 epilogue warp 0), event i
constexpr int kTraceSlots = 128;
__device__ __forceinline__ void trace_event(const GemmArgs& a, int role, int i) {
  if (a.trace && i < 40) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(size_t)blockIdx.x * kTraceSlots + 1 + role * 40 + i] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory rings: the A k-blocks come from DRAM (about 1.5 us away under load), so their ring is deep; the W
// k-blocks are re-read from L2 by every tile and need only a shallow ring.  Each ring has its own producer lane.

//
// SPLIT (3xTF32, ub_linear_tf32x3): fp32-grade products from three TF32 MMAs per k-step.  kind::tf32 reads the upper 19
// bits of each fp32 operand, so  a * w  ~=  a_hi * w_hi + a_lo * w_hi + a_hi * w_lo  with x_hi = x & ~0x1fff and
// x_lo = x - x_hi (exact; rounded to TF32): the dropped term a_lo * w_lo is 2^-22 relative.  W arrives pre-split from the
// host (two tensors, W ring slots alternate hi / lo k-blocks); the A k-blocks are split on chip by three converter warps
// (warp 2 and two extra warps): they mask the TMA-written A slot in place (a_hi, so the result does not depend on how
// the tensor core rounds its operands) and write a_lo into a second, identically laid out (swizzle and all: the
// transform is element-wise) ring, which the MMA lane consumes through the same kind of descriptors.  All three products
// accumulate into the same TMEM columns, so residual-via-tcgen05.cp and the epilogues are unchanged.
//
// MODE 2 (fp16 x3, ub_linear_f16x3): the same three-product scheme on kind::f16 MMAs, which run at twice the TF32 rate and
// move half the operand bytes: a = a_hi + a_lo with a_hi = fp16(a), a_lo = fp16(a - a_hi) (2 x 11 significand bits, the same
// as the TF32 split) -- valid where the caller can BOUND the operands below the fp16 range (plugin/fused.py derives the
// bounds from the weights; everything else stays on MODE 1).  W_hi / W_lo arrive as fp16 tensors (64-element k-blocks); the
// converters turn every two fp32 A k-blocks into one 64-element fp16 operand slot {a_hi tile, a_lo tile}, writing the
// 128-byte swizzle themselves (the element size changes, so the layout does); the MMA lane never reads the A ring.
template <int MODE>
__global__ void __launch_bounds__(MODE == 2 ? kGemmThreadsF16S : (MODE ? kGemmThreadsSplit : kGemmThreads), 1)
    gemm_tf32_kernel(const GemmArgs a, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_wl) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_fa[kMaxStages], s_ea[kMaxStages], s_fw[kMaxStages], s_ew[kMaxStages], s_tfull[2],
      s_tempty[2], s_fl[4], s_el[4];
  __shared__ uint32_t s_tmem;
  __shared__ float2 s_stat[4][kBM];   // LayerNorm partial (sum, sum of squares) per epilogue warp of a lane quarter

  constexpr bool SPLIT = MODE == 1, F16S = MODE == 2;
  constexpr int CW = F16S ? kConvWarpsF16 : kConvWarps;   // converter warps: warp 2 and the warps after the epilogue warps
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SA = a.SA, SW = a.SW;
  const uint32_t a_bytes = kBM * 128, w_bytes = (uint32_t)a.BN * 128;
  const int SL = MODE ? a.SL : 0;
  const uint32_t l_bytes = F16S ? 2u * a_bytes : a_bytes;      // MODE 2: an operand slot = {a_hi tile, a_lo tile} of 64 k
  const uint32_t sm_a = smem_u32(smem), sm_l = sm_a + (uint32_t)SA * a_bytes, sm_w = sm_l + (uint32_t)SL * l_bytes;
  const uint32_t sm_stage_buf = sm_w + (uint32_t)SW * w_bytes;                  // [16 warps] x 2 KB (none in MODE 2)
  float* s_par = reinterpret_cast<float*>(smem + (size_t)SA * a_bytes + (size_t)SL * l_bytes + (size_t)SW * w_bytes +
                                          (a.no_staging ? 0 : kEpiWarps * kStageBuf));
  const int k_blocks = a.K / a.kb_elems;                       // A k-blocks (MODE 2: 32 fp32 each, two per W k-block)
  const int w_blocks = F16S ? a.K / 64 : k_blocks, w_elems = F16S ? 64 : a.kb_elems;
  // Work = groups of `cs` consecutive row tiles of one column tile; cluster c takes groups c, c + n_clusters, ...
  // and CTA rank r of the cluster the r-th row tile of the group (possibly past M: loads zero-fill, nothing is stored).
  const int cs = a.cs, rank = (int)blockIdx.x % cs, cluster_id = (int)blockIdx.x / cs, n_clusters = (int)gridDim.x / cs;
  const int groups_m = (a.n_tiles_m + cs - 1) / cs, n_groups = groups_m * a.n_tiles_n;
  const uint16_t cta_mask = (uint16_t)((1u << cs) - 1u);
  const uint32_t w_slice_rows = (uint32_t)(a.BN / cs), w_slice_bytes = w_slice_rows * 128u;
  // Balanced mode (one column tile, no clusters): every CTA owns a contiguous range of M / grid rows and walks it
  // in 128-row tiles; the last tile of a range is partial (its extra rows belong to the next CTA and are computed
  // but not stored), so all CTAs carry the same load instead of 2 or 3 whole tiles.
  int r_beg = 0, r_end = a.M, n_iter;
  // With several column tiles (n_split > 1) the CTAs blockIdx = g * n_split + j share row range g and take column
  // tile j each: neighbours in time and space, so the second read of an A tile is an L2 hit, and every CTA keeps
  // its own W tile resident.
  const int my_n0 = a.balanced ? ((int)blockIdx.x % a.n_split) * a.BN : 0;
  if (a.balanced) {
    const int n_ranges = (int)gridDim.x / a.n_split, rg = (int)blockIdx.x / a.n_split;
    const int base = a.M / n_ranges, rem = a.M % n_ranges;
    r_beg = rg * base + min(rg, rem);
    r_end = r_beg + base + (rg < rem ? 1 : 0);
    n_iter = rg < n_ranges ? (r_end - r_beg + kBM - 1) / kBM : 0;
  } else {
    n_iter = cluster_id < n_groups ? (n_groups - cluster_id + n_clusters - 1) / n_clusters : 0;
  }
  auto tile_m0 = [&](int i) {
    if (a.balanced) return r_beg + i * kBM;
    const int g = cluster_id + i * n_clusters;
    return ((g % groups_m) * cs + rank) * kBM;
  };
  auto tile_n0 = [&](int i) { return a.balanced ? my_n0 : ((cluster_id + i * n_clusters) / groups_m) * a.BN; };

  float a_scale = a.a_scale;
  if (F16S && a.bound_ptr) {
    pdl_wait();   // the bound comes from a predecessor kernel
    const float bnd = fmaf(__ldg(a.bound_ptr), a.bound_mul, a.bound_add);
    int e = 11;   // a_scale = the power of two that brings the bound just below 2^15 (clamped to 2^-10 .. 2^10)
    if (bnd > 0.f) frexpf(32768.f / bnd, &e);
    a_scale = ldexpf(1.f, max(-10, min(10, e - 1)));
  }
  for (int i = tid; i < a.N; i += (int)blockDim.x) {
    s_par[i] = a.bias ? a.bias[i] : 0.f;
    if (a.ln) s_par[a.N + i] = a.gamma[i], s_par[2 * a.N + i] = a.beta[i];
    if (F16S) s_par[3 * a.N + i] = a.bound_ptr ? a.cscale[i] / a_scale : (a.cscale ? a.cscale[i] : 1.f / a_scale);
  }
  if (tid == 0) {
    // SPLIT: a slot is free again once the MMAs that read it have completed AND every converter warp has passed it
    // (also the residual slots, which the converters do not touch): no waiter can then fall a whole phase behind
    for (int i = 0; i < SA; ++i) mbar_init(smem_u32(&s_fa[i]), 1), mbar_init(smem_u32(&s_ea[i]), MODE ? 1 + CW : 1);
    for (int i = 0; i < SW; ++i) mbar_init(smem_u32(&s_fw[i]), 1), mbar_init(smem_u32(&s_ew[i]), cs);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&s_tfull[i]), 1), mbar_init(smem_u32(&s_tempty[i]), kEpiWarps);
    for (int i = 0; i < SL; ++i) mbar_init(smem_u32(&s_fl[i]), CW), mbar_init(smem_u32(&s_el[i]), 1);
    mbar_init_fence();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // every CTA's barriers exist before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  pdl_trigger();   // the next kernel in the stream may be scheduled onto SMs as they drain
  if (a.trace && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(size_t)blockIdx.x * kTraceSlots] = t;
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer: per tile the residual boxes
    // (copied into the accumulator by the MMA lane), then the A k-blocks, all through the same ring
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      if (a.res_chunks) tma_prefetch_desc(&map_r);
      // A and the residual come from the predecessor kernel (and this kernel's outputs may alias buffers it still
      // reads): everything up to here -- barriers, TMEM, parameters, the resident weight tile -- overlapped its tail
      pdl_wait();
      if (a.stagger_ns && (cluster_id & 1)) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        } while (t - t0 < (unsigned long long)a.stagger_ns);
      }
      // Split modes hold few A k-blocks in flight (the W_hi / W_lo and a_lo rings take the shared memory), too few to cover the
      // DRAM latency: a tile's 128 rows (contiguous in A, and in the residual when its rows are dense) are pulled into L2 as
      // one sequential stream a whole tile ahead, so the k-block loads only see L2 latency
      // (issued in k_blocks slices next to the k-block loads: a tile-sized prefetch in one go would sit in front of them in
      // the copy engine's queue)
      auto prefetch_slice = [&](int i, int kb) {
        if (!MODE || !a.A || i >= n_iter) return;
        const int m0 = tile_m0(i);
        if (m0 >= a.M) return;
        const size_t rows = (size_t)min(kBM, a.M - m0);
        const char* p = reinterpret_cast<const char*>(a.A) + (size_t)m0 * a.K * 4 + (size_t)kb * rows * 128;
        bulk_prefetch_l2(p, (uint32_t)(rows * 128));
        if (a.residual && a.ldr == a.N && a.N == a.BN && kb * 4 < a.N / 8) {   // residual tile: N * 4 / 128 slices of rows x 128 B
          const char* r = reinterpret_cast<const char*>(a.residual) + (size_t)m0 * a.N * 4 + (size_t)kb * rows * 128;
          if ((size_t)(kb + 1) * 128 <= (size_t)a.N * 4) bulk_prefetch_l2(r, (uint32_t)(rows * 128));
        }
      };
      auto prefetch_tile = [&](int i) {
        for (int kb = 0; kb < k_blocks; ++kb) prefetch_slice(i, kb);
      };
      prefetch_tile(0);
      int stage = 0, ev = 0;
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
          tma_load_2d(sm_a + (uint32_t)stage * a_bytes, &map_a, bar, kb * a.kb_elems, m0);
          prefetch_slice(i + 1, kb);
          if (!F16S) trace_event(a, 0, ev++);
          if (++stage == SA) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ W producer
    if (lane == 0) {
      tma_prefetch_desc(&map_w);
      if (a.w_res) {
        // the whole W tile (all k-blocks) fits: load it once, it is never released
        if (n_iter > 0)
          for (int kb = 0; kb < k_blocks; ++kb) {
            const uint32_t bar = smem_u32(&s_fw[kb]);
            mbar_arrive_expect_tx(bar, w_bytes);
            tma_load_2d(sm_w + (uint32_t)kb * w_bytes, &map_w, bar, kb * a.kb_elems, my_n0);
          }
      } else {
        int stage = 0;
        uint32_t phase = 0;
        if (MODE) tma_prefetch_desc(&map_wl);
        for (int i = 0; i < n_iter; ++i) {
          const int n0 = tile_n0(i);
          for (int kb = 0; kb < w_blocks; ++kb) {
#pragma unroll
            for (int part = 0; part < (MODE ? 2 : 1); ++part) {   // split modes: the hi k-block, then the lo k-block
              const CUtensorMap* mw = part ? &map_wl : &map_w;
              mbar_wait(smem_u32(&s_ew[stage]), phase ^ 1u);   // every CTA of the cluster has consumed the slot
              const uint32_t bar = smem_u32(&s_fw[stage]);
              const uint32_t dst = sm_w + (uint32_t)stage * w_bytes;
              mbar_arrive_expect_tx(bar, w_bytes);             // all `cs` slices of the W k-block
              if (cs == 1)
                tma_load_2d(dst, mw, bar, kb * w_elems, n0);
              else
                tma_load_2d_multicast(dst + (uint32_t)rank * w_slice_bytes, mw, bar, kb * w_elems,
                                      n0 + rank * (int)w_slice_rows, cta_mask);
              if (++stage == SW) stage = 0, phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (a.f16 || F16S) ? idesc_f16(a.BN) : idesc_tf32(a.BN);
      int sa = 0, sw = 0, sl = 0, ev = 0;
      uint32_t pa = 0, pw = 0, pl = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int acc = it & 1;
        mbar_wait(smem_u32(&s_tempty[acc]), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        // accumulator <- residual tile, 32 columns per ring slot (four copies of eight columns each)
        for (int rc = 0; rc < a.res_chunks; ++rc) {
          mbar_wait(smem_u32(&s_fa[sa]), pa);
          tc_fence_after();
          const uint64_t rdesc = smem_desc_k128(sm_a + (uint32_t)sa * a_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) tmem_cp_128x256b(tmem_d + (uint32_t)(rc * 32 + k * 8), rdesc + (uint64_t)(k * 2));
          mma_commit(smem_u32(&s_ea[sa]));
          if (++sa == SA) sa = 0, pa ^= 1u;
        }
        const uint32_t acc0 = a.res_chunks ? 1u : 0u;       // accumulate on top of the preloaded residual
        if (F16S) {
          for (int kb = 0; kb < w_blocks; ++kb) {
            trace_event(a, 1, ev++);
            mbar_wait(smem_u32(&s_fl[sl]), pl);             // operand slot: both A k-blocks converted
            trace_event(a, 1, ev++);
            // the two A slots were only read by the converters: this lane's arrival is the one the MMA commit is elsewhere
            mbar_arrive(smem_u32(&s_ea[sa]));
            if (++sa == SA) sa = 0, pa ^= 1u;
            mbar_arrive(smem_u32(&s_ea[sa]));
            if (++sa == SA) sa = 0, pa ^= 1u;
            mbar_wait(smem_u32(&s_fw[sw]), pw);             // W_hi k-block
            trace_event(a, 1, ev++);
            tc_fence_after();
            const uint64_t hdesc = smem_desc_k128(sm_l + (uint32_t)sl * l_bytes);
            const uint64_t ldesc = smem_desc_k128(sm_l + (uint32_t)sl * l_bytes + a_bytes);
            uint64_t bdesc = smem_desc_k128(sm_w + (uint32_t)sw * w_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16(tmem_d, hdesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : acc0);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_f16(tmem_d, ldesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
            if (cs == 1)
              mma_commit(smem_u32(&s_ew[sw]));
            else
              mma_commit_multicast(smem_u32(&s_ew[sw]), cta_mask);
            if (++sw == SW) sw = 0, pw ^= 1u;
            mbar_wait(smem_u32(&s_fw[sw]), pw);             // W_lo k-block
            trace_event(a, 1, ev++);
            tc_fence_after();
            bdesc = smem_desc_k128(sm_w + (uint32_t)sw * w_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_f16(tmem_d, hdesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
            if (cs == 1)
              mma_commit(smem_u32(&s_ew[sw]));
            else
              mma_commit_multicast(smem_u32(&s_ew[sw]), cta_mask);
            if (++sw == SW) sw = 0, pw ^= 1u;
            mma_commit(smem_u32(&s_el[sl]));
            if (++sl == SL) sl = 0, pl ^= 1u;
          }
        } else if (SPLIT) {
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(smem_u32(&s_fa[sa]), pa);
            trace_event(a, 1, ev++);
            mbar_wait(smem_u32(&s_fl[sl]), pl);             // the converters have masked the A slot and filled the lo slot
            trace_event(a, 1, ev++);
            mbar_wait(smem_u32(&s_fw[sw]), pw);             // W_hi k-block
            trace_event(a, 1, ev++);
            tc_fence_after();
            const uint64_t adesc = smem_desc_k128(sm_a + (uint32_t)sa * a_bytes);
            const uint64_t ldesc = smem_desc_k128(sm_l + (uint32_t)sl * a_bytes);
            uint64_t bdesc = smem_desc_k128(sm_w + (uint32_t)sw * w_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : acc0);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32(tmem_d, ldesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
            if (cs == 1)
              mma_commit(smem_u32(&s_ew[sw]));
            else
              mma_commit_multicast(smem_u32(&s_ew[sw]), cta_mask);
            if (++sw == SW) sw = 0, pw ^= 1u;
            mbar_wait(smem_u32(&s_fw[sw]), pw);             // W_lo k-block
            trace_event(a, 1, ev++);
            tc_fence_after();
            bdesc = smem_desc_k128(sm_w + (uint32_t)sw * w_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
            if (cs == 1)
              mma_commit(smem_u32(&s_ew[sw]));
            else
              mma_commit_multicast(smem_u32(&s_ew[sw]), cta_mask);
            if (++sw == SW) sw = 0, pw ^= 1u;
            mma_commit(smem_u32(&s_ea[sa]));
            mma_commit(smem_u32(&s_el[sl]));
            if (++sa == SA) sa = 0, pa ^= 1u;
            if (++sl == SL) sl = 0, pl ^= 1u;
          }
        } else
        for (int kb = 0; kb < k_blocks; ++kb) {
          if (a.w_res) sw = kb, pw = 0;                     // resident: slot kb, its only phase
          mbar_wait(smem_u32(&s_fw[sw]), pw);
          mbar_wait(smem_u32(&s_fa[sa]), pa);
          trace_event(a, 1, ev++);
          tc_fence_after();
          const uint64_t adesc = smem_desc_k128(sm_a + (uint32_t)sa * a_bytes);
          const uint64_t bdesc = smem_desc_k128(sm_w + (uint32_t)sw * w_bytes);
          // 8 tf32 / 16 fp16 = 32 bytes per MMA along K: advance the descriptors' start address by 32 B
          if (a.f16) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : acc0);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_tf32(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : acc0);
          }
          mma_commit(smem_u32(&s_ea[sa]));
          if (++sa == SA) sa = 0, pa ^= 1u;
          if (!a.w_res) {
            if (cs == 1)
              mma_commit(smem_u32(&s_ew[sw]));
            else
              mma_commit_multicast(smem_u32(&s_ew[sw]), cta_mask);   // the W slot is shared: release it everywhere
            if (++sw == SW) sw = 0, pw ^= 1u;
          }
        }
        mma_commit(smem_u32(&s_tfull[acc]));
      }
    }
  } else if (F16S && (warp == 2 || warp >= 4 + kEpiWarps)) {
    // ------------------------------------------------------------------ fp16 x3: two fp32 A k-blocks -> {a_hi, a_lo} fp16 tiles
    const int ct = (warp == 2 ? 0 : warp - (4 + kEpiWarps) + 1) * 32 + lane;   // 0 .. 32 CW - 1
    // A k-block = 1024 16-byte chunks (4 floats) over the 32 CW converter threads.  Chunk sidx sits at stored position sidx & 7
    // of row `row` (128-byte swizzle: logical chunk c = position ^ (row & 7)); its four fp16 values go to 16-byte chunk
    // 4 hsel + c / 2 (XOR-swizzled with the row) of the operand tile's row, half c & 1 of it.  The rows of one warp
    // instruction are permuted so that their fp16 stores fall into both 64-byte halves of the shared-memory banks.  Source
    // and destination offsets do not depend on the k-block: computed once, packed into one register per chunk
    // (destination for hsel = 1: XOR 64).
    constexpr int kPer = (1024 + 32 * CW - 1) / (32 * CW);
    uint32_t offs[kPer];
#pragma unroll
    for (int m = 0; m < kPer; ++m) {
      const uint32_t sidx = min((uint32_t)ct + 32u * CW * m, 1023u);
      const uint32_t rr = sidx >> 3, row = (rr & ~7u) | ((rr & 1u) << 2) | ((rr & 6u) >> 1), cp = sidx & 7u;
      const uint32_t c = cp ^ (row & 7u);
      offs[m] = (row * 128u + cp * 16u) | ((row * 128u + (((c >> 1) ^ (row & 7u)) << 4) + ((c & 1u) << 3)) << 16);
    }
    const int n_mine = ((uint32_t)ct + 32u * CW * (kPer - 1) < 1024u) ? kPer : kPer - 1;
    int cev = 0;
    int sa = 0, sl = 0;
    uint32_t pa = 0, pl = 0;
    for (int it = 0; it < n_iter; ++it) {
      for (int rc = 0; rc < a.res_chunks; ++rc) {          // residual boxes pass through the ring untouched
        mbar_wait(smem_u32(&s_fa[sa]), pa);
        if (lane == 0) mbar_arrive(smem_u32(&s_ea[sa]));
        if (++sa == SA) sa = 0, pa ^= 1u;
      }
      for (int kb = 0; kb < w_blocks; ++kb) {
        mbar_wait(smem_u32(&s_el[sl]), pl ^ 1u);
        if (ct == 0) trace_event(a, 0, cev++);               // (MODE 2 traces its first converter warp in the producer's slots)
        unsigned char* hi_p = smem + (size_t)SA * a_bytes + (size_t)sl * l_bytes;
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {              // the two 32-float k-blocks of this 64-element operand slot
          mbar_wait(smem_u32(&s_fa[sa]), pa);
          if (ct == 0) trace_event(a, 0, cev++);
          const unsigned char* a_slot = smem + (size_t)sa * a_bytes;
          // all of a thread's loads first (plain C++ accesses: they stay in flight together), then conversions and stores
          float4 xs[kPer];
#pragma unroll
          for (int m = 0; m < kPer; ++m) xs[m] = *reinterpret_cast<const float4*>(a_slot + (offs[m] & 0xffffu));
#pragma unroll
          for (int m = 0; m < kPer; ++m) {
            const float4 x = make_float4(xs[m].x * a_scale, xs[m].y * a_scale, xs[m].z * a_scale, xs[m].w * a_scale);
            const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2half2_rn(x.z - f23.x, x.w - f23.y);
            const uint32_t off = (offs[m] >> 16) ^ (hsel ? 64u : 0u);
            if (m < n_mine) {
              *reinterpret_cast<uint2*>(hi_p + off) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
              *reinterpret_cast<uint2*>(hi_p + a_bytes + off) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
          }
          __syncwarp();                                      // every lane has read the A slot
          if (lane == 0) mbar_arrive(smem_u32(&s_ea[sa]));
          if (++sa == SA) sa = 0, pa ^= 1u;
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_fl[sl]));
        if (ct == 0) trace_event(a, 0, cev++);
        if (++sl == SL) sl = 0, pl ^= 1u;
      }
    }
  } else if (SPLIT && (warp == 2 || warp >= 4 + kEpiWarps)) {
    // ------------------------------------------------------------------ 3xTF32: A k-block -> (a_hi in place, a_lo slot)
    const int ct = (warp == 2 ? 0 : warp - (4 + kEpiWarps) + 1) * 32 + lane;   // 0 .. 32 kConvWarps - 1
    int sa = 0, sl = 0;
    uint32_t pa = 0, pl = 0;
    for (int it = 0; it < n_iter; ++it) {
      for (int rc = 0; rc < a.res_chunks; ++rc) {          // residual boxes pass through the ring untouched
        mbar_wait(smem_u32(&s_fa[sa]), pa);
        if (lane == 0) mbar_arrive(smem_u32(&s_ea[sa]));
        if (++sa == SA) sa = 0, pa ^= 1u;
      }
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(smem_u32(&s_fa[sa]), pa);
        mbar_wait(smem_u32(&s_el[sl]), pl ^ 1u);
        // 1024 16-byte chunks over the 96 converter threads: all of a thread's loads first (plain C++ accesses, so the
        // compiler keeps them in flight together), then the splits and stores
        constexpr int kPer = (1024 + 32 * kConvWarps - 1) / (32 * kConvWarps);
        unsigned char* a_slot = smem + (size_t)sa * a_bytes;
        unsigned char* l_slot = smem + (size_t)SA * a_bytes + (size_t)sl * a_bytes;
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
            if (a.x3_inplace) *reinterpret_cast<uint4*>(a_slot + j) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(l_slot + j) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_fl[sl])), mbar_arrive(smem_u32(&s_ea[sa]));
        if (++sa == SA) sa = 0, pa ^= 1u;
        if (++sl == SL) sl = 0, pl ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ------------------------------------------------------------------ epilogue
    // Thread = one row of the tile (its TMEM lane); the four warps of a lane quarter take the 16-column chunks
    // round-robin.  Rows meet global memory through one swizzled 2 KB buffer per warp: coalesced 16-byte accesses
    // on the global side (8 rows x 64 B per instruction), whole rows on the thread side, __syncwarp in between.
    const int ew = warp - 4, q = ew & 3, part = ew >> 2;
    const uint32_t buf = sm_stage_buf + (uint32_t)ew * kStageBuf;
    const int n_chunks = a.BN / kChunk;
    const int crow = lane >> 2, ccol = lane & 3;              // coalesced side: rows crow + 8 i, 16-byte column ccol
    for (int it = 0; it < n_iter; ++it) {
      const int acc = it & 1;
      const int m0 = tile_m0(it), n0 = tile_n0(it);
      const int row0 = m0 + q * 32;                           // this warp's 32 rows
      const uint32_t tbase = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
      const bool rows_live = row0 < r_end;                      // warp-uniform (and equal for the warps of the quarter)

      auto store_rows = [&](int c, const float (&f)[16]) {    // this thread's row chunk -> global
        if (a.out && a.direct_store) {          // fp32 rows: 64 contiguous bytes per thread, two 32-byte stores
          const int row = row0 + lane;
          if (row < r_end) {
            auto put = [&](float* p) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + 8 * j), "f"(f[8 * j]),
                             "f"(f[8 * j + 1]), "f"(f[8 * j + 2]), "f"(f[8 * j + 3]), "f"(f[8 * j + 4]), "f"(f[8 * j + 5]),
                             "f"(f[8 * j + 6]), "f"(f[8 * j + 7])
                             : "memory");
            };
            if (!a.scatter) {
              put(a.out + (size_t)row * a.ldc + n0 + c * kChunk);
            } else {
              const int bi = row / a.sc_rows;
              const int* dl = a.scatter + (size_t)(row - bi * a.sc_rows) * a.sc_r;
              for (int j = 0; j < a.sc_r; ++j) {
                const int d = __ldg(dl + j);
                if (d < 0) break;
                put(a.out + ((size_t)bi * a.sc_dst_rows + d) * a.ldc + n0 + c * kChunk);
              }
            }
          }
        } else if (a.out) {                                   // fp32 rows, coalesced through the buffer
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(swz(buf, lane, j)), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                         "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                         : "memory");
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = row0 + crow + 8 * i;
            float4 o;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                         : "r"(swz(buf, crow + 8 * i, ccol)));
            if (row < r_end) {
              if (!a.scatter) {
                st_stream4(a.out + (size_t)row * a.ldc + n0 + c * kChunk + ccol * 4, o);
              } else {
                const int bi = row / a.sc_rows;
                const int* dl = a.scatter + (size_t)(row - bi * a.sc_rows) * a.sc_r;
                for (int j = 0; j < a.sc_r; ++j) {
                  const int d = __ldg(dl + j);
                  if (d < 0) break;
                  st_stream4(a.out + ((size_t)bi * a.sc_dst_rows + d) * a.ldc + n0 + c * kChunk + ccol * 4, o);
                }
              }
            }
          }
          __syncwarp();
        }
        if (a.out16 && row0 + lane < r_end)                     // fp16 copy: 32 bytes of this thread's row, one store
          st_half16(a.out16 + (size_t)(row0 + lane) * a.ldc16 + n0 + c * kChunk, f);
      };
      auto params4 = [&](const float* p, int j) {             // 16-byte broadcast read of bias / gamma / beta
        return *reinterpret_cast<const float4*>(p + 4 * j);
      };

      mbar_wait(smem_u32(&s_tfull[acc]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (ew == 0 && lane == 0) trace_event(a, 2, 2 * it);

      float sum = 0.f, sumsq = 0.f;
      // ---- pass A: acc (+ residual, already in the accumulator) + bias; LayerNorm: row statistics only
      //              otherwise: activation and store
      for (int c = part; c < n_chunks; c += 4) {
        uint32_t v[16];
        tmem_ld16(tbase + (uint32_t)(c * kChunk), v);
        if (!rows_live) continue;
        float f[16];
        const float* bias = s_par + n0 + c * kChunk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = params4(bias, j);
          if (F16S) {   // undo the operand scaling (exact powers of two)
            const float4 s4 = params4(s_par + 3 * a.N + n0 + c * kChunk, j);
            f[4 * j] = fmaf(__uint_as_float(v[4 * j]), s4.x, b4.x), f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), s4.y, b4.y);
            f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), s4.z, b4.z), f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), s4.w, b4.w);
          } else {
            f[4 * j] = __uint_as_float(v[4 * j]) + b4.x, f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z, f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
        }
        if (F16S && a.epi_res && row0 + lane < r_end) {   // residual: 64 contiguous bytes of this thread's row
          const float* rp = a.residual + (size_t)(row0 + lane) * a.ldr + n0 + c * kChunk;
          float r8[8];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            ld_stream8(rp + 8 * j, r8);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[8 * j + e] += r8[e];
          }
        }
        if (a.ln) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum += f[j];
            sumsq = fmaf(f[j], f[j], sumsq);
          }
          if (F16S) {   // park the unscaled pre-LayerNorm values for pass B
            uint32_t w16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w16[j] = __float_as_uint(f[j]);
            tmem_st16(tbase + (uint32_t)(c * kChunk), w16);
          }
        } else if (a.planes) {
          // fp16 head-major planes: chunk c of the row is half (c & 1) of head (n0 / 32 + c / 2) of token (row % Nv)
          // in group row / Nv
          const int row = row0 + lane;
          if (row < r_end) {
            const int gi = row / a.Nv, tok = row - gi * a.Nv;
            st_half16(a.planes + (((int64_t)gi * a.H + (n0 / 32 + (c >> 1))) * a.Nv + tok) * 32 + (c & 1) * 16, f);
          }
        } else if (a.planes32) {
          // fp32 half-head planes: chunk c of the row is half-plane (n0 / 16 + c) of token (row % Nv) in group row / Nv
          const int row = row0 + lane;
          if (row < r_end) {
            const int gi = row / a.Nv, tok = row - gi * a.Nv;
            float* p = a.planes32 + (((int64_t)gi * (a.N / 16) + (n0 / 16 + c)) * a.Nv + tok) * 16;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + 8 * j), "f"(f[8 * j]),
                           "f"(f[8 * j + 1]), "f"(f[8 * j + 2]), "f"(f[8 * j + 3]), "f"(f[8 * j + 4]), "f"(f[8 * j + 5]),
                           "f"(f[8 * j + 6]), "f"(f[8 * j + 7])
                           : "memory");
          }
        } else {
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          store_rows(c, f);
        }
      }
      // ---- pass B (LayerNorm): the four warps of the quarter exchange their partial statistics, then each
      //      normalises and stores its chunks of the parked row
      if (a.ln) {
        s_stat[part][q * 32 + lane] = make_float2(sum, sumsq);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
        float tsum = 0.f, tsq = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 o = s_stat[p][q * 32 + lane];
          tsum += o.x, tsq += o.y;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // s_stat is reused by the next tile
        if (rows_live) {
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
              const float4 g4 = params4(gam, j), b4 = params4(bet, j);
              const float4 c4 = F16S ? make_float4(0.f, 0.f, 0.f, 0.f) : params4(bias, j);   // (MODE 2: already added in pass A)
              f[4 * j] = (__uint_as_float(v[4 * j]) + c4.x - mean) * rstd * g4.x + b4.x;
              f[4 * j + 1] = (__uint_as_float(v[4 * j + 1]) + c4.y - mean) * rstd * g4.y + b4.y;
              f[4 * j + 2] = (__uint_as_float(v[4 * j + 2]) + c4.z - mean) * rstd * g4.z + b4.z;
              f[4 * j + 3] = (__uint_as_float(v[4 * j + 3]) + c4.w - mean) * rstd * g4.w + b4.w;
            }
            store_rows(c, f);
          }
        }
      }
      // the accumulator is drained
      tc_fence_before();
      __syncwarp();
      if (ew == 0 && lane == 0) trace_event(a, 2, 2 * it + 1);
      if (lane == 0) mbar_arrive(smem_u32(&s_tempty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace ub

namespace ub {
int launch_x3_pair(const char* fn, const float* A, const float* W_hi, const float* W_lo, const float* bias, const float* residual,
                   int ldr, const float* gamma, const float* beta, float eps, float* out, int ldc, float* planes32, int Nv,
                   int M, int N, int K, int relu, int ln, cudaStream_t stream);   // gemm_pair.cu
}

using namespace ub;

static int g_gemm_cluster = 4;
static int g_x3_prefetch = 0;   // split modes: L2 prefetch of the next row tile (A and residual); measured neutral: off
extern "C" int ub_set_gemm_x3_prefetch(int on) {
  g_x3_prefetch = on ? 1 : 0;
  return UB_OK;
}
static int g_x3_pair = 0;   // 3xTF32 on CTA pairs (tcgen05.mma.cta_group::2, gemm_pair.cu) where the shape is covered
extern "C" int ub_set_gemm_x3_pair(int on) {
  g_x3_pair = on ? 1 : 0;
  return UB_OK;
}
static int g_x3_inplace = 1, g_x3_direct = 1, g_x3_cluster = 2, g_x3_stagger_ns = 0;   // measured best (profiles/r2_gemm_x3.txt)
// A/B knobs of the 3xTF32 mode (tools/bench_gemm_x3.py): in-place a_hi write-back, direct result stores, cluster size,
// start offset of every other cluster
extern "C" int ub_set_gemm_x3(int inplace, int direct_store, int cluster, int stagger_ns) {
  UB_REQUIRE(cluster == 1 || cluster == 2 || cluster == 4, "ub_set_gemm_x3: cluster size must be 1, 2 or 4");
  UB_REQUIRE(stagger_ns >= 0 && stagger_ns <= 100000, "ub_set_gemm_x3: stagger must be in 0 .. 100 us");
  g_x3_inplace = inplace ? 1 : 0, g_x3_direct = direct_store ? 1 : 0, g_x3_cluster = cluster, g_x3_stagger_ns = stagger_ns;
  return UB_OK;
}
static int g_gemm_stream_w_res = 1;   // stream W (deeper A / residual ring) when a residual rides the ring: 103 vs 107 us
extern "C" int ub_set_gemm_stream_w_with_residual(int on) {
  g_gemm_stream_w_res = on ? 1 : 0;
  return UB_OK;
}
static unsigned long long* g_gemm_trace = nullptr;
// debugging aid (tools/trace_gemm.py): device buffer of 148 x 128 u64 that the next launches fill with timestamps
extern "C" int ub_set_gemm_trace(void* buf) {
  g_gemm_trace = reinterpret_cast<unsigned long long*>(buf);
  return UB_OK;
}
// CTAs per cluster of ub_linear_tf32 (1, 2 or 4): a performance knob, results do not depend on it
extern "C" int ub_set_gemm_cluster(int cs) {
  UB_REQUIRE(cs == 1 || cs == 2 || cs == 4, "ub_set_gemm_cluster: cluster size must be 1, 2 or 4");
  g_gemm_cluster = cs;
  return UB_OK;
}

// Shared launcher.  f16 = 0: A / W fp32 (TF32 MMA);  f16 = 1: A / W fp16.  W_lo != null: 3xTF32 (W = W_hi, see the kernel).
static int launch_linear(const char* fn, int f16, const void* A, const void* W, const float* W_lo, const float* bias,
                         const float* residual, int ldr, const float* gamma, const float* beta, float eps, float* out,
                         int ldc, void* out16, int ldc16, void* planes, float* planes32, int Nv, int M, int N, int K,
                         int flags, ub_stream_t stream, const int* scatter = nullptr, int sc_r = 0, int sc_rows = 0,
                         int sc_dst_rows = 0, bool f16s = false, float a_scale = 1.f, const float* col_scale = nullptr,
                         const float* bound_dev = nullptr, float bound_mul = 1.f, float bound_add = 0.f) {
  UB_REQUIRE(!bound_dev || (f16s && col_scale), "%s: a device-side bound needs the fp16 x3 mode and col_scale", fn);
  // f16s: fp16 x3 (kernel MODE 2): A fp32, W / W_lo fp16 (64-element k-blocks)
  const int relu = flags & 1, ln = (flags >> 1) & 1;
  const int kb_elems = f16 ? 64 : 32, esize = f16 ? 2 : 4;
  const bool split = W_lo != nullptr && !f16s;
  UB_REQUIRE(A && W && (out || out16 || planes || planes32), "%s: null pointer", fn);
  UB_REQUIRE(!(split && f16), "%s: the 3xTF32 mode takes fp32 operands", fn);
  UB_REQUIRE(!scatter || (out && !out16 && !planes && !planes32 && !ln && sc_r > 0 && sc_rows > 0 && M % sc_rows == 0 &&
                          sc_dst_rows > 0),
             "%s: a row scatter takes a plain fp32 output and M %% rows_per_item == 0", fn);
  UB_REQUIRE(M > 0 && N > 0 && K > 0, "%s: non-positive dimension", fn);
  UB_REQUIRE(!ln || (gamma && beta), "%s: layernorm needs gamma and beta", fn);
  UB_REQUIRE_ALIGNED16(A);
  UB_REQUIRE_ALIGNED16(W);
  if (K % (f16s ? 64 : kb_elems) != 0 || N % 32 != 0 || (N > 256 && N % 256 != 0) || (ln && N > 256) ||
      (planes && (ln || relu || residual || out16 || planes32)) || ((planes || planes32) && (Nv <= 0 || M % Nv != 0)) ||
      (planes32 && (ln || relu || residual || out16 || out || (reinterpret_cast<uintptr_t>(planes32) & 31u))) ||
      ((split || f16s) && (reinterpret_cast<uintptr_t>(W_lo) & 15u)) ||
      (f16s && (out16 || planes || (out && (ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 31u))) ||
                (residual && (ldr % 8 != 0 || (reinterpret_cast<uintptr_t>(residual) & 31u))) || !(a_scale > 0.f))) ||
      (out && (ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(out) & 15u))) ||
      (out16 && (ldc16 % 16 != 0 || (reinterpret_cast<uintptr_t>(out16) & 31u))) ||
      (planes && (reinterpret_cast<uintptr_t>(planes) & 31u)) ||
      (residual && (ldr % 4 != 0 || (reinterpret_cast<uintptr_t>(residual) & 15u))) || N > 1024) {
    set_error("%s: shape not covered (M=%d N=%d K=%d flags=%d)", fn, M, N, K, flags);
    return ub::unsupported();
  }
  if (split && g_x3_pair && !scatter && !out16 && !planes && M >= 2 * kBM * 16) {
    const int rc = launch_x3_pair(fn, reinterpret_cast<const float*>(A), reinterpret_cast<const float*>(W), W_lo, bias, residual,
                                  ldr, gamma, beta, eps, out, ldc, planes32, Nv, M, N, K, relu, ln, (cudaStream_t)stream);
    if (rc != UB_EUNSUPPORTED) return rc;   // (not covered: the single-CTA kernel below; not a generic fallback)
  }
  GemmArgs a;
  a.bias = bias, a.gamma = gamma, a.beta = beta, a.planes = reinterpret_cast<__half*>(planes);
  a.planes32 = planes32, a.SL = (split || f16s) ? 2 : 0;
  a.scatter = scatter, a.sc_r = sc_r, a.sc_rows = sc_rows, a.sc_dst_rows = sc_dst_rows;
  a.a_scale = a_scale, a.cscale = col_scale, a.epi_res = 0;
  a.bound_ptr = bound_dev, a.bound_mul = bound_mul, a.bound_add = bound_add;
  a.x3_inplace = g_x3_inplace, a.stagger_ns = split ? g_x3_stagger_ns : 0;
  a.direct_store = f16s || (split && g_x3_direct && out && ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0);
  a.Nv = Nv, a.H = N / 32;
  a.M = M, a.N = N, a.K = K, a.BN = N > 256 ? 256 : N;
  a.n_tiles_m = (M + kBM - 1) / kBM, a.n_tiles_n = N / a.BN;
  a.eps = eps, a.relu = relu, a.ln = ln;
  a.A = (split || f16s) && g_x3_prefetch ? reinterpret_cast<const float*>(A) : nullptr;
  a.residual = residual, a.out = out, a.ldr = ldr, a.ldc = ldc;
  a.out16 = reinterpret_cast<__half*>(out16), a.ldc16 = ldc16;
  a.f16 = f16, a.kb_elems = kb_elems;
  a.trace = g_gemm_trace;
  const int k_blocks = K / kb_elems;
  const bool direct_ok = out && ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0;
  const bool stage_free = f16s || (split && ((g_x3_direct && direct_ok) || planes32));   // results leave straight from registers
  a.no_staging = stage_free;
  const size_t fixed = (stage_free ? 0 : kEpiWarps * kStageBuf) + (size_t)(f16s ? 4 : 3) * N * sizeof(float);   // (MODE 2 stores straight from registers)
  const size_t budget = 232448 - 5120 - 1024;   // minus static shared memory and slack
  // W resident: every k-block of the (single) W tile stays in shared memory, with at least 3 A stages next to it
  // (with several column tiles: one CTA per column tile and row range, see n_split in the kernel)
  const int n_sms = sm_count();
  const bool split_ok = a.n_tiles_n > 1 && n_sms % a.n_tiles_n == 0 && M >= 4 * n_sms;
  a.w_res = !split && !f16s && (a.n_tiles_n == 1 || split_ok) && k_blocks <= kMaxStages && !(residual && ln && g_gemm_stream_w_res) &&
            fixed + (size_t)k_blocks * a.BN * 128 + 3 * (size_t)kBM * 128 <= budget;
  // cluster size (streaming W only): the W k-block is split into `cs` slices of whole 8-row swizzle groups
  a.cs = a.w_res ? 1 : ((split || f16s) ? g_x3_cluster : g_gemm_cluster);
  while (a.cs > 1 && ((a.BN / a.cs) % 8 != 0 || a.BN % a.cs != 0 || a.n_tiles_m < a.cs)) a.cs >>= 1;
  CUtensorMap ma, mw, mr, mwl;
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)K * esize};
    const uint32_t box[2] = {(uint32_t)kb_elems, kBM};
    if (int rc = make_tensor_map(&ma, dt, 2, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K * (f16s ? 2 : esize)};
    const uint32_t box[2] = {(uint32_t)(f16s ? 64 : kb_elems), (uint32_t)(a.BN / a.cs)};
    const CUtensorMapDataType wdt = f16s ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : dt;
    if (int rc = make_tensor_map(&mw, wdt, 2, W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    mwl = mw;
    if (split || f16s)
      if (int rc = make_tensor_map(&mwl, wdt, 2, W_lo, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  a.res_chunks = 0;
  mr = ma;
  if (residual && f16s) {
    a.epi_res = 1;   // added in the epilogue
  } else if (residual) {   // fp32 boxes of 32 columns x 128 rows, laid out in shared memory like a TF32 A k-block
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ldr * 4};
    const uint32_t box[2] = {32, kBM};
    if (int rc = make_tensor_map(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, residual, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
    a.res_chunks = a.BN / 32;
  }
  a.balanced = (a.n_tiles_n == 1 || (split_ok && a.w_res)) && a.cs == 1 && M >= 4 * n_sms;
  a.n_split = a.balanced ? a.n_tiles_n : 1;
  // ring depths: W resident or shallow, A as deep as the 227 KB of shared memory allow (up to 8)
  a.SW = a.w_res ? k_blocks : (a.BN > 128 ? 3 : 4);
  if (split || f16s) {
    // W slots alternate hi / lo k-blocks: as many as fit next to three A stages and the A_lo ring (at least three:
    // one k-block in use, half of the next in flight); MODE 2: the SL operand slots are twice as large
    const size_t rest = budget - fixed - (size_t)(3 + a.SL * (f16s ? 2 : 1)) * kBM * 128;
    a.SW = (int)(rest / ((size_t)a.BN * 128));
    if (a.SW > 6) a.SW = 6;
    if (a.SW < 3) {
      set_error("%s: no room for the operand rings (N=%d)", fn, N);
      return ub::unsupported();
    }
  }
  a.SA = (int)((budget - fixed - (size_t)a.SW * a.BN * 128) / (kBM * 128)) - a.SL * (f16s ? 2 : 1);
  if (a.SA > kMaxStages) a.SA = kMaxStages;
  if ((split || f16s) && a.SA > 4) a.SA = 4;   // a k-block lasts three MMA groups: a shallow ring already covers the DRAM latency
  if (a.SA < 2) {
    set_error("%s: no room for the operand rings (N=%d)", fn, N);
    return ub::unsupported();
  }
  const size_t smem = (size_t)(a.SA + a.SL * (f16s ? 2 : 1)) * kBM * 128 + (size_t)a.SW * a.BN * 128 + fixed;
  const int mode = f16s ? 2 : (split ? 1 : 0);
  auto kernel = f16s ? gemm_tf32_kernel<2> : (split ? gemm_tf32_kernel<1> : gemm_tf32_kernel<0>);
  if (int rc = ensure_smem(kernel, smem, fn)) return rc;
  const int n_groups = ((a.n_tiles_m + a.cs - 1) / a.cs) * a.n_tiles_n;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(mode == 2 ? kGemmThreadsF16S : (mode ? kGemmThreadsSplit : kGemmThreads));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.cs, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // persistent grid: as many clusters as can be resident at once (one CTA per SM; clusters do not span GPCs)
  // (a property of the kernel variant, cluster size and shared-memory size; every B200 answers the same)
  static std::atomic<int> max_clusters[3][5] = {};
  static std::atomic<size_t> max_clusters_smem[3][5] = {};
  if (max_clusters[mode][a.cs].load() == 0 || max_clusters_smem[mode][a.cs].load() != smem) {
    cfg.gridDim = dim3(n_sms / a.cs * a.cs);
    int n = 0;
    if (a.cs == 1 || cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = a.cs == 1 ? n_sms : (n_sms / a.cs) * 3 / 4;
    }
    max_clusters[mode][a.cs].store(n), max_clusters_smem[mode][a.cs].store(smem);
  }
  int clusters = max_clusters[mode][a.cs].load();
  if (clusters > n_groups && !a.balanced) clusters = n_groups;
  cfg.gridDim = dim3(clusters * a.cs);
  if (cudaLaunchKernelEx(&cfg, kernel, a, ma, mw, mr, mwl) != cudaSuccess) {
    set_error("%s: launch failed: %s", fn, cudaGetErrorString(cudaGetLastError()));
    return UB_ECUDA;
  }
  return check_launch(fn);
}

// out = epilogue(A (M, K) @ W (N, K)^T), fp32 operands (TF32 MMA).  flags: bit 0 relu, bit 1 layernorm (gamma, beta,
// N <= 256).  planes != NULL: fp16 head-major output (G = M / Nv groups, H = N / 32 heads), `out` ignored.
extern "C" int ub_linear_tf32(const float* A, const float* W, const float* bias, const float* residual, int ldr,
                              const float* gamma, const float* beta, float eps, float* out, int ldc, void* planes,
                              int Nv, int M, int N, int K, int flags, ub_stream_t stream) {
  return launch_linear("ub_linear_tf32", 0, A, W, nullptr, bias, residual, ldr, gamma, beta, eps, out, ldc, nullptr, 0, planes,
                       nullptr, Nv, M, N, K, flags, stream);
}
// ub_linear_tf32 that also writes the fp16 copy `out16` (row stride ldc16) of the result rows
extern "C" int ub_linear_tf32_dual(const float* A, const float* W, const float* bias, const float* residual, int ldr,
                                   const float* gamma, const float* beta, float eps, float* out, int ldc, void* out16,
                                   int ldc16, int M, int N, int K, int flags, ub_stream_t stream) {
  return launch_linear("ub_linear_tf32_dual", 0, A, W, nullptr, bias, residual, ldr, gamma, beta, eps, out, ldc, out16, ldc16,
                       nullptr, nullptr, 0, M, N, K, flags, stream);
}

// The same with fp16 operands A (M, K) / W (N, K) (same 11-bit significand as TF32, half the bytes: W usually stays
// resident in shared memory) and an optional fp16 copy `out16` of the result, the A operand of the next projection.
extern "C" int ub_linear_f16(const void* A16, const void* W16, const float* bias, const float* residual, int ldr,
                             const float* gamma, const float* beta, float eps, float* out, int ldc, void* out16, int ldc16,
                             void* planes, int Nv, int M, int N, int K, int flags, ub_stream_t stream) {
  return launch_linear("ub_linear_f16", 1, A16, W16, nullptr, bias, residual, ldr, gamma, beta, eps, out, ldc, out16, ldc16,
                       planes, nullptr, Nv, M, N, K, flags, stream);
}

// fp32-grade projection on the tensor cores (3xTF32, see the kernel): out = epilogue(A (M, K) @ (W_hi + W_lo) (N, K)^T) with
// A, W_hi, W_lo fp32; W_hi / W_lo from ub_split_tf32.  Same epilogues as ub_linear_tf32; planes32 != NULL: fp32 half-head
// planes (G = M / Nv, N / 16, Nv, 16) for the fp32 window-staged sampling kernels, `out` ignored.
extern "C" int ub_linear_tf32x3(const float* A, const float* W_hi, const float* W_lo, const float* bias, const float* residual,
                                int ldr, const float* gamma, const float* beta, float eps, float* out, int ldc,
                                float* planes32, int Nv, int M, int N, int K, int flags, ub_stream_t stream) {
  UB_REQUIRE(W_lo, "ub_linear_tf32x3: null pointer");
  return launch_linear("ub_linear_tf32x3", 0, A, W_hi, W_lo, bias, residual, ldr, gamma, beta, eps, planes32 ? nullptr : out, ldc,
                       nullptr, 0, nullptr, planes32, Nv, M, N, K, flags, stream);
}
// ub_linear_tf32x3 whose result rows leave in another order: row b * rows_per_item + q of the product is written to the
// rows b * dst_rows_per_item + scatter[q * scatter_r + j] of `out` for j = 0, 1, ... up to the first negative entry (none:
// the row is dropped).  The camera cross-attention's offset|logit projection writes its rows in hit-list order this way
// (scatter = q_dst of ub_hit_order), so that the sampling kernel reads them as TMA tiles.
extern "C" int ub_linear_tf32x3_scatter(const float* A, const float* W_hi, const float* W_lo, const float* bias, float* out,
                                        int ldc, const int* scatter, int scatter_r, int rows_per_item, int dst_rows_per_item,
                                        int M, int N, int K, ub_stream_t stream) {
  UB_REQUIRE(W_lo && scatter, "ub_linear_tf32x3_scatter: null pointer");
  return launch_linear("ub_linear_tf32x3_scatter", 0, A, W_hi, W_lo, bias, nullptr, 0, nullptr, nullptr, 0.f, out, ldc, nullptr,
                       0, nullptr, nullptr, 0, M, N, K, 0, stream, scatter, scatter_r, rows_per_item, dst_rows_per_item);
}

// The same on kind::f16 MMAs (fp16 x3): every product as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with x_hi = fp16(x), x_lo =
// fp16(x - x_hi) -- two 11-bit significands, the precision class of ub_linear_tf32x3, at twice the tensor-core rate.
// VALID ONLY where every |A| element (and |W| element) is below the fp16 range (65504): the caller guarantees that bound
// (unibev_b200/plugin/fused.py derives it from the weights); values below 2^-14 x 2^11 lose relative (not absolute) precision.
// A (M, K) fp32; W16_hi / W16_lo (N, K) fp16 from ub_split_f16; K % 64 == 0; epilogues as ub_linear_tf32x3 (+ optional row
// scatter as ub_linear_tf32x3_scatter when scatter != NULL).
extern "C" int ub_linear_f16x3(const float* A, float a_scale, const void* W16_hi, const void* W16_lo, const float* col_scale,
                               const float* bias, const float* residual, int ldr, const float* gamma, const float* beta,
                               float eps, float* out, int ldc, float* planes32, int Nv, const int* scatter, int scatter_r,
                               int rows_per_item, int dst_rows_per_item, int M, int N, int K, int flags, ub_stream_t stream) {
  UB_REQUIRE(W16_lo, "ub_linear_f16x3: null pointer");
  return launch_linear("ub_linear_f16x3", 0, A, W16_hi, reinterpret_cast<const float*>(W16_lo), bias, residual, ldr, gamma, beta,
                       eps, planes32 ? nullptr : out, ldc, nullptr, 0, nullptr, planes32, Nv, M, N, K, flags, stream, scatter,
                       scatter_r, rows_per_item, dst_rows_per_item, true, a_scale, col_scale);
}
// ub_linear_f16x3 whose activation bound is only known on the device: |A| <= *bound_dev * bound_mul + bound_add (bound_dev: a
// device float written by an earlier kernel on the stream, e.g. ub_flatten_feats_max).  The kernel derives a_scale from it;
// col_scale = 1 / s_n from ub_split_f16 with a_scale = 1.  No ReLU / scatter variants needed by the encoder here.
extern "C" int ub_linear_f16x3_dyn(const float* A, const float* bound_dev, float bound_mul, float bound_add, const void* W16_hi,
                                   const void* W16_lo, const float* col_scale, const float* bias, const float* residual, int ldr,
                                   const float* gamma, const float* beta, float eps, float* out, int ldc, float* planes32,
                                   int Nv, int M, int N, int K, int flags, ub_stream_t stream) {
  UB_REQUIRE(W16_lo && bound_dev && col_scale, "ub_linear_f16x3_dyn: null pointer");
  UB_REQUIRE(bound_mul >= 0.f && bound_add >= 0.f, "ub_linear_f16x3_dyn: negative bound terms");
  return launch_linear("ub_linear_f16x3_dyn", 0, A, W16_hi, reinterpret_cast<const float*>(W16_lo), bias, residual, ldr, gamma,
                       beta, eps, planes32 ? nullptr : out, ldc, nullptr, 0, nullptr, planes32, Nv, M, N, K, flags, stream,
                       nullptr, 0, 0, 0, true, 1.f, col_scale, bound_dev, bound_mul, bound_add);
}

namespace ub {
// one CTA per weight row: s = the power of two that brings the row's largest magnitude into [4096, 8192), hi = fp16(w s),
// lo = fp16(w s - hi), col_scale = 1 / (s a_scale)
__global__ void __launch_bounds__(128) split_f16_kernel(const float* __restrict__ w, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, float* __restrict__ col_scale, int cols,
                                                        float a_scale) {
  __shared__ float s_max[4];
  const float* row = w + (size_t)blockIdx.x * cols;
  float mx = 0.f;
  for (int i = threadIdx.x; i < cols; i += 128) mx = fmaxf(mx, fabsf(row[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_max[0], s_max[1]), fmaxf(s_max[2], s_max[3]));
  int e = 0;
  if (mx > 0.f && mx < 3.0e38f) frexpf(mx, &e);          // mx = m 2^e, m in [0.5, 1)
  const float s = col_scale ? ldexpf(1.f, 13 - e) : 1.f;  // mx s in [4096, 8192)
  for (int i = threadIdx.x; i < cols; i += 128) {
    const float x = row[i] * s;
    const __half h = __float2half_rn(x);
    hi[(size_t)blockIdx.x * cols + i] = h, lo[(size_t)blockIdx.x * cols + i] = __float2half_rn(x - __half2float(h));
  }
  if (col_scale && threadIdx.x == 0) col_scale[blockIdx.x] = 1.f / (s * a_scale);
}
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = w[i];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - h));
    hi[i] = h, lo[i] = __uint_as_float(l);
  }
}
}  // namespace ub

// w (rows, cols) fp32 -> hi16 / lo16 (rows, cols) fp16 of w[n, :] s_n with s_n the power of two that brings the row's largest
// magnitude into [4096, 8192) (hi = fp16, lo = fp16 of the remainder: both normal numbers for every element within 2^-16 of
// the row maximum), col_scale (rows) = 1 / (s_n a_scale): the epilogue factor of ub_linear_f16x3.  col_scale == NULL: s_n = 1.
extern "C" int ub_split_f16(const float* w, void* hi16, void* lo16, float* col_scale, int rows, int cols, float a_scale,
                            ub_stream_t stream) {
  UB_REQUIRE(w && hi16 && lo16 && rows > 0 && cols > 0 && a_scale > 0.f, "ub_split_f16: bad argument");
  split_f16_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<__half*>(hi16), reinterpret_cast<__half*>(lo16),
                                                           col_scale, cols, a_scale);
  return check_launch("ub_split_f16");
}

// w (n) -> hi = w with the 13 low mantissa bits cleared (exactly what a TF32 MMA reads of w), lo = tf32(w - hi)
extern "C" int ub_split_tf32(const float* w, float* hi, float* lo, int64_t n, ub_stream_t stream) {
  UB_REQUIRE(w && hi && lo && n > 0, "ub_split_tf32: bad argument");
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  split_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, hi, lo, n);
  return check_launch("ub_split_tf32");
}
