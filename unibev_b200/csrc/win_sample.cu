// Window-staged deformable sampling (the fast path of the UniBEV BEV encoder, head dim 32, 4 or 8 points).
//
//   ub_value_to_half       : value map (rows, H*32) fp32 -> head-major fp16 planes (G, H, fH*fW, 32)
//   ub_bev_sample_win_fwd  : BEV self-attention / LiDAR cross-attention sampling                     [R4]
//   ub_build_hits          : per-camera hit lists + 1/count from the visibility bits                 [R3]
//   ub_img_sample_win_fwd  : camera cross-attention sampling, summed over cameras / count            [R3]
//
// Why: with fp32 value rows a bilinear sample moves 4 x 128 B through the SM's L1/shared data pipe (one
// 128-byte wavefront per clock), which caps the fp32 kernels of tile_sample.cu at about a third of the HBM
// roofline.  Here the value map is staged as fp16 planes: one head of one pixel is 64 B, so the two horizontal
// neighbours of a sample are ONE 128-byte shared-memory row segment (2 wavefronts per sample), and the window a
// tile of queries can reach is brought into shared memory by one TMA box copy per (tile, head) -- no demand
// misses, zero padding outside the map for free (TMA fills out-of-bounds box elements with zeros).
//
// Worker warp w of a CTA (16 per CTA, one CTA per SM) owns 16 items (query, head) of the current unit:
//   P1  one lane per SAMPLE (query, point): offsets / logits (the warp's TMA-staged row of the fused
//       offset|logit GEMM output in BEV mode, register-prefetched direct loads in camera mode), softmax across the
//       P adjacent lanes, then a descriptor in the warp's shared-memory slice: a 16-bit pixel index into the
//       window and two half2 words {w_top, w_bottom} x {left, right} (attention weight folded in).
//   P2  8 lanes per ITEM: lanes 0-3 own the left pixel, 4-7 the right one, 8 channels each; per point 1 LDS.32
//       (weights) + 2 LDS.128 (top / bottom pixel pair) + 16 mixed-precision FMAs (fma.rn.f32.f16: fp16 value x
//       fp16 weight, fp32 accumulation); the halves are combined with four shuffles and written as one 128-byte
//       row.
// There is no CTA-wide barrier in the loop: windows are handed over through full / empty mbarriers, so the warps
// drift apart and descriptor building (ALU / SFU bound) of some overlaps the gather (shared-memory bound) of
// others, while the next window streams in behind both (double-buffered in BEV mode).
//
// BEV mode: unit = 16 x 16 BEV queries x one head, handed out by a scheduler warp (atomic counter, re-armed by the
// last CTA) that also streams the windows.
// Samples whose 2 x 2 footprint is not inside the staged window (offsets larger than the halo) take a slow,
// exact path straight from global memory (fp32 weights), so any offsets are handled; the halo is a tuning knob.
//
// Camera mode: unit = 256 hits of one camera x one head; the whole (camera, head) plane plus a one-pixel zero
// halo is the window (reloaded when the CTA's contiguous unit range crosses into a new plane); contributions are
// pre-scaled by 1/count.  The hit lists are split by rank: every query's first camera is written with plain
// stores (pass 0, which also zero-fills the rows no camera sees), only the further cameras of a query (pass 1)
// use red.global.add.v4.f32 -- no zero-fill pass and no read-modify-write for the bulk of the pairs.
#include <cuda_fp16.h>

#include "win_common.cuh"

namespace ub {

static int g_bev_halo = 0;       // 0 = default (P + 1)
int g_bev_halo_shared() { return g_bev_halo; }   // the fp32 twin (win_sample32.cu) honours the same knob
static int g_img_two_win = 0;    // camera mode: double-buffer the plane windows when two fit
static int g_img_vec_ref = 1;
static int g_img_stage = 0;      // camera mode: P1 inputs through bulk async copies instead of register prefetch; off:
                                 // 768 small copies per unit cost the TMA engine more than the LSU loads (247 vs 218 us)

// ---------------------------------------------------------------------------------------------------------
// fp32 token-major value rows -> fp16 head-major planes.  Thread = 8 channels of one (row, head).
__global__ void __launch_bounds__(256) value_to_half_kernel(const float* __restrict__ in, uint4* __restrict__ out, int G,
                                                            int Nv, int H, int Dh) {
  const int cq_n = Dh / 8;
  const int64_t total = (int64_t)G * H * Nv * cq_n;
  const int C = H * Dh;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int cq = (int)(w % cq_n);
    int64_t t = w / cq_n;
    const int row = (int)(t % Nv);
    t /= Nv;
    const int h = (int)(t % H), g = (int)(t / H);
    const float* src = in + ((int64_t)g * Nv + row) * C + h * Dh + cq * 8;
    const float4 a = ld_stream4(src), b = ld_stream4(src + 4);
    const float lim = 65504.f;
    auto sat = [lim](float v) { return fminf(fmaxf(v, -lim), lim); };
    const __half2 h0 = __floats2half2_rn(sat(a.x), sat(a.y)), h1 = __floats2half2_rn(sat(a.z), sat(a.w));
    const __half2 h2 = __floats2half2_rn(sat(b.x), sat(b.y)), h3 = __floats2half2_rn(sat(b.z), sat(b.w));
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h0), o.y = *reinterpret_cast<const uint32_t*>(&h1);
    o.z = *reinterpret_cast<const uint32_t*>(&h2), o.w = *reinterpret_cast<const uint32_t*>(&h3);
    out[w] = o;
  }
}

// ---------------------------------------------------------------------------------------------------------
// shared pieces of the two sampling kernels

// Per-warp descriptor buffer of the warp's 16 items: weight pairs (item stride padded by two words so the four
// items a warp reads in one instruction fall into different banks), then 16-bit window pixel indices.
template <int PP>
struct Desc {
  static constexpr int w_stride = PP * 2 + 2;  // words per item
  static constexpr int w_bytes = kWarpItems * w_stride * 4;
  static constexpr int idx_bytes = kWarpItems * PP * 2;
  static constexpr int bytes = w_bytes + idx_bytes;
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One sample -> descriptor words (branch-free).  (h_im, w_im): pixel coordinates in the value map; aw: attention
// weight (already scaled, 0 for an invalid item); window origin (wy0, wx0) / size (WW, WH) in value-map pixels.
// Returns true when the sample touches the map but its 2 x 2 footprint is not inside the window ("far").
__device__ __forceinline__ bool make_desc(bool ok, float h_im, float w_im, float aw, int fH, int fW, int wy0, int wx0,
                                          int WW, int WH, uint32_t& wl, uint32_t& wr, uint32_t& idx) {
  const bool inmap = ok & (h_im > -1.f) & (w_im > -1.f) & (h_im < (float)fH) & (w_im < (float)fW);
  const int y0 = __float2int_rd(h_im), x0 = __float2int_rd(w_im);   // saturating: wild coordinates are harmless
  const float lh = h_im - (float)y0, lw = w_im - (float)x0;
  const int yy = y0 - wy0, xx = x0 - wx0;
  const bool inwin = ((unsigned)xx < (unsigned)(WW - 1)) & ((unsigned)yy < (unsigned)(WH - 1));
  const bool use = inmap & inwin;
  const float a2 = use ? aw : 0.f;
  const float wb = a2 * lh, wt = a2 - wb;
  const float wbr = wb * lw, wtr = wt * lw;
  wl = pack_h2(wt - wtr, wb - wbr), wr = pack_h2(wtr, wbr);
  idx = use ? (uint32_t)(yy * WW + xx) : 0u;
  return inmap & !inwin;
}

// the lane's PPL descriptors -> the warp's descriptor arrays (weight pairs at item stride Desc::w_stride, indices)
template <int PP>
__device__ __forceinline__ void store_descs(uint32_t sm_w, uint32_t sm_idx, int item, int p0, const uint32_t (&wl)[PP / 2],
                                            const uint32_t (&wr)[PP / 2], const uint32_t (&idx)[PP / 2]) {
  constexpr int PPL = PP / 2;
  const uint32_t wa = sm_w + (uint32_t)(item * Desc<PP>::w_stride + p0 * 2) * 4u;
#pragma unroll
  for (int i = 0; i < PPL; ++i)
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(wa + i * 8), "r"(wl[i]), "r"(wr[i]) : "memory");
  const uint32_t ia = sm_idx + (uint32_t)(item * PP + p0) * 2u;
  if (PPL == 4)
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(ia), "r"(idx[0] | (idx[1] << 16)), "r"(idx[2] | (idx[PPL - 1] << 16))
                 : "memory");
  else
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ia), "r"(idx[0] | (idx[1] << 16)) : "memory");
}

// acc[0..7] += fp16x8 (v) * one fp16 weight (low half of w when TOP, high half otherwise), fp32 accumulation
template <bool TOP>
__device__ __forceinline__ void fhfma8(float (&acc)[8], const uint4& v, uint32_t w) {
  if (TOP) {
    asm("{\n"
        ".reg .b16 a0,a1,a2,a3,a4,a5,a6,a7,wl,wh;\n"
        "mov.b32 {a0,a1}, %8;\n"
        "mov.b32 {a2,a3}, %9;\n"
        "mov.b32 {a4,a5}, %10;\n"
        "mov.b32 {a6,a7}, %11;\n"
        "mov.b32 {wl,wh}, %12;\n"
        "fma.rn.f32.f16 %0, a0, wl, %0;\n"
        "fma.rn.f32.f16 %1, a1, wl, %1;\n"
        "fma.rn.f32.f16 %2, a2, wl, %2;\n"
        "fma.rn.f32.f16 %3, a3, wl, %3;\n"
        "fma.rn.f32.f16 %4, a4, wl, %4;\n"
        "fma.rn.f32.f16 %5, a5, wl, %5;\n"
        "fma.rn.f32.f16 %6, a6, wl, %6;\n"
        "fma.rn.f32.f16 %7, a7, wl, %7;\n"
        "}"
        : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
        : "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(w));
  } else {
    asm("{\n"
        ".reg .b16 a0,a1,a2,a3,a4,a5,a6,a7,wl,wh;\n"
        "mov.b32 {a0,a1}, %8;\n"
        "mov.b32 {a2,a3}, %9;\n"
        "mov.b32 {a4,a5}, %10;\n"
        "mov.b32 {a6,a7}, %11;\n"
        "mov.b32 {wl,wh}, %12;\n"
        "fma.rn.f32.f16 %0, a0, wh, %0;\n"
        "fma.rn.f32.f16 %1, a1, wh, %1;\n"
        "fma.rn.f32.f16 %2, a2, wh, %2;\n"
        "fma.rn.f32.f16 %3, a3, wh, %3;\n"
        "fma.rn.f32.f16 %4, a4, wh, %4;\n"
        "fma.rn.f32.f16 %5, a5, wh, %5;\n"
        "fma.rn.f32.f16 %6, a6, wh, %6;\n"
        "fma.rn.f32.f16 %7, a7, wh, %7;\n"
        "}"
        : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
        : "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(w));
  }
}

// fp16 output rows: this lane's four channels as 8 bytes
__device__ __forceinline__ void st_half4(__half* p, const float4& v) {
  asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(pack_h2(v.x, v.y)), "r"(pack_h2(v.z, v.w)) : "memory");
}
__device__ __forceinline__ void red_add_half4(__half* p, const float4& v) {
  asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1,%2};" ::"l"(p), "r"(pack_h2(v.x, v.y)), "r"(pack_h2(v.z, v.w))
               : "memory");
}

// The gather of one warp's 16 items: lane group `grp` (8 lanes) reduces items grp, grp + 4, grp + 8, grp + 12, two
// at a time.  sm_w / sm_idx: the warp's descriptor arrays (shared-space byte addresses); win: window base
// + sub * 16; ROWB > 0: bytes per window row known at compile time (immediate offset of the bottom-row load).
// emit(item, o) receives this lane's four output channels (channel cq * 8 + half * 4 ...).
template <int PP, int ROWB, typename Emit>
__device__ __forceinline__ void gather_warp(uint32_t sm_w, uint32_t sm_idx, uint32_t win, uint32_t row_rt, int grp,
                                            int half, Emit emit) {
  const uint32_t row_b = ROWB > 0 ? (uint32_t)ROWB : row_rt;
#pragma unroll 1
  for (int k = 0; k < 4; k += 2) {
    const int item[2] = {grp + k * 4, grp + (k + 1) * 4};
    uint32_t ix[2][4], wa[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      wa[j] = sm_w + (uint32_t)(item[j] * Desc<PP>::w_stride + half) * 4u;
      if (PP == 8) {
        const uint4 t = lds128(sm_idx + item[j] * 16);
        ix[j][0] = t.x, ix[j][1] = t.y, ix[j][2] = t.z, ix[j][3] = t.w;
      } else {
        uint32_t a, b;
        asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(sm_idx + item[j] * 8));
        ix[j][0] = a, ix[j][1] = b, ix[j][2] = 0, ix[j][3] = 0;
      }
    }
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      uint4 top[2], bot[2];
      uint32_t w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t word = ix[j][p >> 1];
        const uint32_t id = (p & 1) ? (word >> 16) : (word & 0xffffu);
        const uint32_t a = win + id * 64u;
        w[j] = lds32(wa[j] + p * 8);
        top[j] = lds128(a);
        bot[j] = lds128(a + row_b);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        fhfma8<true>(acc[j], top[j], w[j]);
        fhfma8<false>(acc[j], bot[j], w[j]);
      }
    }
    // combine the left / right pixel halves: after the exchange half 0 owns channels +0..3, half 1 channels +4..7
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float s0 = half ? acc[j][0] : acc[j][4], s1 = half ? acc[j][1] : acc[j][5];
      const float s2 = half ? acc[j][2] : acc[j][6], s3 = half ? acc[j][3] : acc[j][7];
      const float r0 = __shfl_xor_sync(0xffffffffu, s0, 4), r1 = __shfl_xor_sync(0xffffffffu, s1, 4);
      const float r2 = __shfl_xor_sync(0xffffffffu, s2, 4), r3 = __shfl_xor_sync(0xffffffffu, s3, 4);
      float4 o;
      o.x = (half ? acc[j][4] : acc[j][0]) + r0;
      o.y = (half ? acc[j][5] : acc[j][1]) + r1;
      o.z = (half ? acc[j][6] : acc[j][2]) + r2;
      o.w = (half ? acc[j][7] : acc[j][3]) + r3;
      emit(item[j], o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct BevWinArgs {
  const __half* value16;  // (B*H, fH, fW, 32): far path
  float* out;             // (B, Nq, H*32) fp32 rows, or null when out16 is set
  __half* out16;          // (B, Nq, H*32) fp16 rows (the A operand of the fp16 output projection), or null
  int* counters;          // [0] next unit, [1] CTAs done
  int B, bev_h, bev_w, fH, fW, H;
  int tiles_x, tiles_y, n_units;
  int WW, WH, R;
  int off_col, logit_col;
  float sx, sy;
  int round_tf32;
};

template <int PP>
struct BevSmem {
  static constexpr int slice_off_bytes = kWarpItems * PP * 8, slice_lg_bytes = kWarpItems * PP * 4;
  static constexpr int slice_bytes = slice_off_bytes + slice_lg_bytes;      // one query row of the offset|logit tile
  static constexpr int warp_bytes = slice_bytes + ((Desc<PP>::bytes + 127) & ~127);
  static size_t total(int win_bytes) { return (size_t)2 * win_bytes + (size_t)kWorkerWarps * warp_bytes; }
};

// Worker warp w owns query row ty0 + w of the unit's 16 x 16 tile.  Its loop per unit k:
//   P1  descriptors of its 128 / 64 samples from its TMA-staged slice of the offset|logit rows (own mbarrier)
//   --  issue the slice of unit k + 1 (the buffer is free), slow path for far samples
//   P2  wait for window k (full[k & 1]), gather its 16 items, arrive on empty[k & 1]
// No CTA-wide barrier: the warps drift apart, so some build descriptors (ALU / SFU) while others gather (LDS).
// The scheduler warp hands out units (atomic counter), publishes them two units ahead through a 4-slot ring and
// streams window k into buffer k & 1 as soon as every worker has released it (unit k - 2).
template <int PP, int ROWB>
__global__ void __launch_bounds__(kBevThreads, 1)
    bev_sample_win_kernel(const BevWinArgs a, const __grid_constant__ CUtensorMap map_val,
                          const __grid_constant__ CUtensorMap map_off, const __grid_constant__ CUtensorMap map_lg) {
  using S = BevSmem<PP>;
  using D = Desc<PP>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], s_unit[4], s_qp[kWorkerWarps];
  __shared__ UnitInfo s_ring[4];

  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  const uint32_t sm_win = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Nq = a.bev_h * a.bev_w, C = a.H * 32;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&s_full[i]), 1), mbar_init(smem_u32(&s_empty[i]), kWorkerWarps);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s_unit[i]), 1);
    for (int i = 0; i < kWorkerWarps; ++i) mbar_init(smem_u32(&s_qp[i]), 1);
    mbar_init_fence();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();   // value planes / offset|logit rows come from the predecessor kernels

  if (warp == kWorkerWarps) {
    // ---------------- scheduler warp (one lane)
    if (lane != 0) return;
    tma_prefetch_desc(&map_val);
    const int n_tiles = a.tiles_x * a.tiles_y;
    for (int k = 0;; ++k) {
      const int u = atomicAdd(&a.counters[0], 1);
      UnitInfo w;
      w.pad = 0;
      if (u >= a.n_units) {
        w.u = -1, w.b = w.h = w.tx0 = w.ty0 = w.wx0 = w.wy0 = 0;
      } else {
        w.u = u;
        w.h = u % a.H;
        const int t = (u / a.H) % n_tiles;
        w.b = u / (a.H * n_tiles);
        w.tx0 = (t % a.tiles_x) * kTQ, w.ty0 = (t / a.tiles_x) * kTQ;
        w.wx0 = (int)floorf(((float)w.tx0 + 0.5f) * a.sx - 0.5f) - a.R;
        w.wy0 = (int)floorf(((float)w.ty0 + 0.5f) * a.sy - 0.5f) - a.R;
      }
      // ring slot k & 3 held unit k - 4, which every worker released before window k - 2 was issued
      s_ring[k & 3] = w;
      mbar_arrive(smem_u32(&s_unit[k & 3]));
      if (w.u < 0) break;
      if (k >= 2) mbar_wait(smem_u32(&s_empty[k & 1]), (uint32_t)(((k >> 1) - 1) & 1));
      const uint32_t bar = smem_u32(&s_full[k & 1]);
      mbar_arrive_expect_tx(bar, (uint32_t)(a.WW * a.WH * 64));
      tma_load_4d(sm_win + (uint32_t)(k & 1) * win_bytes, &map_val, bar, 0, w.wx0, w.wy0, w.b * a.H + w.h);
    }
    // the last CTA to leave re-arms the unit counter for the next launch
    __threadfence();
    const int done = atomicAdd(&a.counters[1], 1);
    if (done == (int)gridDim.x - 1) {
      a.counters[0] = 0;
      a.counters[1] = 0;
      __threadfence();
    }
    return;
  }

  // ---------------- worker warps
  const uint32_t sm_slice = sm_win + 2u * win_bytes + (uint32_t)warp * S::warp_bytes;  // {offsets, logits}
  const uint32_t sm_w = sm_slice + S::slice_bytes, sm_idx = sm_w + D::w_bytes;
  const uint32_t bar_qp = smem_u32(&s_qp[warp]);
  const uint32_t bar_unit = smem_u32(&s_unit[0]), bar_full = smem_u32(&s_full[0]), bar_empty = smem_u32(&s_empty[0]);
  constexpr int PPL = PP / 2;                // sampling points per lane in P1 (two lanes per item)
  const int item_l = lane >> 1, p0 = (lane & 1) * PPL;
  const int grp = lane >> 3, sub = lane & 7, half = sub >> 2, cq = sub & 3;

  auto issue_slice = [&](const UnitInfo& w) {  // lane 0: this warp's query row of the offset|logit tile
    mbar_arrive_expect_tx(bar_qp, (uint32_t)S::slice_bytes);
    tma_load_4d(sm_slice, &map_off, bar_qp, a.off_col + w.h * PP * 2, w.tx0, w.ty0 + warp, w.b);
    tma_load_4d(sm_slice + S::slice_off_bytes, &map_lg, bar_qp, a.logit_col + w.h * PP, w.tx0, w.ty0 + warp, w.b);
  };

  mbar_wait(bar_unit, 0u);
  UnitInfo w = s_ring[0];
  if (w.u >= 0 && lane == 0) issue_slice(w);

  for (int k = 0; w.u >= 0; ++k) {
    const int qy = w.ty0 + warp;
    const bool row_ok = qy < a.bev_h;
    const float hbase = (float)qy + 0.5f;
    // ---- P1
    mbar_wait(bar_qp, (uint32_t)(k & 1));
    const int qx1 = w.tx0 + item_l;
    const bool ok = row_ok & (qx1 < a.bev_w);
    float off[PPL * 2], lg[PPL], far_h[PPL], far_w[PPL], far_a[PPL];
    {
      const uint32_t oa = sm_slice + (uint32_t)(item_l * PP + p0) * 8u;
      const uint32_t la = sm_slice + S::slice_off_bytes + (uint32_t)(item_l * PP + p0) * 4u;
#pragma unroll
      for (int i = 0; i < PPL / 2; ++i)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(off[4 * i]), "=f"(off[4 * i + 1]), "=f"(off[4 * i + 2]), "=f"(off[4 * i + 3])
                     : "r"(oa + i * 16));
      if (PPL == 4)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(lg[0]), "=f"(lg[1]), "=f"(lg[2]), "=f"(lg[PPL - 1]) : "r"(la));
      else
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(lg[0]), "=f"(lg[1]) : "r"(la));
    }
    softmax_pair<PPL>(lg, ok, 1.f, far_a);
    unsigned far_bits = 0;
    {
      // pixel = ((q + .5) / bev + off / f) * f - .5  ==  (q + .5) * (f / bev) + off - .5
      const float wbase = (float)qx1 + 0.5f;
      uint32_t wl[PPL], wr[PPL], idx[PPL];
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        far_w[i] = fmaf(wbase, a.sx, off[2 * i] - 0.5f);
        far_h[i] = fmaf(hbase, a.sy, off[2 * i + 1] - 0.5f);
        if (make_desc(ok, far_h[i], far_w[i], far_a[i], a.fH, a.fW, w.wy0, w.wx0, a.WW, a.WH, wl[i], wr[i], idx[i]))
          far_bits |= 1u << i;
      }
      store_descs<PP>(sm_w, sm_idx, item_l, p0, wl, wr, idx);
    }
    const bool any_far = __any_sync(0xffffffffu, far_bits != 0u);
    __syncwarp();   // descriptor stores visible to the whole warp
    // ---- the slice buffer is free: stream in the next unit's row
    mbar_wait(bar_unit + 8u * (uint32_t)((k + 1) & 3), (uint32_t)(((k + 1) >> 2) & 1));
    const UnitInfo wn = s_ring[(k + 1) & 3];
    if (wn.u >= 0 && lane == 0) issue_slice(wn);

    // ---- slow path for samples outside the staged window (exact: fp32 weights, per-corner bounds checks)
    if (any_far) {
      if (row_ok) {
#pragma unroll
        for (int e = lane; e < kWarpItems * 8; e += 32) {
          const int item = e >> 3, c4 = e & 7;
          if (w.tx0 + item < a.bev_w) {
            const int64_t o = ((int64_t)w.b * Nq + qy * a.bev_w + w.tx0 + item) * C + w.h * 32 + c4 * 4;
            if (a.out16)
              *reinterpret_cast<uint2*>(a.out16 + o) = make_uint2(0u, 0u);
            else
              *reinterpret_cast<float4*>(a.out + o) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      __syncwarp();
      const __half* plane = a.value16 + (int64_t)(w.b * a.H + w.h) * a.fH * a.fW * 32;
#pragma unroll
      for (int r = 0; r < PPL; ++r) {
        unsigned m = __ballot_sync(0xffffffffu, (far_bits >> r) & 1u);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float h_im = __shfl_sync(0xffffffffu, far_h[r], src);
          const float w_im = __shfl_sync(0xffffffffu, far_w[r], src);
          const float aw = __shfl_sync(0xffffffffu, far_a[r], src);
          const int item = src >> 1;
          const int corner = lane >> 3, c4 = lane & 7, dy = corner >> 1, dx = corner & 1;
          const float hf = floorf(h_im), wf = floorf(w_im);
          const float lh = h_im - hf, lw = w_im - wf;
          const int y = (int)hf + dy, x = (int)wf + dx;
          const float wgt = aw * ((dy ? lh : 1.f - lh) * (dx ? lw : 1.f - lw));
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (y >= 0 && y < a.fH && x >= 0 && x < a.fW) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(plane + ((int64_t)y * a.fW + x) * 32 + c4 * 4));
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            v = make_float4(wgt * f0.x, wgt * f0.y, wgt * f1.x, wgt * f1.y);
          }
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o), v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, o), v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
          }
          if (lane < 8) {
            const int64_t o = ((int64_t)w.b * Nq + qy * a.bev_w + w.tx0 + item) * C + w.h * 32 + c4 * 4;
            if (a.out16)
              red_add_half4(a.out16 + o, v);
            else
              red_add4(a.out + o, v);
          }
        }
      }
      __syncwarp();
    }

    // ---- P2
    // element offset of this lane's four channels in the first query of the warp's row
    const int64_t out0 = ((int64_t)w.b * Nq + qy * a.bev_w + w.tx0) * C + w.h * 32 + cq * 8 + half * 4;
    mbar_wait(bar_full + 8u * (uint32_t)(k & 1), (uint32_t)((k >> 1) & 1));
    gather_warp<PP, ROWB>(sm_w, sm_idx, sm_win + (uint32_t)(k & 1) * win_bytes + sub * 16u, (uint32_t)a.WW * 64u, grp,
                          half, [&](int item, const float4& o) {
                            const int qx = w.tx0 + item;
                            if (row_ok && qx < a.bev_w) {
                              const int64_t oo = out0 + item * C;
                              if (a.out16) {
                                if (any_far)
                                  red_add_half4(a.out16 + oo, o);
                                else
                                  st_half4(a.out16 + oo, o);
                                return;
                              }
                              float* dst = a.out + oo;
                              if (any_far) {
                                red_add4(dst, o);
                              } else if (!a.round_tf32) {
                                st_stream4(dst, o);
                              } else {
                                // the consumer is the output projection's TF32 tensor-core GEMM, which truncates
                                // its operands: round to nearest here instead
                                float4 r;
                                r.x = round_tf32(o.x), r.y = round_tf32(o.y), r.z = round_tf32(o.z), r.w = round_tf32(o.w);
                                st_stream4(dst, r);
                              }
                            }
                          });
    __syncwarp();   // every lane is done with the window and the descriptors
    if (lane == 0) mbar_arrive(bar_empty + 8u * (uint32_t)(k & 1));
    w = wn;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Per-camera hit lists from batch item 0's visibility bits (reference quirk, spatial_cross_attention_img.py:142),
// split by RANK: a hit (n, q) is "first" when no camera n' < n sees q, "later" otherwise.  Row n of hit_idx holds
// the first hits ascending from the front and the later hits from the back (hit_idx[n][Nq - 1 - k]); row N lists the
// queries no camera sees.  hit_cnt = {first counts (N), later counts (N), unseen count}.  The sampling kernel
// writes first hits with plain stores (and zero rows for the unseen queries), then accumulates the later ones:
// no zero-fill pass, no read-modify-write for the ~88 % of pairs that are a query's only camera.
// inv_cnt = 1 / max(1, #cameras that see (b, q)) (:209-212).
__global__ void __launch_bounds__(1024) build_hits_kernel(const uint8_t* __restrict__ mask, int* __restrict__ hit_idx,
                                                          int* __restrict__ hit_cnt, float* __restrict__ inv_cnt, int B,
                                                          int N, int Nq) {
  __shared__ int s_warp[2][32];
  __shared__ int s_base[2];
  pdl_trigger();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((int)blockIdx.x <= N) {
    const int n = blockIdx.x;          // n == N: the unseen list
    if (tid < 2) s_base[tid] = 0;
    __syncthreads();
    for (int q0 = 0; q0 < Nq; q0 += 1024) {
      const int q = q0 + tid;
      bool below = false, own = false;
      if (q < Nq) {
        for (int m = 0; m < n; ++m) below |= mask[(int64_t)q * N + m] != 0;
        own = n < N ? mask[(int64_t)q * N + n] != 0 : !below;
      }
      const bool k0 = own && !below;                 // first hit of q (or: unseen, when n == N)
      const bool k1 = own && below && n < N;         // later hit
      const unsigned b0 = __ballot_sync(0xffffffffu, k0), b1 = __ballot_sync(0xffffffffu, k1);
      if (lane == 0) s_warp[0][warp] = __popc(b0), s_warp[1][warp] = __popc(b1);
      __syncthreads();
      if (warp < 2) {
        int v = s_warp[warp][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += t;
        }
        s_warp[warp][lane] = v;  // inclusive
      }
      __syncthreads();
      const unsigned below_me = (1u << lane) - 1u;
      if (k0) hit_idx[(int64_t)n * Nq + s_base[0] + (warp ? s_warp[0][warp - 1] : 0) + __popc(b0 & below_me)] = q;
      if (k1) hit_idx[(int64_t)n * Nq + Nq - 1 - (s_base[1] + (warp ? s_warp[1][warp - 1] : 0) + __popc(b1 & below_me))] = q;
      __syncthreads();
      if (tid < 2) s_base[tid] += s_warp[tid][31];
      __syncthreads();
    }
    if (tid == 0) {
      if (n < N)
        hit_cnt[n] = s_base[0], hit_cnt[N + n] = s_base[1];
      else
        hit_cnt[2 * N] = s_base[0];
    }
  } else {
    const int64_t nb = gridDim.x - (N + 1);
    for (int64_t i = (int64_t)(blockIdx.x - (N + 1)) * blockDim.x + tid; i < (int64_t)B * Nq; i += nb * blockDim.x) {
      int c = 0;
      for (int n = 0; n < N; ++n) c += mask[i * N + n] != 0 ? 1 : 0;
      inv_cnt[i] = 1.f / (float)max(c, 1);
    }
  }
}

// 1 / count in hit-list order: hit_ic[b][n][pos] = inv_cnt[b][hit_idx[n][pos]] for the valid positions of row n, so the
// sampling kernel reads it with the same coalesced access as the hit list (instead of one scattered load per hit).
__global__ void __launch_bounds__(256) hit_ic_kernel(const int* __restrict__ hit_idx, const int* __restrict__ hit_cnt,
                                                     const float* __restrict__ inv_cnt, float* __restrict__ hit_ic, int B,
                                                     int N, int Nq) {
  pdl_trigger();
  const int64_t total = (int64_t)B * N * Nq;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pos = (int)(i % Nq), n = (int)((i / Nq) % N), b = (int)(i / ((int64_t)Nq * N));
    if (pos < hit_cnt[n] || pos >= Nq - hit_cnt[N + n])
      hit_ic[i] = inv_cnt[(int64_t)b * Nq + hit_idx[(int64_t)n * Nq + pos]];
  }
}

// ---------------------------------------------------------------------------------------------------------
struct ImgWinArgs {
  const float* qproj;
  const float* ref_cam;   // (B, Nq, N, D, 2)
  const float* inv_cnt;   // (B, Nq)
  const float* hit_ic;    // (B, N, Nq) 1 / count in hit-list order, or null
  const int* hit_idx;     // (N + 1, Nq): first hits from the front, later hits from the back; row N = unseen queries
  const int* hit_cnt;     // (2 N + 1): first counts, later counts, unseen count
  float* out;             // (B, Nq, H*32) fp32 rows, or null when out16 is set
  __half* out16;          // (B, Nq, H*32) fp16 rows, or null
  int B, N, Nq, fH, fW, H, P, D, ld, off_col, logit_col;
  int WW, WH;
  int part;               // 0: first hits (plain stores) + zero rows of the unseen queries; 1: later hits (red.add)
  int two_win;            // two window buffers fit: the next camera plane streams in behind the current one
  int vec_ref;            // anchors of a lane are consecutive and may be read with 16-byte loads
  int vec8;               // P = 8, D = 4 and 32-byte aligned rows: a lane's offsets / anchors are one 256-bit load each
};

template <int PP>
struct ImgSmem {
  static constexpr int desc_bytes = (Desc<PP>::bytes + kWarpItems * 4 + 127) & ~127;
  // staged P1 inputs of the warp's 16 items: offsets (PP x 8 B), logits (PP x 4 B), projected anchors (up to 8 x 8 B)
  static constexpr int slice_off = kWarpItems * PP * 8, slice_lg = kWarpItems * PP * 4, slice_ref = kWarpItems * 64;
  static constexpr int slice_bytes = slice_off + slice_lg + slice_ref;
  __host__ __device__ static constexpr int warp_bytes(bool stage) { return desc_bytes + (stage ? slice_bytes : 0); }
  static size_t total(int win_bytes, int n_win, bool stage) {
    return (size_t)n_win * win_bytes + (size_t)kWorkerWarps * warp_bytes(stage);
  }
};

// global -> shared bulk copy (TMA engine, no tensor map); completes `bar` with `bytes`.  16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// Camera mode.  Units (b, camera, chunk of 256 hits) x head, a contiguous range of chunks per group of H CTAs (one
// head each, see below); warp w owns hits chunk * 256 + 16 w .. + 15.  Per unit and warp: descriptors from registers prefetched during the previous gather
// (hit index -> offsets / logits / projected anchor / 1/count straight from global memory), then the gather.
// The only CTA-wide barrier is at a plane change (the window is reloaded once every warp has left the old plane).
// Launched twice per call: part 0 stores the contribution of every query's FIRST camera (and zero rows for the
// queries no camera sees), part 1 accumulates the remaining (camera, query) pairs on top.
// STAGE: the hit-ordered P1 inputs (offset / logit pieces of the query's GEMM row, projected anchors) are brought into
// a per-warp shared-memory slice by bulk async copies one unit ahead (TMA engine: they cost no LSU / L1 wavefronts and
// no registers); otherwise they are prefetched into registers with ordinary loads.
template <int PP, int ROWB, bool STAGE>
__global__ void __launch_bounds__(kImgThreads, 1)
    img_sample_win_kernel(const ImgWinArgs a, const __grid_constant__ CUtensorMap map_val) {
  using D = Desc<PP>;
  using SM = ImgSmem<PP>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar[2], s_qp[kWorkerWarps];

  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sm_win = smem_u32(smem), bar = smem_u32(&s_bar[0]);
  const uint32_t sm_w = sm_win + (uint32_t)(a.two_win ? 2 : 1) * (uint32_t)win_bytes + (uint32_t)warp * SM::warp_bytes(STAGE);
  const uint32_t sm_idx = sm_w + D::w_bytes;
  const uint32_t sl_off = sm_w + SM::desc_bytes, sl_lg = sl_off + SM::slice_off, sl_ref = sl_lg + SM::slice_lg;
  const uint32_t bar_qp = smem_u32(&s_qp[warp]);
  const uint32_t ref_bytes = (uint32_t)a.D * 8u;      // per item (a multiple of 16 in STAGE mode)
  const uint32_t sm_q = sm_idx + D::idx_bytes;   // query index per item
  const int C = a.H * 32;

  pdl_trigger();
  pdl_wait();   // hit lists, planes, offset|logit rows come from the predecessor kernels; `out` may alias their inputs
  const int* cnt_p = a.hit_cnt + a.part * a.N;
  if (a.part == 0) {   // rows of the queries no camera sees
    const int n_zero = __ldg(a.hit_cnt + 2 * a.N);
    const int per_row = C / 4;
    for (int64_t e = (int64_t)blockIdx.x * kImgThreads + tid; e < (int64_t)a.B * n_zero * per_row;
         e += (int64_t)gridDim.x * kImgThreads) {
      const int c4 = (int)(e % per_row);
      const int64_t r = e / per_row;
      const int q = __ldg(a.hit_idx + (int64_t)a.N * a.Nq + (int)(r % n_zero));
      const int64_t o = ((r / n_zero) * a.Nq + q) * C + c4 * 4;
      if (a.out16)
        *reinterpret_cast<uint2*>(a.out16 + o) = make_uint2(0u, 0u);
      else
        st_stream4(a.out + o, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  // Units = (b, camera, chunk of 256 hits) x this CTA's head.  CTAs are grouped by H: the H CTAs of a group walk the
  // same contiguous range of chunks, one head each, at about the same time -- so a query's offset|logit row, its
  // projected anchors and its hit-list entry come from DRAM once and from L2 for the other heads -- and a CTA
  // changes its window only when the range crosses into another camera.
  int chunks_tot = 0;
  for (int n = 0; n < a.N; ++n) chunks_tot += (__ldg(cnt_p + n) + kUnitItems - 1) / kUnitItems;
  const int n_grp = (int)gridDim.x / a.H, cgrp = (int)blockIdx.x / a.H, my_h = (int)blockIdx.x % a.H;
  if (cgrp >= n_grp) return;
  const int total = chunks_tot * a.B;
  const int per = total / n_grp, rem = total % n_grp;
  const int u_beg = cgrp * per + min(cgrp, rem);
  const int u_end = u_beg + per + (cgrp < rem ? 1 : 0);

  struct Unit {
    int b, n, h, chunk, cnt, plane;
  };
  auto decode = [&](int uu) {
    Unit w;
    w.b = uu / chunks_tot;
    int r = uu % chunks_tot;
    w.n = 0, w.cnt = 0, w.h = my_h, w.chunk = 0;
    for (int n = 0; n < a.N; ++n) {
      const int cnt = __ldg(cnt_p + n), ch = (cnt + kUnitItems - 1) / kUnitItems;
      if (r < ch) {
        w.n = n, w.cnt = cnt, w.chunk = r;
        break;
      }
      r -= ch;
    }
    w.plane = (w.b * a.N + w.n) * a.H + w.h;
    return w;
  };

  if (u_beg >= u_end) return;   // uniform per CTA
  Unit w = decode(u_beg);
  // first plane other than `plane` among the units v >= u_from of this CTA's range (-1: none)
  auto next_plane = [&](int u_from, int plane) {
    for (int v = u_from; v < u_end; ++v) {
      const int p = decode(v).plane;
      if (p != plane) return p;
    }
    return -1;
  };
  auto load_plane = [&](int buf, int plane) {   // one thread
    mbar_arrive_expect_tx(bar + 8u * buf, (uint32_t)(a.WW * a.WH * 64));
    tma_load_4d(sm_win + (uint32_t)buf * win_bytes, &map_val, bar + 8u * buf, 0, -1, -1, plane);
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8u, 1);
    for (int i = 0; i < kWorkerWarps; ++i) mbar_init(smem_u32(&s_qp[i]), 1);
    mbar_init_fence();
    tma_prefetch_desc(&map_val);
    load_plane(0, w.plane);
    if (a.two_win) {
      const int np = next_plane(u_beg + 1, w.plane);
      if (np >= 0) load_plane(1, np);
    }
  }
  __syncthreads();

  constexpr int PPL = PP / 2;                // sampling points per lane in P1 (two lanes per item)
  const int item_l = lane >> 1, p0 = (lane & 1) * PPL;
  const int grp = lane >> 3, sub = lane & 7, half = sub >> 2, cq = sub & 3;
  float off[PPL * 2], lg[PPL], ref[PPL * 2], ic;
  int qq;
  auto lookup = [&](const Unit& wu) {   // the query of this lane pair's item in unit wu, or -1
    const int ord = wu.chunk * kUnitItems + warp * kWarpItems + item_l;
    return ord < wu.cnt ? __ldg(a.hit_idx + (int64_t)wu.n * a.Nq + (a.part ? a.Nq - 1 - ord : ord)) : -1;
  };
  auto inv_count = [&](const Unit& wu, int q) {   // 1 / #cameras of the item: list-ordered copy when the caller has one
    if (a.hit_ic) {
      const int ord = wu.chunk * kUnitItems + warp * kWarpItems + item_l;
      return ord < wu.cnt ? __ldg(a.hit_ic + ((int64_t)wu.b * a.N + wu.n) * a.Nq + (a.part ? a.Nq - 1 - ord : ord)) : 0.f;
    }
    return q >= 0 ? __ldg(a.inv_cnt + (int64_t)wu.b * a.Nq + q) : 0.f;
  };
  // !STAGE: everything P1 needs into registers.  `q` (the item's query in unit wu) was itself loaded one unit earlier,
  // so the row loads below go out at once instead of waiting for the hit-list entry
  auto prefetch = [&](const Unit& wu, int q) {
    qq = q;
    ic = inv_count(wu, qq);
#pragma unroll
    for (int i = 0; i < PPL; ++i) off[2 * i] = 0.f, off[2 * i + 1] = 0.f, lg[i] = 0.f, ref[2 * i] = 0.f, ref[2 * i + 1] = 0.f;
    if (qq >= 0) {
      const int64_t bq = (int64_t)wu.b * a.Nq + qq;
      const float* rowp = a.qproj + bq * a.ld;
      const float4* op = reinterpret_cast<const float4*>(rowp + a.off_col + (wu.h * PP + p0) * 2);
      if (PPL == 4 && a.vec8) {
        float t[8];
        ld_stream8(reinterpret_cast<const float*>(op), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) off[i % (PPL * 2)] = t[i];
      } else {
#pragma unroll
        for (int i = 0; i < PPL / 2; ++i) {
          const float4 t = ld_stream4(reinterpret_cast<const float*>(op + i));
          off[4 * i] = t.x, off[4 * i + 1] = t.y, off[4 * i + 2] = t.z, off[4 * i + 3] = t.w;
        }
      }
      const float* lp = rowp + a.logit_col + wu.h * PP + p0;
      if (PPL == 4) {
        const float4 t = ld_stream4(lp);
        lg[0] = t.x, lg[1] = t.y, lg[2] = t.z, lg[PPL - 1] = t.w;
      } else {
        const float2 t = ld_stream2(lp);
        lg[0] = t.x, lg[1] = t.y;
      }
      const float2* rp = reinterpret_cast<const float2*>(a.ref_cam) + (bq * a.N + wu.n) * a.D;
      if (PPL == 4 && a.vec8) {
        float t[8];
        ld_stream8(reinterpret_cast<const float*>(rp), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) ref[i % (PPL * 2)] = t[i];
      } else if (a.vec_ref) {   // the lane's PPL anchors are consecutive: 16-byte loads
        const float4* r4 = reinterpret_cast<const float4*>(rp + p0 % a.D);
#pragma unroll
        for (int i = 0; i < PPL / 2; ++i) {
          const float4 t = __ldg(r4 + i);
          ref[4 * i] = t.x, ref[4 * i + 1] = t.y, ref[4 * i + 2] = t.z, ref[4 * i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          const float2 t = __ldg(rp + (p0 + i) % a.D);
          ref[2 * i] = t.x, ref[2 * i + 1] = t.y;
        }
      }
    }
  };
  // STAGE: bulk copies of unit wu's P1 inputs into the warp's slice (even lane of a pair: offsets, odd lane: logits and
  // anchors); lane 0 posts the byte count of the warp's valid items
  auto issue = [&](const Unit& wu, int q) {
    const int valid = min(kWarpItems, max(0, wu.cnt - (wu.chunk * kUnitItems + warp * kWarpItems)));
    if (lane == 0) mbar_arrive_expect_tx(bar_qp, (uint32_t)valid * ((uint32_t)PP * 12u + ref_bytes));
    if (q >= 0) {
      const int64_t bq = (int64_t)wu.b * a.Nq + q;
      const float* rowp = a.qproj + bq * a.ld;
      if ((lane & 1) == 0) {
        bulk_g2s(sl_off + (uint32_t)item_l * (PP * 8), rowp + a.off_col + wu.h * PP * 2, PP * 8, bar_qp);
      } else {
        bulk_g2s(sl_lg + (uint32_t)item_l * (PP * 4), rowp + a.logit_col + wu.h * PP, PP * 4, bar_qp);
        bulk_g2s(sl_ref + (uint32_t)item_l * 64u, a.ref_cam + ((bq * a.N + wu.n) * a.D) * 2, ref_bytes, bar_qp);
      }
    }
  };
  int qn = -1;        // STAGE: query of the item in the NEXT unit
  float icn = 0.f;    //        and its 1 / count
  if (STAGE) {
    qq = lookup(w);
    issue(w, qq);
    ic = inv_count(w, qq);
    if (u_beg + 1 < u_end) qn = lookup(decode(u_beg + 1));
  } else {
    prefetch(w, lookup(w));
    if (u_beg + 1 < u_end) qn = lookup(decode(u_beg + 1));
  }

  int loaded = w.plane, cur = 0;
  uint32_t phbits = 0u;   // bit k: phase of window barrier k
  bool fresh = true;
  for (int u = u_beg; u < u_end; ++u) {
    // ---- P1 from the staged slice / the prefetched registers
    {
      const bool ok = qq >= 0;
      if (STAGE) {
        mbar_wait(bar_qp, (uint32_t)((u - u_beg) & 1));
        const uint32_t oa = sl_off + (uint32_t)(item_l * PP + p0) * 8u, la = sl_lg + (uint32_t)(item_l * PP + p0) * 4u;
#pragma unroll
        for (int i = 0; i < PPL / 2; ++i)
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(off[4 * i]), "=f"(off[4 * i + 1]), "=f"(off[4 * i + 2]), "=f"(off[4 * i + 3])
                       : "r"(oa + i * 16));
        if (PPL == 4)
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(lg[0]), "=f"(lg[1]), "=f"(lg[2]), "=f"(lg[PPL - 1]) : "r"(la));
        else
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(lg[0]), "=f"(lg[1]) : "r"(la));
        const uint32_t ra = sl_ref + (uint32_t)item_l * 64u;
        if (a.vec_ref) {
#pragma unroll
          for (int i = 0; i < PPL / 2; ++i)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(ref[4 * i]), "=f"(ref[4 * i + 1]), "=f"(ref[4 * i + 2]), "=f"(ref[4 * i + 3])
                         : "r"(ra + (uint32_t)(p0 % a.D) * 8u + i * 16));
        } else {
#pragma unroll
          for (int i = 0; i < PPL; ++i)
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(ref[2 * i]), "=f"(ref[2 * i + 1])
                         : "r"(ra + (uint32_t)((p0 + i) % a.D) * 8u));
        }
      }
      float aw[PPL];
      softmax_pair<PPL>(lg, ok, ic, aw);
      uint32_t wl[PPL], wr[PPL], idx[PPL];
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        const float w_im = fmaf(ref[2 * i], (float)a.fW, off[2 * i] - 0.5f);
        const float h_im = fmaf(ref[2 * i + 1], (float)a.fH, off[2 * i + 1] - 0.5f);
        // the window is the whole plane plus a one-pixel zero halo (origin (-1, -1)): nothing is ever far
        make_desc(ok, h_im, w_im, aw[i], a.fH, a.fW, -1, -1, a.WW, a.WH, wl[i], wr[i], idx[i]);
      }
      store_descs<PP>(sm_w, sm_idx, item_l, p0, wl, wr, idx);
      if (p0 == 0) asm volatile("st.shared.b32 [%0], %1;" ::"r"(sm_q + (uint32_t)item_l * 4u), "r"(qq) : "memory");
    }
    __syncwarp();
    const Unit w_cur = w;
    if (u + 1 < u_end) {
      w = decode(u + 1);
      if (STAGE) {
        // the slice has been consumed: stream in the next unit's inputs, then look one unit further ahead
        issue(w, qn);
        icn = inv_count(w, qn);
        const int q2 = u + 2 < u_end ? lookup(decode(u + 2)) : -1;
        qq = qn, ic = icn, qn = q2;
      } else {
        prefetch(w, qn);
        qn = u + 2 < u_end ? lookup(decode(u + 2)) : -1;
      }
    }
    // ---- P2
    if (fresh) {
      mbar_wait(bar + 8u * cur, (phbits >> cur) & 1u);
      phbits ^= 1u << cur;
      fresh = false;
    }
    gather_warp<PP, ROWB>(sm_w, sm_idx, sm_win + (uint32_t)cur * win_bytes + sub * 16u, (uint32_t)a.WW * 64u, grp, half,
                          [&](int item, const float4& o) {
                            int q;
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(q) : "r"(sm_q + (uint32_t)item * 4u));
                            if (q >= 0) {
                              const int64_t oo = ((int64_t)w_cur.b * a.Nq + q) * C + w_cur.h * 32 + cq * 8 + half * 4;
                              if (a.out16) {
                                if (a.part)
                                  red_add_half4(a.out16 + oo, o);
                                else
                                  st_half4(a.out16 + oo, o);
                              } else if (a.part) {
                                red_add4(a.out + oo, o);
                              } else {
                                st_stream4(a.out + oo, o);
                              }
                            }
                          });
    __syncwarp();
    if (u + 1 < u_end && w.plane != loaded) {   // uniform per CTA: every warp leaves the old plane first
      __syncthreads();
      if (a.two_win) {
        // the next plane is already (being) loaded into the other buffer; the one just left takes the plane after it
        if (tid == 0) {
          const int np = next_plane(u + 2, w.plane);
          if (np >= 0) load_plane(cur, np);
        }
        cur ^= 1;
      } else if (tid == 0) {
        load_plane(0, w.plane);
      }
      loaded = w.plane;
      fresh = true;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side

template <int PP, int ROWB>
static int launch_bev_win_v(BevWinArgs& a, const CUtensorMap& mv, const CUtensorMap& mo, const CUtensorMap& ml,
                            size_t smem, cudaStream_t s) {
  const char* fn = "ub_bev_sample_win_fwd";
  if (int rc = ensure_smem(bev_sample_win_kernel<PP, ROWB>, smem, fn)) return rc;
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  launch_pdl(bev_sample_win_kernel<PP, ROWB>, dim3(grid), dim3(kBevThreads), smem, s, a, mv, mo, ml);
  return check_launch(fn);
}

template <int PP>
static int launch_bev_win(BevWinArgs& a, const void* value16, const float* qproj, int ld, cudaStream_t s) {
  const char* fn = "ub_bev_sample_win_fwd";
  // preferred row pitch with a compile-time kernel variant; otherwise shrink the halo until two windows fit
  // (samples beyond it stay exact through the slow path)
  constexpr int kPrefWW = PP == 8 ? 36 : 28;
  auto smem_for = [&]() { return BevSmem<PP>::total((a.WW * a.WH * 64 + 127) & ~127); };
  bool fixed = false;
  if (a.WW <= kPrefWW) {
    const int keep = a.WW;
    a.WW = kPrefWW;
    if (smem_for() <= kSmemBudget)
      fixed = true;
    else
      a.WW = keep;
  }
  while (smem_for() > kSmemBudget && a.R > 1) --a.R, a.WW -= 2, a.WH -= 2;
  const size_t smem = smem_for();
  if (smem > kSmemBudget) {
    set_error("%s: window %d x %d needs %zu bytes of shared memory", fn, a.WW, a.WH, smem);
    return ub::unsupported();
  }
  CUtensorMap mv, mo, ml;
  {
    const uint64_t dims[4] = {32, (uint64_t)a.fW, (uint64_t)a.fH, (uint64_t)a.B * a.H};
    const uint64_t str[3] = {64, (uint64_t)a.fW * 64, (uint64_t)a.fH * a.fW * 64};
    const uint32_t box[4] = {32, (uint32_t)a.WW, (uint32_t)a.WH, 1};
    if (int rc = make_tensor_map(&mv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, value16, dims, str, box,
                                 CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)ld, (uint64_t)a.bev_w, (uint64_t)a.bev_h, (uint64_t)a.B};
    const uint64_t str[3] = {(uint64_t)ld * 4, (uint64_t)a.bev_w * ld * 4, (uint64_t)a.bev_h * a.bev_w * ld * 4};
    const uint32_t box_o[4] = {2 * PP, kTQ, 1, 1}, box_l[4] = {PP, kTQ, 1, 1};  // one query row per worker warp
    if (int rc = make_tensor_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, qproj, dims, str, box_o,
                                 CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
    if (int rc = make_tensor_map(&ml, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, qproj, dims, str, box_l,
                                 CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  }
  if (fixed) return launch_bev_win_v<PP, kPrefWW * 64>(a, mv, mo, ml, smem, s);
  return launch_bev_win_v<PP, 0>(a, mv, mo, ml, smem, s);
}

template <int PP, int ROWB, bool STAGE>
static int launch_img_win_v(ImgWinArgs& a, const CUtensorMap& mv, size_t smem, cudaStream_t s) {
  const char* fn = "ub_img_sample_win_fwd";
  if (int rc = ensure_smem(img_sample_win_kernel<PP, ROWB, STAGE>, smem, fn)) return rc;
  a.part = 0;
  launch_pdl(img_sample_win_kernel<PP, ROWB, STAGE>, dim3(sm_count()), dim3(kImgThreads), smem, s, a, mv);
  if (int rc = check_launch(fn)) return rc;
  a.part = 1;   // (camera, query) pairs beyond a query's first camera: ~12 % of the pairs on the nuScenes rig
  launch_pdl(img_sample_win_kernel<PP, ROWB, STAGE>, dim3(sm_count()), dim3(kImgThreads), smem, s, a, mv);
  return check_launch(fn);
}

template <int PP>
static int launch_img_win(ImgWinArgs& a, const void* value16, cudaStream_t s) {
  const char* fn = "ub_img_sample_win_fwd";
  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  // off by default: the second 100 KB window leaves the scattered P1 loads almost no L1 (measured 279 vs 196 us)
  // bulk-staged P1 inputs need 16-byte pieces: an even number of Z-anchors and a 16-byte aligned anchor tensor
  const bool stage = g_img_stage && a.D % 2 == 0 && (reinterpret_cast<uintptr_t>(a.ref_cam) & 15u) == 0 &&
                     ImgSmem<PP>::total(win_bytes, 1, true) <= kSmemBudget;
  a.two_win = g_img_two_win && ImgSmem<PP>::total(win_bytes, 2, stage) <= kSmemBudget ? 1 : 0;
  const size_t smem = ImgSmem<PP>::total(win_bytes, a.two_win ? 2 : 1, stage);
  if (smem > kSmemBudget) {
    set_error("%s: plane %d x %d needs %zu bytes of shared memory", fn, a.fH, a.fW, smem);
    return ub::unsupported();
  }
  CUtensorMap mv;
  const uint64_t dims[4] = {32, (uint64_t)a.fW, (uint64_t)a.fH, (uint64_t)a.B * a.N * a.H};
  const uint64_t str[3] = {64, (uint64_t)a.fW * 64, (uint64_t)a.fH * a.fW * 64};
  const uint32_t box[4] = {32, (uint32_t)a.WW, (uint32_t)a.WH, 1};
  if (int rc = make_tensor_map(&mv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, value16, dims, str, box,
                               CU_TENSOR_MAP_SWIZZLE_NONE))
    return rc;
  if (a.WW == 52)   // nuScenes 1600 x 928 / 32 -> 50 + 2
    return stage ? launch_img_win_v<PP, 52 * 64, true>(a, mv, smem, s) : launch_img_win_v<PP, 52 * 64, false>(a, mv, smem, s);
  return stage ? launch_img_win_v<PP, 0, true>(a, mv, smem, s) : launch_img_win_v<PP, 0, false>(a, mv, smem, s);
}

}  // namespace ub

using namespace ub;

extern "C" int ub_set_window_halo(int halo) {
  UB_REQUIRE(halo >= 0 && halo <= 64, "ub_set_window_halo: halo must be in 0..64 (0 = default, points + 1)");
  g_bev_halo = halo;
  return UB_OK;
}

extern "C" int ub_set_img_two_windows(int on) {
  g_img_two_win = on ? 1 : 0;
  return UB_OK;
}
extern "C" int ub_set_img_vec_ref(int on) {
  g_img_vec_ref = on ? 1 : 0;
  return UB_OK;
}
extern "C" int ub_set_img_stage(int on) {
  g_img_stage = on ? 1 : 0;
  return UB_OK;
}

extern "C" int ub_value_to_half(const float* value, void* value16, int G, int Nv, int H, int Dh, ub_stream_t stream) {
  UB_REQUIRE(value && value16, "ub_value_to_half: null pointer");
  UB_REQUIRE(G > 0 && Nv > 0 && H > 0 && Dh > 0 && Dh % 8 == 0, "ub_value_to_half: need positive dims and Dh %% 8 == 0");
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(value16);
  const int64_t total = (int64_t)G * H * Nv * (Dh / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  value_to_half_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(value, reinterpret_cast<uint4*>(value16), G, Nv, H, Dh);
  return check_launch("ub_value_to_half");
}

extern "C" int ub_bev_sample_win_fwd(const void* value16, const float* qproj, void* out, int out_f16, int B, int bev_h,
                                     int bev_w, int fH, int fW, int H, int Dh, int P, int ld, int off_col, int logit_col,
                                     int* workspace, int flags, ub_stream_t stream) {
  const char* fn = "ub_bev_sample_win_fwd";
  UB_REQUIRE(value16 && qproj && out && workspace, "%s: null pointer", fn);
  UB_REQUIRE(B > 0 && bev_h > 0 && bev_w > 0 && fH >= 2 && fW >= 2 && H > 0, "%s: non-positive dimension", fn);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small", fn, ld);
  UB_REQUIRE_ALIGNED16(value16);
  UB_REQUIRE_ALIGNED16(qproj);
  UB_REQUIRE_ALIGNED16(out);
  if (Dh != 32 || (P != 4 && P != 8) || ld % 4 != 0 || off_col % 4 != 0 || logit_col % 4 != 0 ||
      (int64_t)B * H > (1 << 20)) {
    set_error("%s: shape not covered by the window kernels (Dh=%d P=%d ld=%d)", fn, Dh, P, ld);
    return ub::unsupported();
  }
  BevWinArgs a;
  a.value16 = reinterpret_cast<const __half*>(value16), a.counters = workspace;
  a.out = out_f16 ? nullptr : reinterpret_cast<float*>(out), a.out16 = out_f16 ? reinterpret_cast<__half*>(out) : nullptr;
  a.B = B, a.bev_h = bev_h, a.bev_w = bev_w, a.fH = fH, a.fW = fW, a.H = H;
  a.tiles_x = (bev_w + kTQ - 1) / kTQ, a.tiles_y = (bev_h + kTQ - 1) / kTQ;
  a.n_units = B * H * a.tiles_x * a.tiles_y;
  a.sx = (float)fW / (float)bev_w, a.sy = (float)fH / (float)bev_h;
  a.R = g_bev_halo > 0 ? g_bev_halo : P + 1;
  a.WW = (int)ceilf((kTQ - 1) * a.sx) + 2 * a.R + 3;
  a.WH = (int)ceilf((kTQ - 1) * a.sy) + 2 * a.R + 3;
  a.off_col = off_col, a.logit_col = logit_col;
  a.round_tf32 = flags & 1;
  if (a.WW > 256 || a.WH > 256) {
    set_error("%s: window %d x %d exceeds the TMA box limit", fn, a.WW, a.WH);
    return ub::unsupported();
  }
  return P == 8 ? launch_bev_win<8>(a, value16, qproj, ld, (cudaStream_t)stream)
                : launch_bev_win<4>(a, value16, qproj, ld, (cudaStream_t)stream);
}

extern "C" int ub_build_hits(const uint8_t* mask, int* hit_idx, int* hit_cnt, float* inv_cnt, float* hit_ic, int B, int N,
                             int Nq, ub_stream_t stream) {
  UB_REQUIRE(mask && hit_idx && hit_cnt && inv_cnt, "ub_build_hits: null pointer");
  UB_REQUIRE(B > 0 && N > 0 && N <= 32 && Nq > 0, "ub_build_hits: bad dimension (B=%d N=%d Nq=%d)", B, N, Nq);
  int extra = (int)(((int64_t)B * Nq + 1023) / 1024);
  if (extra > sm_count()) extra = sm_count();
  build_hits_kernel<<<N + 1 + extra, 1024, 0, (cudaStream_t)stream>>>(mask, hit_idx, hit_cnt, inv_cnt, B, N, Nq);
  if (int rc = check_launch("ub_build_hits")) return rc;
  if (hit_ic) {
    int blocks = (int)(((int64_t)B * N * Nq + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    hit_ic_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(hit_idx, hit_cnt, inv_cnt, hit_ic, B, N, Nq);
    return check_launch("ub_build_hits");
  }
  return UB_OK;
}

extern "C" int ub_img_sample_win_fwd(const void* value16, const float* qproj, const float* ref_cam, const int* hit_idx,
                                     const int* hit_cnt, const float* inv_cnt, const float* hit_ic, void* out, int out_f16,
                                     int B, int N,
                                     int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P, int D, int ld, int off_col,
                                     int logit_col, ub_stream_t stream) {
  const char* fn = "ub_img_sample_win_fwd";
  UB_REQUIRE(value16 && qproj && ref_cam && hit_idx && hit_cnt && inv_cnt && out, "%s: null pointer", fn);
  UB_REQUIRE(B > 0 && N > 0 && N <= 32 && bev_h > 0 && bev_w > 0 && fH >= 2 && fW >= 2 && H > 0 && D > 0 && D <= 8,
             "%s: bad dimension (B=%d N=%d D=%d)", fn, B, N, D);
  UB_REQUIRE(P % D == 0, "%s: num_points %d must be a multiple of the %d Z-anchors", fn, P, D);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small", fn, ld);
  UB_REQUIRE_ALIGNED16(value16);
  UB_REQUIRE_ALIGNED16(out);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(ref_cam) & 7u) == 0 && (reinterpret_cast<uintptr_t>(qproj) & 7u) == 0,
             "%s: ref_cam / qproj not 8-byte aligned", fn);
  if (Dh != 32 || (P != 4 && P != 8) || ld % 4 != 0 || off_col % 4 != 0 || logit_col % 4 != 0 ||
      (reinterpret_cast<uintptr_t>(qproj) & 15u) != 0 || fW + 2 > 256 || fH + 2 > 256 ||
      (int64_t)(fH + 2) * (fW + 2) > 65535 || (int64_t)B * N * H > (1 << 20) || H > sm_count()) {
    set_error("%s: shape not covered by the window kernels (Dh=%d P=%d fH=%d fW=%d)", fn, Dh, P, fH, fW);
    return ub::unsupported();
  }
  ImgWinArgs a;
  a.qproj = qproj, a.ref_cam = ref_cam, a.inv_cnt = inv_cnt, a.hit_ic = hit_ic, a.hit_idx = hit_idx, a.hit_cnt = hit_cnt;
  a.out = out_f16 ? nullptr : reinterpret_cast<float*>(out), a.out16 = out_f16 ? reinterpret_cast<__half*>(out) : nullptr;
  a.B = B, a.N = N, a.Nq = bev_h * bev_w, a.fH = fH, a.fW = fW, a.H = H, a.P = P, a.D = D;
  a.ld = ld, a.off_col = off_col, a.logit_col = logit_col;
  a.WW = fW + 2, a.WH = fH + 2;
  a.vec_ref = g_img_vec_ref && D % (P / 2) == 0 && (reinterpret_cast<uintptr_t>(ref_cam) & 15u) == 0;
  a.vec8 = g_img_vec_ref && P == 8 && D == 4 && (reinterpret_cast<uintptr_t>(ref_cam) & 31u) == 0 &&
           ((reinterpret_cast<uintptr_t>(qproj) + (size_t)off_col * 4) & 31u) == 0 && (ld * 4) % 32 == 0;
  return P == 8 ? launch_img_win<8>(a, value16, (cudaStream_t)stream) : launch_img_win<4>(a, value16, (cudaStream_t)stream);
}
