// fp32 window-staged deformable sampling: the sampling kernels of the default ('fp32') precision class.
//
//   ub_bev_sample_win32_fwd  : BEV self-attention / LiDAR cross-attention sampling, fp32 value maps       [R4]
//
// Same organisation as the fp16-staged kernels of win_sample.cu (scheduler warp + 16 worker warps per persistent CTA,
// windows handed over through full / empty mbarriers, descriptors once per sample, eight lanes per item) with the
// arithmetic of the reference: fp32 value maps, fp32 bilinear x attention weights, expf / true division in the softmax.
//
// An fp32 head of one pixel is 128 B, and a window of fp32 heads does not fit shared memory twice.  The value map is
// therefore stored as HALF-HEAD planes (G, 2 H, fH*fW, 16) fp32 -- written in that layout by the value projection's
// epilogue (ub_linear_tf32x3, planes32) -- so one pixel of one half-head is 64 B, byte for byte the geometry of the fp16
// kernels: the two horizontal neighbours of a sample are one 128-byte shared-memory row segment.  A unit (16 x 16
// queries x one head) streams its two half-head windows through the two window buffers: the workers build the
// descriptors once, gather the first half-head while the second window lands, gather the second while the next unit's
// first window lands, then write whole 128-byte rows (both halves) with one store per lane.
// Cost model: 4 shared-memory wavefronts per sample (2 rows x 128 B per half-head) against 2 for fp16 planes: the
// L1 / shared-memory data pipe (one 128-byte wavefront per clock and SM), not HBM, bounds these kernels at
// 8 points; see DESIGN.md section 3.
#include "win_common.cuh"

namespace ub {

// Per-warp descriptor buffer of the warp's 16 items: per sample four fp32 weights {left top, left bottom, right top,
// right bottom} (attention weight folded in), then 16-bit window pixel indices.  The weights are stored per pixel SIDE and
// per PAIR of sampling points -- {top, bottom} of point 2 j, {top, bottom} of point 2 j + 1 -- so that a gathering lane
// (which owns one side) fetches two points' weights with one 16-byte load.  Item stride and side offset are chosen so
// that the 4 items x 2 sides x 16 bytes one warp instruction reads fall into 32 different banks (one wavefront):
//   P = 8: [side][pair] (side offset 16 words), item stride 36 words;  P = 4: [pair][side] (side offset 4), stride 24.
template <int PP>
struct Desc32 {
  static constexpr int w_stride = PP == 8 ? 36 : PP * 4 + 8;  // words per item
  static constexpr int w_bytes = kWarpItems * w_stride * 4;
  static constexpr int idx_bytes = kWarpItems * PP * 2;
  static constexpr int bytes = w_bytes + idx_bytes;
  // word offset of the float4 {w_top(2 j), w_bot(2 j), w_top(2 j + 1), w_bot(2 j + 1)} of pixel side `side`
  __device__ static constexpr int pair_off(int side, int j) { return PP == 8 ? side * 16 + j * 4 : j * 8 + side * 4; }
};

// softmax over the item's P logits (two lanes per item, PPL points each): expf and a true division
template <int PPL>
__device__ __forceinline__ void softmax_pair32(const float (&lg)[PPL], bool ok, float scale, float (&aw)[PPL]) {
  float mx = lg[0];
#pragma unroll
  for (int i = 1; i < PPL; ++i) mx = fmaxf(mx, lg[i]);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    aw[i] = expf(lg[i] - mx);
    sum += aw[i];
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = ok ? scale / sum : 0.f;
#pragma unroll
  for (int i = 0; i < PPL; ++i) aw[i] *= inv;
}

// One sample -> four fp32 corner weights + window pixel index (branch-free).  Returns true when the sample touches the
// map but its 2 x 2 footprint is not inside the window ("far").  Zero padding: the window holds zeros outside the map
// (TMA out-of-bounds fill), so corners beyond the border contribute nothing, as in the reference.
__device__ __forceinline__ bool make_desc32(bool ok, float h_im, float w_im, float aw, int fH, int fW, int wy0, int wx0,
                                            int WW, int WH, float4& w4, uint32_t& idx) {
  const bool inmap = ok & (h_im > -1.f) & (w_im > -1.f) & (h_im < (float)fH) & (w_im < (float)fW);
  const int y0 = __float2int_rd(h_im), x0 = __float2int_rd(w_im);   // saturating: wild coordinates are harmless
  const float lh = h_im - (float)y0, lw = w_im - (float)x0;
  const int yy = y0 - wy0, xx = x0 - wx0;
  const bool inwin = ((unsigned)xx < (unsigned)(WW - 1)) & ((unsigned)yy < (unsigned)(WH - 1));
  const bool use = inmap & inwin;
  const float a2 = use ? aw : 0.f;
  const float hh = 1.f - lh, hw = 1.f - lw;
  w4 = make_float4(a2 * (hh * hw), a2 * (lh * hw), a2 * (hh * lw), a2 * (lh * lw));   // {lt, lb, rt, rb}
  idx = use ? (uint32_t)(yy * WW + xx) : 0u;
  return inmap & !inwin;
}

template <int PP>
__device__ __forceinline__ void store_descs32(uint32_t sm_w, uint32_t sm_idx, int item, int p0, const float4 (&w4)[PP / 2],
                                              const uint32_t (&idx)[PP / 2]) {
  constexpr int PPL = PP / 2;
  const uint32_t wa = sm_w + (uint32_t)(item * Desc32<PP>::w_stride) * 4u;
#pragma unroll
  for (int i = 0; i < PPL; i += 2) {   // this lane's point pairs p0 / 2 + i / 2
    const int j = p0 / 2 + i / 2;
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(wa + (uint32_t)Desc32<PP>::pair_off(0, j) * 4u), "f"(w4[i].x),
                 "f"(w4[i].y), "f"(w4[i + 1].x), "f"(w4[i + 1].y)
                 : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(wa + (uint32_t)Desc32<PP>::pair_off(1, j) * 4u), "f"(w4[i].z),
                 "f"(w4[i].w), "f"(w4[i + 1].z), "f"(w4[i + 1].w)
                 : "memory");
  }
  const uint32_t ia = sm_idx + (uint32_t)(item * PP + p0) * 2u;
  if (PPL == 4)
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(ia), "r"(idx[0] | (idx[1] << 16)), "r"(idx[2] | (idx[PPL - 1] << 16))
                 : "memory");
  else
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ia), "r"(idx[0] | (idx[1] << 16)) : "memory");
}

__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// The gather of one warp's 16 items from ONE half-head window: lane group `grp` (8 lanes) reduces items grp, grp + 4,
// grp + 8, grp + 12, two at a time; lanes 0-3 of a group own the left pixel, 4-7 the right one, four channels each.
// res[m] = this lane's four channels (cq * 4 ...) of item grp + 4 m, both pixel sides summed (held by both side lanes).
// n_live (warp-uniform): only items < n_live are real (tile columns beyond the BEV grid / list positions beyond the hit
// list); a pair of item quads with no real item is skipped (res left untouched).
template <int PP, int ROWB>
__device__ __forceinline__ void gather_warp32(uint32_t sm_w, uint32_t sm_idx, uint32_t win, uint32_t row_rt, int grp,
                                              int side, int n_live, float4 (&res)[4]) {
  const uint32_t row_b = ROWB > 0 ? (uint32_t)ROWB : row_rt;
#pragma unroll
  for (int k = 0; k < 4; k += 2) {
    if (k * 4 >= n_live) break;
    const int item[2] = {grp + k * 4, grp + (k + 1) * 4};
    uint32_t ix[2][4], wa[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      wa[j] = sm_w + (uint32_t)(item[j] * Desc32<PP>::w_stride + side * Desc32<PP>::pair_off(1, 0)) * 4u;
      if (PP == 8) {
        const uint4 t = lds128(sm_idx + item[j] * 16);
        ix[j][0] = t.x, ix[j][1] = t.y, ix[j][2] = t.z, ix[j][3] = t.w;
      } else {
        uint32_t a, b;
        asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(sm_idx + item[j] * 8));
        ix[j][0] = a, ix[j][1] = b, ix[j][2] = 0, ix[j][3] = 0;
      }
    }
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
#pragma unroll
    for (int pp = 0; pp < PP / 2; ++pp) {
      float4 w[2];   // {top, bottom} of point 2 pp, {top, bottom} of point 2 pp + 1, this lane's pixel side
#pragma unroll
      for (int j = 0; j < 2; ++j)
        w[j] = lds128f(wa[j] + (uint32_t)Desc32<PP>::pair_off(0, pp) * 4u);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float4 top[2], bot[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t word = ix[j][pp];
          const uint32_t id = q ? (word >> 16) : (word & 0xffffu);
          const uint32_t a = win + id * 64u;
          top[j] = lds128f(a);
          bot[j] = lds128f(a + row_b);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          fma4(acc[j], q ? w[j].z : w[j].x, top[j]);
          fma4(acc[j], q ? w[j].w : w[j].y, bot[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {   // left + right pixel
      float4 o;
      o.x = acc[j].x + __shfl_xor_sync(0xffffffffu, acc[j].x, 4);
      o.y = acc[j].y + __shfl_xor_sync(0xffffffffu, acc[j].y, 4);
      o.z = acc[j].z + __shfl_xor_sync(0xffffffffu, acc[j].z, 4);
      o.w = acc[j].w + __shfl_xor_sync(0xffffffffu, acc[j].w, 4);
      res[k + j] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct BevWin32Args {
  const float* planes;    // (B * 2H, fH, fW, 16): far path
  float* out;             // (B, Nq, H*32) fp32 rows
  int* counters;          // [0] next unit, [1] CTAs done (caller-owned, zero between calls)
  int B, bev_h, bev_w, fH, fW, H;
  int tiles_x, tiles_y, n_units;
  int WW, WH, R;
  int off_col, logit_col;
  float sx, sy;
  int NB;                 // window buffers: 2, or 3 where shared memory allows (hides the window's load latency)
};

template <int PP>
struct BevSmem32 {
  static constexpr int slice_off_bytes = kWarpItems * PP * 8, slice_lg_bytes = kWarpItems * PP * 4;
  static constexpr int slice_bytes = slice_off_bytes + slice_lg_bytes;      // one query row of the offset|logit tile
  static constexpr int warp_bytes = slice_bytes + ((Desc32<PP>::bytes + 127) & ~127);
  static size_t total(int win_bytes, int nb = 2) { return (size_t)nb * win_bytes + (size_t)kWorkerWarps * warp_bytes; }
};

// Worker warp w owns query row ty0 + w of the unit's 16 x 16 tile.  Its loop per unit k:
//   P1  descriptors of its 128 / 64 samples from its TMA-staged slice of the offset|logit rows (own mbarrier)
//   --  issue the slice of unit k + 1 (the buffer is free), slow path for far samples
//   P2  wait for half-head window 0 (full[0]), gather, release it (empty[0]); the same for window 1; write the rows
// The scheduler warp hands out units (atomic counter), publishes them through a 4-slot ring and streams half-head
// window s of unit k into buffer s as soon as every worker has released it (unit k - 1).
template <int PP, int ROWB>
__global__ void __launch_bounds__(kBevThreads, 1)
    bev_sample_win32_kernel(const BevWin32Args a, const __grid_constant__ CUtensorMap map_val,
                            const __grid_constant__ CUtensorMap map_off, const __grid_constant__ CUtensorMap map_lg) {
  using S = BevSmem32<PP>;
  using D = Desc32<PP>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[3], s_empty[3], s_unit[4], s_qp[kWorkerWarps];
  __shared__ UnitInfo s_ring[4];

  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  const uint32_t sm_win = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Nq = a.bev_h * a.bev_w, C = a.H * 32;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&s_full[i]), 1), mbar_init(smem_u32(&s_empty[i]), kWorkerWarps);
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
    int sched_buf = 0, sched_use = 0;
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
      // ring slot k & 3 held unit k - 4; every worker finished unit k - 2 before the windows of unit k - 1 were issued
      s_ring[k & 3] = w;
      mbar_arrive(smem_u32(&s_unit[k & 3]));
      if (w.u < 0) break;
      // half-head window i = 2 k + s of the CTA's stream goes to buffer i % NB; its n-th use (n = i / NB) waits for the
      // workers' release of use n - 1
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (sched_use > 0) mbar_wait(smem_u32(&s_empty[sched_buf]), (uint32_t)((sched_use - 1) & 1));
        const uint32_t bar = smem_u32(&s_full[sched_buf]);
        mbar_arrive_expect_tx(bar, (uint32_t)(a.WW * a.WH * 64));
        tma_load_4d(sm_win + (uint32_t)sched_buf * win_bytes, &map_val, bar, 0, w.wx0, w.wy0, (w.b * a.H + w.h) * 2 + s);
        if (++sched_buf == a.NB) sched_buf = 0, ++sched_use;
      }
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
  const uint32_t sm_slice = sm_win + (uint32_t)a.NB * win_bytes + (uint32_t)warp * S::warp_bytes;  // {offsets, logits}
  const uint32_t sm_w = sm_slice + S::slice_bytes, sm_idx = sm_w + D::w_bytes;
  const uint32_t bar_qp = smem_u32(&s_qp[warp]);
  const uint32_t bar_unit = smem_u32(&s_unit[0]), bar_full = smem_u32(&s_full[0]), bar_empty = smem_u32(&s_empty[0]);
  constexpr int PPL = PP / 2;                // sampling points per lane in P1 (two lanes per item)
  const int item_l = lane >> 1, p0 = (lane & 1) * PPL;
  const int grp = lane >> 3, sub = lane & 7, side = sub >> 2, cq = sub & 3;

  auto issue_slice = [&](const UnitInfo& w) {  // lane 0: this warp's query row of the offset|logit tile
    mbar_arrive_expect_tx(bar_qp, (uint32_t)S::slice_bytes);
    tma_load_4d(sm_slice, &map_off, bar_qp, a.off_col + w.h * PP * 2, w.tx0, w.ty0 + warp, w.b);
    tma_load_4d(sm_slice + S::slice_off_bytes, &map_lg, bar_qp, a.logit_col + w.h * PP, w.tx0, w.ty0 + warp, w.b);
  };

  mbar_wait(bar_unit, 0u);
  UnitInfo w = s_ring[0];
  if (w.u >= 0 && lane == 0) issue_slice(w);
  int w_buf = 0, w_use = 0;   // buffer and use count of the next half-head window of this CTA's stream

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
    softmax_pair32<PPL>(lg, ok, 1.f, far_a);
    unsigned far_bits = 0;
    {
      // pixel = ((q + .5) / bev + off / f) * f - .5  ==  (q + .5) * (f / bev) + off - .5
      const float wbase = (float)qx1 + 0.5f;
      float4 w4[PPL];
      uint32_t idx[PPL];
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        far_w[i] = fmaf(wbase, a.sx, off[2 * i] - 0.5f);
        far_h[i] = fmaf(hbase, a.sy, off[2 * i + 1] - 0.5f);
        if (make_desc32(ok, far_h[i], far_w[i], far_a[i], a.fH, a.fW, w.wy0, w.wx0, a.WW, a.WH, w4[i], idx[i]))
          far_bits |= 1u << i;
      }
      store_descs32<PP>(sm_w, sm_idx, item_l, p0, w4, idx);
    }
    const bool any_far = __any_sync(0xffffffffu, far_bits != 0u);
    __syncwarp();   // descriptor stores visible to the whole warp
    // ---- the slice buffer is free: stream in the next unit's row
    mbar_wait(bar_unit + 8u * (uint32_t)((k + 1) & 3), (uint32_t)(((k + 1) >> 2) & 1));
    const UnitInfo wn = s_ring[(k + 1) & 3];
    if (wn.u >= 0 && lane == 0) issue_slice(wn);

    // ---- slow path for samples outside the staged window (per-corner bounds checks, straight from global memory)
    if (any_far) {
      if (row_ok) {
#pragma unroll
        for (int e = lane; e < kWarpItems * 8; e += 32) {
          const int item = e >> 3, c4 = e & 7;
          if (w.tx0 + item < a.bev_w) {
            const int64_t o = ((int64_t)w.b * Nq + qy * a.bev_w + w.tx0 + item) * C + w.h * 32 + c4 * 4;
            *reinterpret_cast<float4*>(a.out + o) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      __syncwarp();
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
          // lane = corner (2 bits) x channel quad (3 bits: half-head c4 >> 2, quad c4 & 3)
          const int corner = lane >> 3, c4 = lane & 7, dy = corner >> 1, dx = corner & 1;
          const float hf = floorf(h_im), wf = floorf(w_im);
          const float lh = h_im - hf, lw = w_im - wf;
          const int y = (int)hf + dy, x = (int)wf + dx;
          const float wgt = aw * ((dy ? lh : 1.f - lh) * (dx ? lw : 1.f - lw));
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (y >= 0 && y < a.fH && x >= 0 && x < a.fW) {
            const float* plane = a.planes + (int64_t)((w.b * a.H + w.h) * 2 + (c4 >> 2)) * a.fH * a.fW * 16;
            const float4 f = ldg4(plane + ((int64_t)y * a.fW + x) * 16 + (c4 & 3) * 4);
            v = make_float4(wgt * f.x, wgt * f.y, wgt * f.z, wgt * f.w);
          }
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o), v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, o), v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
          }
          if (lane < 8) {
            const int64_t o = ((int64_t)w.b * Nq + qy * a.bev_w + w.tx0 + item) * C + w.h * 32 + c4 * 4;
            red_add4(a.out + o, v);
          }
        }
      }
      __syncwarp();
    }

    // ---- P2: the two half-head windows of the unit
    float4 r0[4], r1[4];
    // real items of this warp's query row: none below the grid, 16 or fewer in the last tile column
    const int n_live = row_ok ? min(kWarpItems, a.bev_w - w.tx0) : 0;
    mbar_wait(bar_full + 8u * (uint32_t)w_buf, (uint32_t)(w_use & 1));
    gather_warp32<PP, ROWB>(sm_w, sm_idx, sm_win + (uint32_t)w_buf * win_bytes + sub * 16u, (uint32_t)a.WW * 64u, grp, side, n_live,
                            r0);
    __syncwarp();   // every lane is done with window 0
    if (lane == 0) mbar_arrive(bar_empty + 8u * (uint32_t)w_buf);
    if (++w_buf == a.NB) w_buf = 0, ++w_use;
    mbar_wait(bar_full + 8u * (uint32_t)w_buf, (uint32_t)(w_use & 1));
    gather_warp32<PP, ROWB>(sm_w, sm_idx, sm_win + (uint32_t)w_buf * win_bytes + sub * 16u, (uint32_t)a.WW * 64u, grp, side, n_live,
                            r1);
    __syncwarp();   // ... and with window 1 and the descriptors
    if (lane == 0) mbar_arrive(bar_empty + 8u * (uint32_t)w_buf);
    if (++w_buf == a.NB) w_buf = 0, ++w_use;
    // rows: lanes of pixel side 0 write the first half-head's channels, side 1 the second's: 128 B per item and store
    if (row_ok) {
      const int64_t out0 = ((int64_t)w.b * Nq + qy * a.bev_w + w.tx0) * C + w.h * 32 + side * 16 + cq * 4;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int item = grp + 4 * m;
        if (w.tx0 + item < a.bev_w) {
          const float4 o = side ? r1[m] : r0[m];
          float* dst = a.out + out0 + (int64_t)item * C;
          if (any_far)
            red_add4(dst, o);
          else
            st_stream4(dst, o);
        }
      }
    }
    w = wn;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side

template <int PP, int ROWB>
static int launch_bev_win32_v(BevWin32Args& a, const CUtensorMap& mv, const CUtensorMap& mo, const CUtensorMap& ml,
                              size_t smem, cudaStream_t s) {
  const char* fn = "ub_bev_sample_win32_fwd";
  if (int rc = ensure_smem(bev_sample_win32_kernel<PP, ROWB>, smem, fn)) return rc;
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  launch_pdl(bev_sample_win32_kernel<PP, ROWB>, dim3(grid), dim3(kBevThreads), smem, s, a, mv, mo, ml);
  return check_launch(fn);
}

static int g_bev_win32_buffers = 0;   // A/B knob (tools): 0 = as many as fit (up to 3), 2 = always two

template <int PP>
static int launch_bev_win32(BevWin32Args& a, const float* planes, const float* qproj, int ld, cudaStream_t s) {
  const char* fn = "ub_bev_sample_win32_fwd";
  // preferred row pitch with a compile-time kernel variant; otherwise shrink the halo until two windows fit
  // (samples beyond it stay exact through the slow path)
  constexpr int kPrefWW = PP == 8 ? 36 : 28;
  auto smem_for = [&]() { return BevSmem32<PP>::total((a.WW * a.WH * 64 + 127) & ~127); };
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
  size_t smem = smem_for();
  if (smem > kSmemBudget) {
    set_error("%s: window %d x %d needs %zu bytes of shared memory", fn, a.WW, a.WH, smem);
    return ub::unsupported();
  }
  // a third window buffer where it fits (4 points: 28 x 28 windows): the next unit's first window then loads while both
  // windows of the current unit are still in use
  a.NB = 2;
  if (g_bev_win32_buffers != 2 && BevSmem32<PP>::total((a.WW * a.WH * 64 + 127) & ~127, 3) <= kSmemBudget) {
    a.NB = 3;
    smem = BevSmem32<PP>::total((a.WW * a.WH * 64 + 127) & ~127, 3);
  }
  CUtensorMap mv, mo, ml;
  {
    const uint64_t dims[4] = {16, (uint64_t)a.fW, (uint64_t)a.fH, (uint64_t)a.B * a.H * 2};
    const uint64_t str[3] = {64, (uint64_t)a.fW * 64, (uint64_t)a.fH * a.fW * 64};
    const uint32_t box[4] = {16, (uint32_t)a.WW, (uint32_t)a.WH, 1};
    if (int rc = make_tensor_map(&mv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, planes, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
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
  if (fixed) return launch_bev_win32_v<PP, kPrefWW * 64>(a, mv, mo, ml, smem, s);
  return launch_bev_win32_v<PP, 0>(a, mv, mo, ml, smem, s);
}

extern int g_bev_halo_shared();

}  // namespace ub

using namespace ub;

extern "C" int ub_set_bev_win32_buffers(int n) {
  UB_REQUIRE(n == 0 || n == 2, "ub_set_bev_win32_buffers: 0 (as many as fit) or 2");
  ub::g_bev_win32_buffers = n;
  return UB_OK;
}

extern "C" int ub_bev_sample_win32_fwd(const float* planes32, const float* qproj, float* out, int B, int bev_h, int bev_w,
                                       int fH, int fW, int H, int Dh, int P, int ld, int off_col, int logit_col,
                                       int* workspace, ub_stream_t stream) {
  const char* fn = "ub_bev_sample_win32_fwd";
  UB_REQUIRE(planes32 && qproj && out && workspace, "%s: null pointer", fn);
  UB_REQUIRE(B > 0 && bev_h > 0 && bev_w > 0 && fH >= 2 && fW >= 2 && H > 0, "%s: non-positive dimension", fn);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small", fn, ld);
  UB_REQUIRE_ALIGNED16(planes32);
  UB_REQUIRE_ALIGNED16(qproj);
  UB_REQUIRE_ALIGNED16(out);
  if (Dh != 32 || (P != 4 && P != 8) || ld % 4 != 0 || off_col % 4 != 0 || logit_col % 4 != 0 ||
      (int64_t)B * H * 2 > (1 << 20)) {
    set_error("%s: shape not covered by the window kernels (Dh=%d P=%d ld=%d)", fn, Dh, P, ld);
    return ub::unsupported();
  }
  BevWin32Args a;
  a.planes = planes32, a.out = out, a.counters = workspace;
  a.B = B, a.bev_h = bev_h, a.bev_w = bev_w, a.fH = fH, a.fW = fW, a.H = H;
  a.tiles_x = (bev_w + kTQ - 1) / kTQ, a.tiles_y = (bev_h + kTQ - 1) / kTQ;
  a.n_units = B * H * a.tiles_x * a.tiles_y;
  a.sx = (float)fW / (float)bev_w, a.sy = (float)fH / (float)bev_h;
  const int halo = g_bev_halo_shared();
  a.R = halo > 0 ? halo : P + 1;
  a.WW = (int)ceilf((kTQ - 1) * a.sx) + 2 * a.R + 3;
  a.WH = (int)ceilf((kTQ - 1) * a.sy) + 2 * a.R + 3;
  a.off_col = off_col, a.logit_col = logit_col;
  if (a.WW > 256 || a.WH > 256) {
    set_error("%s: window %d x %d exceeds the TMA box limit", fn, a.WW, a.WH);
    return ub::unsupported();
  }
  return P == 8 ? launch_bev_win32<8>(a, planes32, qproj, ld, (cudaStream_t)stream)
                : launch_bev_win32<4>(a, planes32, qproj, ld, (cudaStream_t)stream);
}

namespace ub {

// =========================================================================================================
// Camera mode (fp32).
//
// The camera kernels take their per-hit inputs in HIT-LIST ORDER, so that a worker warp's 16 hits are 16 consecutive
// rows that one TMA box copy per tensor brings into its shared-memory slice (no scattered loads in the sampling kernel):
//   qp_hit  (B, N, Nq, ld)   offset|logit rows: written in that order by the projection's epilogue (ub_linear_tf32x3 with
//                            the scatter map q_dst)
//   hit_ref  (B, N, Nq, 2 D)  projected anchors                                   } ub_hit_order, once per frame (they do
//   hit_meta (B, N, Nq, 4)    {query index (int bits), 1 / #cameras, 0, 0}        } not depend on the layer)
// (16-byte records: the hit position is then never the innermost TMA coordinate, whose byte offset must be 16-byte aligned)
// Position p of camera n's row holds a FIRST hit when p < cnt_first[n] and a LATER hit when p >= Nq - cnt_later[n]
// (ub_build_hits' rank-split layout); rows in between are never read as valid.

// q_dst (Nq, N): for query q the destination rows n * Nq + pos of its offset|logit row, one per camera that sees it
// (batch item 0's visibility, the reference's quirk), -1 padded.  hit_ref: anchors gathered into hit-list order.
__global__ void __launch_bounds__(256) hit_order_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ ref_cam,
                                                        const int* __restrict__ hit_idx, const int* __restrict__ hit_cnt,
                                                        const float* __restrict__ inv_cnt, int* __restrict__ q_dst,
                                                        float* __restrict__ hit_ref, float4* __restrict__ hit_meta, int B, int N,
                                                        int Nq, int D2) {
  pdl_trigger();
  const int64_t total = (int64_t)B * N * Nq;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % Nq), n = (int)((t / Nq) % N), b = (int)(t / ((int64_t)Nq * N));
    const bool valid = i < hit_cnt[n] || i >= Nq - hit_cnt[N + n];      // i as a position of camera n's row
    if (valid) {
      const int q = hit_idx[(int64_t)n * Nq + i];
      const float* src = ref_cam + (((int64_t)b * Nq + q) * N + n) * D2;
      float* dst = hit_ref + (((int64_t)b * N + n) * Nq + i) * D2;
      for (int e = 0; e < D2; ++e) dst[e] = src[e];
      hit_meta[((int64_t)b * N + n) * Nq + i] = make_float4(__int_as_float(q), inv_cnt[(int64_t)b * Nq + q], 0.f, 0.f);
      if (b == 0) {
        int slot = 0;
        for (int m = 0; m < n; ++m) slot += mask[(int64_t)q * N + m] != 0 ? 1 : 0;
        q_dst[(int64_t)q * N + slot] = n * Nq + i;
      }
    }
    if (b == 0) {                                                      // i as a query, n as a slot: pad the unused slots
      int cnt = 0;
      for (int m = 0; m < N; ++m) cnt += mask[(int64_t)i * N + m] != 0 ? 1 : 0;
      if (n >= cnt) q_dst[(int64_t)i * N + n] = -1;
    }
  }
}

struct ImgWin32Args {
  const int* hit_idx;     // (N + 1, Nq): row N = unseen queries
  const int* hit_cnt;     // (2 N + 1): first counts, later counts, unseen count
  float* out;             // (B, Nq, H*32)
  int B, N, Nq, fH, fW, H, D, off_col, logit_col;
  int WW, WH;
  int part;               // 0: first hits (plain stores) + zero rows of the unseen queries; 1: later hits (red.add)
};

template <int PP>
struct ImgSmem32 {
  // per warp: staged P1 inputs of its 16 hits, then the descriptors
  static constexpr int sl_off = kWarpItems * PP * 8, sl_lg = kWarpItems * PP * 4, sl_ref = kWarpItems * 64, sl_meta = kWarpItems * 16;
  static constexpr int slice_bytes = sl_off + sl_lg + sl_ref + sl_meta;
  static constexpr int warp_bytes = ((slice_bytes + 127) & ~127) + ((Desc32<PP>::bytes + 127) & ~127);
  static size_t total(int win_bytes) { return (size_t)win_bytes + (size_t)kWorkerWarps * warp_bytes; }
};

// CTAs are grouped by 2 H: the 2 H CTAs of a group walk the same contiguous range of units (b, camera, chunk of 256 hits),
// one HALF-HEAD each -- the window is that half-head's whole camera plane plus a one-pixel zero halo (reloaded when the
// range crosses into another camera) -- so a hit's input rows come from DRAM once and from L2 for the other CTAs of the
// group.  Warp w owns hits 16 w .. 16 w + 15 of the chunk: P1 from its TMA-staged slice (issued one unit ahead), P2
// gathers its 16 items from the window and writes / accumulates 64 bytes (16 channels) per hit.
template <int PP, int ROWB>
__global__ void __launch_bounds__(kImgThreads, 1)
    img_sample_win32_kernel(const ImgWin32Args a, const __grid_constant__ CUtensorMap map_val,
                            const __grid_constant__ CUtensorMap map_off, const __grid_constant__ CUtensorMap map_lg,
                            const __grid_constant__ CUtensorMap map_ref, const __grid_constant__ CUtensorMap map_meta) {
  using D = Desc32<PP>;
  using SM = ImgSmem32<PP>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar, s_qp[kWorkerWarps];

  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sm_win = smem_u32(smem), bar = smem_u32(&s_bar);
  const uint32_t sl_off = sm_win + (uint32_t)win_bytes + (uint32_t)warp * SM::warp_bytes;
  const uint32_t sl_lg = sl_off + SM::sl_off, sl_ref = sl_lg + SM::sl_lg, sl_meta = sl_ref + SM::sl_ref;
  const uint32_t sm_w = sl_off + ((SM::slice_bytes + 127) & ~127), sm_idx = sm_w + D::w_bytes;
  const uint32_t bar_qp = smem_u32(&s_qp[warp]);
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
      st_stream4(a.out + ((r / n_zero) * a.Nq + q) * C + c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  int chunks_tot = 0;
  for (int n = 0; n < a.N; ++n) chunks_tot += (__ldg(cnt_p + n) + kUnitItems - 1) / kUnitItems;
  const int G = 2 * a.H;
  const int n_grp = (int)gridDim.x / G, cgrp = (int)blockIdx.x / G, my_hs = (int)blockIdx.x % G;
  if (cgrp >= n_grp) return;
  const int my_h = my_hs >> 1, my_s = my_hs & 1;
  const int total = chunks_tot * a.B;
  const int per = total / n_grp, rem = total % n_grp;
  const int u_beg = cgrp * per + min(cgrp, rem);
  const int u_end = u_beg + per + (cgrp < rem ? 1 : 0);
  if (u_beg >= u_end) return;   // uniform per CTA

  struct Unit {
    int b, n, pos0, limit, plane;   // pos0: first hit-list position of the chunk; positions < limit are valid
  };
  auto decode = [&](int uu) {
    Unit w;
    w.b = uu / chunks_tot;
    int r = uu % chunks_tot;
    w.n = 0, w.pos0 = 0, w.limit = 0;
    for (int n = 0; n < a.N; ++n) {
      const int cnt = __ldg(cnt_p + n), ch = (cnt + kUnitItems - 1) / kUnitItems;
      if (r < ch) {
        w.n = n;
        w.pos0 = (a.part ? a.Nq - cnt : 0) + r * kUnitItems;
        w.limit = a.part ? a.Nq : cnt;
        break;
      }
      r -= ch;
    }
    w.plane = ((w.b * a.N + w.n) * a.H + my_h) * 2 + my_s;
    return w;
  };
  auto load_plane = [&](int plane) {   // one thread
    mbar_arrive_expect_tx(bar, (uint32_t)(a.WW * a.WH * 64));
    tma_load_4d(sm_win, &map_val, bar, 0, -1, -1, plane);
  };
  Unit w = decode(u_beg);
  if (tid == 0) {
    mbar_init(bar, 1);
    for (int i = 0; i < kWorkerWarps; ++i) mbar_init(smem_u32(&s_qp[i]), 1);
    mbar_init_fence();
    tma_prefetch_desc(&map_val);
    load_plane(w.plane);
  }
  __syncthreads();

  constexpr int PPL = PP / 2;                // sampling points per lane in P1 (two lanes per item)
  const int item_l = lane >> 1, p0 = (lane & 1) * PPL;
  const int grp = lane >> 3, sub = lane & 7, side = sub >> 2, cq = sub & 3;
  const uint32_t ref_bytes = (uint32_t)a.D * 8u;

  // lane 0: the warp's 16 consecutive hit-list rows of unit wu -> its slice (four box copies on one barrier)
  auto issue = [&](const Unit& wu) {
    const int p = wu.pos0 + warp * kWarpItems;
    mbar_arrive_expect_tx(bar_qp, (uint32_t)(SM::sl_off + SM::sl_lg + kWarpItems * ref_bytes + SM::sl_meta));
    tma_load_4d(sl_off, &map_off, bar_qp, a.off_col + my_h * PP * 2, p, wu.n, wu.b);
    tma_load_4d(sl_lg, &map_lg, bar_qp, a.logit_col + my_h * PP, p, wu.n, wu.b);
    tma_load_4d(sl_ref, &map_ref, bar_qp, 0, p, wu.n, wu.b);
    tma_load_4d(sl_meta, &map_meta, bar_qp, 0, p, wu.n, wu.b);
  };
  if (lane == 0) issue(w);

  int loaded = w.plane;
  uint32_t win_phase = 0u;
  bool fresh = true;
  for (int u = u_beg; u < u_end; ++u) {
    // ---- P1 from the staged slice
    mbar_wait(bar_qp, (uint32_t)((u - u_beg) & 1));
    const int pos = w.pos0 + warp * kWarpItems + item_l;
    const bool ok = pos < w.limit;
    float off[PPL * 2], lg[PPL], ref[PPL * 2], ic;
    int qq;
    {
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
      const uint32_t ra = sl_ref + (uint32_t)item_l * ref_bytes;
#pragma unroll
      for (int i = 0; i < PPL; ++i)
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(ref[2 * i]), "=f"(ref[2 * i + 1])
                     : "r"(ra + (uint32_t)((p0 + i) % a.D) * 8u));
      asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(qq), "=f"(ic) : "r"(sl_meta + (uint32_t)item_l * 16u));
      if (!ok) {   // rows beyond the list hold other hits or uninitialised memory
        ic = 0.f;
#pragma unroll
        for (int i = 0; i < PPL; ++i) off[2 * i] = off[2 * i + 1] = ref[2 * i] = ref[2 * i + 1] = lg[i] = 0.f;
      }
    }
    {
      float aw[PPL];
      softmax_pair32<PPL>(lg, ok, ic, aw);
      float4 w4[PPL];
      uint32_t idx[PPL];
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        const float w_im = fmaf(ref[2 * i], (float)a.fW, off[2 * i] - 0.5f);
        const float h_im = fmaf(ref[2 * i + 1], (float)a.fH, off[2 * i + 1] - 0.5f);
        // the window is the whole plane plus a one-pixel zero halo (origin (-1, -1)): nothing is ever far
        make_desc32(ok, h_im, w_im, aw[i], a.fH, a.fW, -1, -1, a.WW, a.WH, w4[i], idx[i]);
      }
      store_descs32<PP>(sm_w, sm_idx, item_l, p0, w4, idx);
    }
    // the items this lane writes in P2: (grp + 4 m) for m = 2 side, 2 side + 1 -> their queries (-1: not a hit)
    int qm[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int item = grp + 4 * (2 * side + j);
      const int q_it = __shfl_sync(0xffffffffu, qq, item * 2);
      const bool ok_it = w.pos0 + warp * kWarpItems + item < w.limit;
      qm[j] = ok_it ? q_it : -1;
    }
    __syncwarp();   // descriptors visible to the whole warp; the slice has been consumed
    const Unit w_cur = w;
    if (u + 1 < u_end) {
      w = decode(u + 1);
      if (lane == 0) issue(w);
    }
    // ---- P2
    if (fresh) {
      mbar_wait(bar, win_phase);
      win_phase ^= 1u;
      fresh = false;
    }
    float4 res[4];
    const int n_live = max(0, min(kWarpItems, w_cur.limit - (w_cur.pos0 + warp * kWarpItems)));
    gather_warp32<PP, ROWB>(sm_w, sm_idx, sm_win + sub * 16u, (uint32_t)a.WW * 64u, grp, side, n_live, res);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (qm[j] >= 0) {
        float* dst = a.out + ((int64_t)w_cur.b * a.Nq + qm[j]) * C + my_h * 32 + my_s * 16 + cq * 4;
        const float4 o = side ? res[2 + j] : res[j];
        if (a.part)
          red_add4(dst, o);
        else
          st_stream4(dst, o);
      }
    }
    __syncwarp();
    if (u + 1 < u_end && w.plane != loaded) {   // uniform per CTA: every warp leaves the old plane first
      __syncthreads();
      if (tid == 0) load_plane(w.plane);
      loaded = w.plane;
      fresh = true;
    }
  }
}

template <int PP, int ROWB>
static int launch_img_win32_v(ImgWin32Args& a, const CUtensorMap* m, size_t smem, cudaStream_t s) {
  const char* fn = "ub_img_sample_win32_fwd";
  if (int rc = ensure_smem(img_sample_win32_kernel<PP, ROWB>, smem, fn)) return rc;
  for (int part = 0; part < 2; ++part) {   // later hits (~12 % of the pairs on the nuScenes rig) accumulate on top
    a.part = part;
    launch_pdl(img_sample_win32_kernel<PP, ROWB>, dim3(sm_count()), dim3(kImgThreads), smem, s, a, m[0], m[1], m[2], m[3], m[4]);
    if (int rc = check_launch(fn)) return rc;
  }
  return UB_OK;
}

}  // namespace ub

extern "C" int ub_hit_order(const uint8_t* mask, const float* ref_cam, const int* hit_idx, const int* hit_cnt,
                            const float* inv_cnt, int* q_dst, float* hit_ref, float* hit_meta, int B, int N, int Nq, int D,
                            ub_stream_t stream) {
  UB_REQUIRE(mask && ref_cam && hit_idx && hit_cnt && inv_cnt && q_dst && hit_ref && hit_meta, "ub_hit_order: null pointer");
  UB_REQUIRE_ALIGNED16(hit_meta);
  UB_REQUIRE(B > 0 && N > 0 && N <= 32 && Nq > 0 && D > 0 && D <= 8, "ub_hit_order: bad dimension (B=%d N=%d Nq=%d D=%d)", B, N,
             Nq, D);
  int blocks = (int)(((int64_t)B * N * Nq + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  hit_order_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mask, ref_cam, hit_idx, hit_cnt, inv_cnt, q_dst, hit_ref,
                                                             reinterpret_cast<float4*>(hit_meta), B, N, Nq, 2 * D);
  return check_launch("ub_hit_order");
}

extern "C" int ub_img_sample_win32_fwd(const float* planes32, const float* qp_hit, const float* hit_ref, const float* hit_meta,
                                       const int* hit_idx, const int* hit_cnt, float* out, int B, int N, int bev_h, int bev_w,
                                       int fH, int fW, int H, int Dh, int P, int D, int ld, int off_col, int logit_col,
                                       ub_stream_t stream) {
  const char* fn = "ub_img_sample_win32_fwd";
  UB_REQUIRE(planes32 && qp_hit && hit_ref && hit_meta && hit_idx && hit_cnt && out, "%s: null pointer", fn);
  UB_REQUIRE(B > 0 && N > 0 && N <= 32 && bev_h > 0 && bev_w > 0 && fH >= 2 && fW >= 2 && H > 0 && D > 0 && D <= 8,
             "%s: bad dimension (B=%d N=%d D=%d)", fn, B, N, D);
  UB_REQUIRE(P % D == 0, "%s: num_points %d must be a multiple of the %d Z-anchors", fn, P, D);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small", fn, ld);
  UB_REQUIRE_ALIGNED16(planes32);
  UB_REQUIRE_ALIGNED16(qp_hit);
  UB_REQUIRE_ALIGNED16(hit_ref);
  UB_REQUIRE_ALIGNED16(hit_meta);
  UB_REQUIRE_ALIGNED16(out);
  const int Nq = bev_h * bev_w;
  const int win_bytes = ((fW + 2) * (fH + 2) * 64 + 127) & ~127;
  if (Dh != 32 || (P != 4 && P != 8) || ld % 4 != 0 || off_col % 4 != 0 || logit_col % 4 != 0 || D % 2 != 0 ||
      fW + 2 > 256 || fH + 2 > 256 || (int64_t)(fH + 2) * (fW + 2) > 65535 || (int64_t)B * N * H * 2 > (1 << 20) ||
      2 * H > sm_count() || ImgSmem32<8>::total(win_bytes) > kSmemBudget) {
    set_error("%s: shape not covered by the window kernels (Dh=%d P=%d D=%d fH=%d fW=%d Nq=%d)", fn, Dh, P, D, fH, fW, Nq);
    return ub::unsupported();
  }
  ImgWin32Args a;
  a.hit_idx = hit_idx, a.hit_cnt = hit_cnt, a.out = out;
  a.B = B, a.N = N, a.Nq = Nq, a.fH = fH, a.fW = fW, a.H = H, a.D = D, a.off_col = off_col, a.logit_col = logit_col;
  a.WW = fW + 2, a.WH = fH + 2, a.part = 0;
  CUtensorMap m[5];
  const CUtensorMapDataType f32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    const uint64_t dims[4] = {16, (uint64_t)fW, (uint64_t)fH, (uint64_t)B * N * H * 2};
    const uint64_t str[3] = {64, (uint64_t)fW * 64, (uint64_t)fH * fW * 64};
    const uint32_t box[4] = {16, (uint32_t)a.WW, (uint32_t)a.WH, 1};
    if (int rc = make_tensor_map(&m[0], f32, 4, planes32, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)ld, (uint64_t)Nq, (uint64_t)N, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ld * 4, (uint64_t)Nq * ld * 4, (uint64_t)N * Nq * ld * 4};
    const uint32_t box_o[4] = {(uint32_t)(2 * P), kWarpItems, 1, 1}, box_l[4] = {(uint32_t)P, kWarpItems, 1, 1};
    if (int rc = make_tensor_map(&m[1], f32, 4, qp_hit, dims, str, box_o, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
    if (int rc = make_tensor_map(&m[2], f32, 4, qp_hit, dims, str, box_l, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)(2 * D), (uint64_t)Nq, (uint64_t)N, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)D * 8, (uint64_t)Nq * D * 8, (uint64_t)N * Nq * D * 8};
    const uint32_t box[4] = {(uint32_t)(2 * D), kWarpItems, 1, 1};
    if (int rc = make_tensor_map(&m[3], f32, 4, hit_ref, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  }
  {
    const uint64_t dims[4] = {4, (uint64_t)Nq, (uint64_t)N, (uint64_t)B};
    const uint64_t str[3] = {16, (uint64_t)Nq * 16, (uint64_t)N * Nq * 16};
    const uint32_t box[4] = {4, kWarpItems, 1, 1};
    if (int rc = make_tensor_map(&m[4], f32, 4, hit_meta, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  }
  const size_t smem = P == 8 ? ImgSmem32<8>::total(win_bytes) : ImgSmem32<4>::total(win_bytes);
  const cudaStream_t s = (cudaStream_t)stream;
  if (a.WW == 52)   // nuScenes 1600 x 928 / 32 -> 50 + 2
    return P == 8 ? launch_img_win32_v<8, 52 * 64>(a, m, smem, s) : launch_img_win32_v<4, 52 * 64>(a, m, smem, s);
  return P == 8 ? launch_img_win32_v<8, 0>(a, m, smem, s) : launch_img_win32_v<4, 0>(a, m, smem, s);
}
