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
// Work unit (BEV mode) = 16 x 16 BEV queries x one head.  512 threads:
//   phase 1  one thread per SAMPLE (query, point): offsets / logits come from a TMA-staged tile of the fused
//            offset|logit GEMM output; softmax across the P adjacent lanes; the sample becomes a descriptor in
//            shared memory: a 16-bit pixel index into the window and two half2 words
//            {w_top, w_bottom} x {left, right} (attention weight folded in).
//   phase 2  8 lanes per ITEM (query, head): lanes 0-3 own the left pixel, 4-7 the right one, 8 channels each;
//            per point 1 LDS.32 (weights) + 2 LDS.128 (top / bottom pixel pair) + 16 mixed-precision FMAs
//            (fma.rn.f32.f16: fp16 value x fp16 weight accumulated in fp32); the halves are combined with four
//            shuffles and written as one 128-byte row.
// Samples whose 2 x 2 footprint is not inside the staged window (offsets larger than the halo) take a slow,
// exact path straight from global memory (fp32 weights), so any offsets are handled; the halo is a tuning knob.
// Windows are double-buffered (the next unit's window streams in during the current unit), units are handed
// out by an atomic counter that the last CTA resets.
//
// Camera mode: a unit is 256 hits of one camera x one head; the whole (camera, head) plane plus a one-pixel zero
// halo is the window (loaded when the CTA's contiguous unit range crosses into a new plane); contributions are
// pre-scaled by 1/count and accumulated with red.global.add.v4.f32 into a zero-filled output.
#include <cuda_fp16.h>

#include "ub_tma.cuh"

namespace ub {

constexpr int kWinThreads = 512;
constexpr int kTQ = 16;                  // BEV tile: 16 x 16 queries
constexpr int kUnitItems = kTQ * kTQ;    // items (query, head) per unit
constexpr int kGroups = kWinThreads / 8; // 8-lane groups per CTA

static int g_bev_halo = 0;  // 0 = default (P + 1)

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

template <int PP>
__device__ __forceinline__ float softmax_pp(float logit, bool ok) {
  float mx = logit;
#pragma unroll
  for (int o = PP / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = ok ? __expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = PP / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  return ok ? __fdividef(e, sum) : 0.f;
}

// word index of a sample's {left, right} weight pair; the XOR spreads the four items of a warp over the banks
template <int PP>
__device__ __forceinline__ int w_word(int item, int p) {
  return (item * PP + (PP == 8 ? (p ^ ((item >> 1) & 1)) : p)) * 2;
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
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

__device__ __forceinline__ void red_add4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Phase 2 of one unit: group `grp` (8 lanes) reduces items grp, grp + 64, grp + 128, grp + 192, two at a time.
// s_w / s_idx: descriptor arrays (shared-space byte addresses); win: window base + sub * 16; row_b: bytes per
// window row.  emit(item, lane_acc) receives this lane's four output channels (item, cq * 8 + half * 4 ...).
template <int PP, typename Emit>
__device__ __forceinline__ void gather_unit(uint32_t s_w, uint32_t s_idx, uint32_t win, uint32_t row_b, int grp, int half,
                                            Emit emit) {
#pragma unroll 1
  for (int k = 0; k < kUnitItems / kGroups; k += 2) {
    const int item[2] = {grp + k * kGroups, grp + (k + 1) * kGroups};
    uint32_t ix[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (PP == 8) {
        const uint4 t = lds128(s_idx + item[j] * 16);
        ix[j][0] = t.x, ix[j][1] = t.y, ix[j][2] = t.z, ix[j][3] = t.w;
      } else {
        uint32_t a, b;
        asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(s_idx + item[j] * 8));
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
        w[j] = lds32(s_w + (uint32_t)(w_word<PP>(item[j], p) + half) * 4u);
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
      float4 o;
      {
        const float s0 = half ? acc[j][0] : acc[j][4], s1 = half ? acc[j][1] : acc[j][5];
        const float s2 = half ? acc[j][2] : acc[j][6], s3 = half ? acc[j][3] : acc[j][7];
        const float r0 = __shfl_xor_sync(0xffffffffu, s0, 4), r1 = __shfl_xor_sync(0xffffffffu, s1, 4);
        const float r2 = __shfl_xor_sync(0xffffffffu, s2, 4), r3 = __shfl_xor_sync(0xffffffffu, s3, 4);
        o.x = (half ? acc[j][4] : acc[j][0]) + r0;
        o.y = (half ? acc[j][5] : acc[j][1]) + r1;
        o.z = (half ? acc[j][6] : acc[j][2]) + r2;
        o.w = (half ? acc[j][7] : acc[j][3]) + r3;
      }
      emit(item[j], o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct BevWinArgs {
  const __half* value16;  // (B*H, fH, fW, 32): far path
  float* out;             // (B, Nq, H*32)
  int* counters;          // [0] next unit, [1] CTAs done
  int B, bev_h, bev_w, fH, fW, H;
  int tiles_x, tiles_y, n_units;
  int WW, WH, R;
  int off_col, logit_col;
  float sx, sy;
};

template <int PP>
struct BevSmem {
  static constexpr int n_samples = kUnitItems * PP;
  static constexpr int off_bytes = n_samples * 8, lg_bytes = n_samples * 4;
  static constexpr int w_bytes = n_samples * 8, idx_bytes = n_samples * 2;
  static size_t total(int win_bytes) { return (size_t)2 * win_bytes + off_bytes + lg_bytes + w_bytes + idx_bytes; }
};

template <int PP>
__global__ void __launch_bounds__(kWinThreads, 1)
    bev_sample_win_kernel(const BevWinArgs a, const __grid_constant__ CUtensorMap map_val,
                          const __grid_constant__ CUtensorMap map_off, const __grid_constant__ CUtensorMap map_lg) {
  using S = BevSmem<PP>;
  constexpr int SPT = S::n_samples / kWinThreads;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar[3];  // window buffers 0 / 1, offset|logit tile
  __shared__ int s_unit[2];

  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  unsigned char* p_off = smem + 2 * (size_t)win_bytes;
  unsigned char* p_lg = p_off + S::off_bytes;
  unsigned char* p_w = p_lg + S::lg_bytes;
  unsigned char* p_idx = p_w + S::w_bytes;
  const uint32_t sm_win = smem_u32(smem), sm_off = smem_u32(p_off), sm_lg = smem_u32(p_lg);
  const uint32_t sm_w = smem_u32(p_w), sm_idx = smem_u32(p_idx);
  const uint32_t bar_win0 = smem_u32(&s_bar[0]), bar_qp = smem_u32(&s_bar[2]);

  const int tid = threadIdx.x, lane = tid & 31;
  const int Nq = a.bev_h * a.bev_w, C = a.H * 32;
  const int n_tiles = a.tiles_x * a.tiles_y;
  const uint32_t qp_bytes = S::off_bytes + S::lg_bytes;
  const uint32_t box_bytes = (uint32_t)(a.WW * a.WH * 64);

  struct Unit {
    int b, h, tx0, ty0, wx0, wy0;
  };
  auto decode = [&](int u) {
    Unit w;
    w.h = u % a.H;
    const int t = (u / a.H) % n_tiles;
    w.b = u / (a.H * n_tiles);
    w.tx0 = (t % a.tiles_x) * kTQ, w.ty0 = (t / a.tiles_x) * kTQ;
    w.wx0 = (int)floorf(((float)w.tx0 + 0.5f) * a.sx - 0.5f) - a.R;
    w.wy0 = (int)floorf(((float)w.ty0 + 0.5f) * a.sy - 0.5f) - a.R;
    return w;
  };
  auto issue_window = [&](int u, int buf) {  // one thread
    const Unit w = decode(u);
    const uint32_t bar = bar_win0 + 8u * buf;
    mbar_arrive_expect_tx(bar, box_bytes);
    tma_load_4d(sm_win + (uint32_t)buf * win_bytes, &map_val, bar, 0, w.wx0, w.wy0, w.b * a.H + w.h);
  };
  auto issue_qproj = [&](int u) {  // one thread
    const Unit w = decode(u);
    mbar_arrive_expect_tx(bar_qp, qp_bytes);
    tma_load_4d(sm_off, &map_off, bar_qp, a.off_col + w.h * PP * 2, w.tx0, w.ty0, w.b);
    tma_load_4d(sm_lg, &map_lg, bar_qp, a.logit_col + w.h * PP, w.tx0, w.ty0, w.b);
  };

  if (tid == 0) {
    mbar_init(bar_win0, 1);
    mbar_init(bar_win0 + 8, 1);
    mbar_init(bar_qp, 1);
    mbar_init_fence();
    tma_prefetch_desc(&map_val);
    tma_prefetch_desc(&map_off);
    tma_prefetch_desc(&map_lg);
    const int u0 = atomicAdd(&a.counters[0], 1);
    const int u1 = atomicAdd(&a.counters[0], 1);
    s_unit[0] = u0, s_unit[1] = u1;
    if (u0 < a.n_units) issue_qproj(u0), issue_window(u0, 0);
    if (u1 < a.n_units) issue_window(u1, 1);
  }
  __syncthreads();

  const int grp = tid >> 3, sub = tid & 7, half = sub >> 2, cq = sub & 3;

  for (int it = 0;; ++it) {
    const int cur = it & 1;
    const int u = s_unit[cur];
    if (u >= a.n_units) break;
    int u_fetch = 0;
    if (tid == 0) u_fetch = atomicAdd(&a.counters[0], 1);  // consumed at the end of the iteration
    const Unit w = decode(u);

    // ---- phase 1
    mbar_wait(bar_qp, (uint32_t)(it & 1));
    float far_h[SPT], far_w[SPT], far_a[SPT];
    unsigned far_bits = 0;
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int s = r * kWinThreads + tid;
      const int p = s % PP, item = s / PP;
      const int qx = w.tx0 + (item & (kTQ - 1)), qy = w.ty0 + (item >> 4);
      const bool ok = qx < a.bev_w && qy < a.bev_h;
      float ox, oy, lg;
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(ox), "=f"(oy) : "r"(sm_off + (uint32_t)s * 8u));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lg) : "r"(sm_lg + (uint32_t)s * 4u));
      const float aw = softmax_pp<PP>(ok ? lg : 0.f, ok);
      // pixel = ((q + .5) / bev + off / f) * f - .5  ==  (q + .5) * (f / bev) + off - .5
      const float h_im = fmaf((float)qy + 0.5f, a.sy, oy - 0.5f), w_im = fmaf((float)qx + 0.5f, a.sx, ox - 0.5f);
      uint32_t wl = 0u, wr = 0u, idx = 0u;
      far_h[r] = h_im, far_w[r] = w_im, far_a[r] = aw;
      if (ok && h_im > -1.f && w_im > -1.f && h_im < (float)a.fH && w_im < (float)a.fW) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const float lh = h_im - hf, lw = w_im - wf;
        const int yy = (int)hf - w.wy0, xx = (int)wf - w.wx0;
        if (xx >= 0 && xx < a.WW - 1 && yy >= 0 && yy < a.WH - 1) {
          const float wt = aw * (1.f - lh), wb = aw * lh;
          wl = pack_h2(wt * (1.f - lw), wb * (1.f - lw));
          wr = pack_h2(wt * lw, wb * lw);
          idx = (uint32_t)(yy * a.WW + xx);
        } else {
          far_bits |= 1u << r;
        }
      }
      asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sm_w + (uint32_t)w_word<PP>(item, p) * 4u), "r"(wl), "r"(wr)
                   : "memory");
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(sm_idx + (uint32_t)s * 2u), "h"((unsigned short)idx) : "memory");
    }
    const int n_far = __syncthreads_count(far_bits != 0u);
    // the offset|logit tile is free again: stream in the next unit's
    if (tid == 0) {
      const int un = s_unit[cur ^ 1];
      if (un < a.n_units) issue_qproj(un);
    }

    // ---- slow path for samples outside the staged window (exact: fp32 weights, per-corner bounds checks)
    if (n_far > 0) {
#pragma unroll
      for (int k = 0; k < kUnitItems * 8 / kWinThreads; ++k) {
        const int e = k * kWinThreads + tid, item = e >> 3, c4 = e & 7;
        const int qx = w.tx0 + (item & (kTQ - 1)), qy = w.ty0 + (item >> 4);
        if (qx < a.bev_w && qy < a.bev_h)
          *reinterpret_cast<float4*>(a.out + ((int64_t)w.b * Nq + qy * a.bev_w + qx) * C + w.h * 32 + c4 * 4) =
              make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
      const __half* plane = a.value16 + (int64_t)(w.b * a.H + w.h) * a.fH * a.fW * 32;
#pragma unroll
      for (int r = 0; r < SPT; ++r) {
        unsigned m = __ballot_sync(0xffffffffu, (far_bits >> r) & 1u);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float h_im = __shfl_sync(0xffffffffu, far_h[r], src), w_im = __shfl_sync(0xffffffffu, far_w[r], src);
          const float aw = __shfl_sync(0xffffffffu, far_a[r], src);
          const int item = (r * kWinThreads + (tid & ~31) + src) / PP;
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
          const int qx = w.tx0 + (item & (kTQ - 1)), qy = w.ty0 + (item >> 4);
          if (lane < 8) red_add4(a.out + ((int64_t)w.b * Nq + qy * a.bev_w + qx) * C + w.h * 32 + c4 * 4, v);
        }
      }
    }

    // ---- phase 2
    mbar_wait(bar_win0 + 8u * cur, (uint32_t)((it >> 1) & 1));
    gather_unit<PP>(sm_w, sm_idx, sm_win + (uint32_t)cur * win_bytes + sub * 16u, (uint32_t)a.WW * 64u, grp, half,
                    [&](int item, const float4& o) {
                      const int qx = w.tx0 + (item & (kTQ - 1)), qy = w.ty0 + (item >> 4);
                      if (qx < a.bev_w && qy < a.bev_h) {
                        float* dst = a.out + ((int64_t)w.b * Nq + qy * a.bev_w + qx) * C + w.h * 32 + cq * 8 + half * 4;
                        if (n_far > 0)
                          red_add4(dst, o);
                        else
                          st_stream4(dst, o);
                      }
                    });
    __syncthreads();  // window buffer `cur` and the descriptors are free
    if (tid == 0) {
      s_unit[cur] = u_fetch;
      if (u_fetch < a.n_units) issue_window(u_fetch, cur);
    }
  }

  // the last CTA to leave re-arms the unit counter for the next launch
  if (tid == 0) {
    __threadfence();
    const int done = atomicAdd(&a.counters[1], 1);
    if (done == (int)gridDim.x - 1) {
      a.counters[0] = 0;
      a.counters[1] = 0;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Per-camera hit lists (ascending query index) from batch item 0's visibility bits (reference quirk,
// spatial_cross_attention_img.py:142) and 1 / max(1, #cameras that see (b, q)) (:209-212).
__global__ void __launch_bounds__(1024) build_hits_kernel(const uint8_t* __restrict__ mask, int* __restrict__ hit_idx,
                                                          int* __restrict__ hit_cnt, float* __restrict__ inv_cnt, int B,
                                                          int N, int Nq) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((int)blockIdx.x < N) {
    const int n = blockIdx.x;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int q0 = 0; q0 < Nq; q0 += 1024) {
      const int q = q0 + tid;
      const bool hit = q < Nq && mask[(int64_t)q * N + n] != 0;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      if (warp == 0) {
        int v = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += t;
        }
        s_warp[lane] = v;  // inclusive
      }
      __syncthreads();
      const int base = s_base + (warp ? s_warp[warp - 1] : 0);
      if (hit) hit_idx[(int64_t)n * Nq + base + __popc(bal & ((1u << lane) - 1u))] = q;
      __syncthreads();
      if (tid == 0) s_base += s_warp[31];
      __syncthreads();
    }
    if (tid == 0) hit_cnt[n] = s_base;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < (int64_t)B * Nq; i += (int64_t)gridDim.x * blockDim.x) {
    int c = 0;
    for (int n = 0; n < N; ++n) c += mask[i * N + n] != 0 ? 1 : 0;
    inv_cnt[i] = 1.f / (float)max(c, 1);
  }
}

// ---------------------------------------------------------------------------------------------------------
struct ImgWinArgs {
  const float* qproj;
  const float* ref_cam;   // (B, Nq, N, D, 2)
  const float* inv_cnt;   // (B, Nq)
  const int* hit_idx;     // (N, Nq)
  const int* hit_cnt;     // (N)
  float* out;             // (B, Nq, H*32), zero-filled
  int B, N, Nq, fH, fW, H, P, D, ld, off_col, logit_col;
  int WW, WH;
};

template <int PP>
__global__ void __launch_bounds__(kWinThreads, 1)
    img_sample_win_kernel(const ImgWinArgs a, const __grid_constant__ CUtensorMap map_val) {
  constexpr int n_samples = kUnitItems * PP, SPT = n_samples / kWinThreads;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_q[kUnitItems];

  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  unsigned char* p_w = smem + win_bytes;
  unsigned char* p_idx = p_w + n_samples * 8;
  const uint32_t sm_win = smem_u32(smem), sm_w = smem_u32(p_w), sm_idx = smem_u32(p_idx), bar = smem_u32(&s_bar);
  const int tid = threadIdx.x;
  const int C = a.H * 32;

  // units: (b, camera, head, chunk of 256 hits), contiguous range per CTA
  int chunks_tot = 0;
  for (int n = 0; n < a.N; ++n) chunks_tot += (a.hit_cnt[n] + kUnitItems - 1) / kUnitItems;
  const int per_b = chunks_tot * a.H, total = per_b * a.B;
  const int per = total / gridDim.x, rem = total % gridDim.x;
  int u = blockIdx.x * per + min((int)blockIdx.x, rem);
  const int u_end = u + per + ((int)blockIdx.x < rem ? 1 : 0);

  struct Unit {
    int b, n, h, chunk, cnt;
  };
  auto decode = [&](int uu) {
    Unit w;
    w.b = uu / per_b;
    int r = uu % per_b;
    w.n = 0, w.cnt = 0, w.h = 0, w.chunk = 0;
    for (int n = 0; n < a.N; ++n) {
      const int cnt = a.hit_cnt[n], ch = (cnt + kUnitItems - 1) / kUnitItems;
      if (r < ch * a.H) {
        w.n = n, w.cnt = cnt, w.h = r / ch, w.chunk = r % ch;
        break;
      }
      r -= ch * a.H;
    }
    return w;
  };

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init_fence();
    tma_prefetch_desc(&map_val);
  }
  __syncthreads();
  if (u >= u_end) return;

  float ox[SPT], oy[SPT], lg[SPT], rx[SPT], ry[SPT], ic[SPT];
  int qq[SPT];
  auto prefetch = [&](const Unit& w) {
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int s = r * kWinThreads + tid;
      const int p = s % PP, item = s / PP;
      const int ord = w.chunk * kUnitItems + item;
      ox[r] = 0.f, oy[r] = 0.f, lg[r] = 0.f, rx[r] = 0.f, ry[r] = 0.f, ic[r] = 0.f, qq[r] = -1;
      if (ord < w.cnt && p < a.P) {
        const int q = __ldg(a.hit_idx + (int64_t)w.n * a.Nq + ord);
        const int64_t bq = (int64_t)w.b * a.Nq + q;
        const float* rowp = a.qproj + bq * a.ld;
        const float2 t = ld_stream2(rowp + a.off_col + (w.h * a.P + p) * 2);
        ox[r] = t.x, oy[r] = t.y;
        lg[r] = ld_stream1(rowp + a.logit_col + w.h * a.P + p);
        const float2 rc = __ldg(reinterpret_cast<const float2*>(a.ref_cam) + (bq * a.N + w.n) * a.D + (p % a.D));
        rx[r] = rc.x, ry[r] = rc.y;
        ic[r] = __ldg(a.inv_cnt + bq);
        qq[r] = q;
      }
    }
  };

  const int grp = tid >> 3, sub = tid & 7, half = sub >> 2, cq = sub & 3;
  Unit w = decode(u);
  prefetch(w);
  int loaded = -1, n_loads = 0;

  for (; u < u_end; ++u) {
    const int plane = (w.b * a.N + w.n) * a.H + w.h;
    const bool reload = plane != loaded;
    if (reload && tid == 0) {  // every thread passed the barrier that ended the previous phase 2
      mbar_arrive_expect_tx(bar, (uint32_t)(a.WW * a.WH * 64));
      tma_load_4d(sm_win, &map_val, bar, 0, -1, -1, plane);
    }
    loaded = plane;
    // ---- phase 1
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int s = r * kWinThreads + tid;
      const int p = s % PP, item = s / PP;
      const bool ok = qq[r] >= 0;
      const float aw = softmax_pp<PP>(lg[r], ok) * ic[r];
      const float h_im = fmaf(ry[r], (float)a.fH, oy[r] - 0.5f), w_im = fmaf(rx[r], (float)a.fW, ox[r] - 0.5f);
      uint32_t wl = 0u, wr = 0u, idx = 0u;
      if (ok && h_im > -1.f && w_im > -1.f && h_im < (float)a.fH && w_im < (float)a.fW) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const float lh = h_im - hf, lw = w_im - wf;
        const float wt = aw * (1.f - lh), wb = aw * lh;
        wl = pack_h2(wt * (1.f - lw), wb * (1.f - lw));
        wr = pack_h2(wt * lw, wb * lw);
        idx = (uint32_t)(((int)hf + 1) * a.WW + (int)wf + 1);  // window origin is pixel (-1, -1)
      }
      asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(sm_w + (uint32_t)w_word<PP>(item, p) * 4u), "r"(wl), "r"(wr)
                   : "memory");
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(sm_idx + (uint32_t)s * 2u), "h"((unsigned short)idx) : "memory");
      if (p == 0) s_q[item] = qq[r];
    }
    __syncthreads();
    const Unit w_cur = w;
    if (u + 1 < u_end) {
      w = decode(u + 1);
      prefetch(w);
    }
    // ---- phase 2
    if (reload) {
      mbar_wait(bar, (uint32_t)(n_loads & 1));
      ++n_loads;
    }
    gather_unit<PP>(sm_w, sm_idx, sm_win + sub * 16u, (uint32_t)a.WW * 64u, grp, half, [&](int item, const float4& o) {
      const int q = s_q[item];
      if (q >= 0) red_add4(a.out + ((int64_t)w_cur.b * a.Nq + q) * C + w_cur.h * 32 + cq * 8 + half * 4, o);
    });
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side

static int* g_counters = nullptr;  // 64 slots x {next unit, CTAs done}
static unsigned g_slot = 0;
static int* counter_slot() {
  if (!g_counters) {
    if (cudaMalloc(&g_counters, 64 * 2 * sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(g_counters, 0, 64 * 2 * sizeof(int));
  }
  return g_counters + 2 * (g_slot++ % 64);
}

template <typename K>
static int set_smem(K kernel, size_t smem, const char* fn) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("%s: cannot reserve %zu bytes of shared memory", fn, smem);
    cudaGetLastError();
    return UB_ECUDA;
  }
  return UB_OK;
}

constexpr size_t kSmemBudget = 220 * 1024;

template <int PP>
static int launch_bev_win(BevWinArgs& a, const void* value16, const float* qproj, int ld, cudaStream_t s) {
  const char* fn = "ub_bev_sample_win_fwd";
  // shrink the halo until two windows fit (samples beyond it stay exact through the slow path)
  auto smem_for = [&]() { return BevSmem<PP>::total((a.WW * a.WH * 64 + 127) & ~127); };
  while (smem_for() > kSmemBudget && a.R > 1) --a.R, a.WW -= 2, a.WH -= 2;
  const size_t smem = smem_for();
  if (smem > kSmemBudget) {
    set_error("%s: window %d x %d needs %zu bytes of shared memory", fn, a.WW, a.WH, smem);
    return UB_EUNSUPPORTED;
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
    const uint32_t box_o[4] = {2 * PP, kTQ, kTQ, 1}, box_l[4] = {PP, kTQ, kTQ, 1};
    if (int rc = make_tensor_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, qproj, dims, str, box_o,
                                 CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
    if (int rc = make_tensor_map(&ml, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, qproj, dims, str, box_l,
                                 CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  }
  static size_t configured = 0;
  if (smem > configured) {
    if (int rc = set_smem(bev_sample_win_kernel<PP>, smem, fn)) return rc;
    configured = smem;
  }
  a.counters = counter_slot();
  if (!a.counters) {
    set_error("%s: cannot allocate the unit counters", fn);
    return UB_ECUDA;
  }
  const int grid = a.n_units < kNumSMs ? a.n_units : kNumSMs;
  bev_sample_win_kernel<PP><<<grid, kWinThreads, smem, s>>>(a, mv, mo, ml);
  return check_launch(fn);
}

template <int PP>
static int launch_img_win(ImgWinArgs& a, const void* value16, cudaStream_t s) {
  const char* fn = "ub_img_sample_win_fwd";
  const int win_bytes = (a.WW * a.WH * 64 + 127) & ~127;
  const size_t smem = (size_t)win_bytes + kUnitItems * PP * 10;
  if (smem > kSmemBudget) {
    set_error("%s: plane %d x %d needs %zu bytes of shared memory", fn, a.fH, a.fW, smem);
    return UB_EUNSUPPORTED;
  }
  CUtensorMap mv;
  const uint64_t dims[4] = {32, (uint64_t)a.fW, (uint64_t)a.fH, (uint64_t)a.B * a.N * a.H};
  const uint64_t str[3] = {64, (uint64_t)a.fW * 64, (uint64_t)a.fH * a.fW * 64};
  const uint32_t box[4] = {32, (uint32_t)a.WW, (uint32_t)a.WH, 1};
  if (int rc = make_tensor_map(&mv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, value16, dims, str, box,
                               CU_TENSOR_MAP_SWIZZLE_NONE))
    return rc;
  static size_t configured = 0;
  if (smem > configured) {
    if (int rc = set_smem(img_sample_win_kernel<PP>, smem, fn)) return rc;
    configured = smem;
  }
  if (cudaMemsetAsync(a.out, 0, (size_t)a.B * a.Nq * a.H * 32 * sizeof(float), s) != cudaSuccess) {
    set_error("%s: cudaMemsetAsync failed", fn);
    return UB_ECUDA;
  }
  img_sample_win_kernel<PP><<<kNumSMs, kWinThreads, smem, s>>>(a, mv);
  return check_launch(fn);
}

}  // namespace ub

using namespace ub;

extern "C" int ub_set_window_halo(int halo) {
  UB_REQUIRE(halo >= 0 && halo <= 64, "ub_set_window_halo: halo must be in 0..64 (0 = default, points + 1)");
  g_bev_halo = halo;
  return UB_OK;
}

extern "C" int ub_value_to_half(const float* value, void* value16, int G, int Nv, int H, int Dh, ub_stream_t stream) {
  UB_REQUIRE(value && value16, "ub_value_to_half: null pointer");
  UB_REQUIRE(G > 0 && Nv > 0 && H > 0 && Dh > 0 && Dh % 8 == 0, "ub_value_to_half: need positive dims and Dh %% 8 == 0");
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(value16);
  const int64_t total = (int64_t)G * H * Nv * (Dh / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  value_to_half_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(value, reinterpret_cast<uint4*>(value16), G, Nv, H, Dh);
  return check_launch("ub_value_to_half");
}

extern "C" int ub_bev_sample_win_fwd(const void* value16, const float* qproj, float* out, int B, int bev_h, int bev_w,
                                     int fH, int fW, int H, int Dh, int P, int ld, int off_col, int logit_col,
                                     ub_stream_t stream) {
  const char* fn = "ub_bev_sample_win_fwd";
  UB_REQUIRE(value16 && qproj && out, "%s: null pointer", fn);
  UB_REQUIRE(B > 0 && bev_h > 0 && bev_w > 0 && fH >= 2 && fW >= 2 && H > 0, "%s: non-positive dimension", fn);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small", fn, ld);
  UB_REQUIRE_ALIGNED16(value16);
  UB_REQUIRE_ALIGNED16(qproj);
  UB_REQUIRE_ALIGNED16(out);
  if (Dh != 32 || (P != 4 && P != 8) || ld % 4 != 0 || off_col % 4 != 0 || logit_col % 4 != 0 ||
      (int64_t)B * H > (1 << 20)) {
    set_error("%s: shape not covered by the window kernels (Dh=%d P=%d ld=%d)", fn, Dh, P, ld);
    return UB_EUNSUPPORTED;
  }
  BevWinArgs a;
  a.value16 = reinterpret_cast<const __half*>(value16), a.out = out, a.counters = nullptr;
  a.B = B, a.bev_h = bev_h, a.bev_w = bev_w, a.fH = fH, a.fW = fW, a.H = H;
  a.tiles_x = (bev_w + kTQ - 1) / kTQ, a.tiles_y = (bev_h + kTQ - 1) / kTQ;
  a.n_units = B * H * a.tiles_x * a.tiles_y;
  a.sx = (float)fW / (float)bev_w, a.sy = (float)fH / (float)bev_h;
  a.R = g_bev_halo > 0 ? g_bev_halo : P + 1;
  a.WW = (int)ceilf((kTQ - 1) * a.sx) + 2 * a.R + 3;
  a.WH = (int)ceilf((kTQ - 1) * a.sy) + 2 * a.R + 3;
  a.off_col = off_col, a.logit_col = logit_col;
  if (a.WW > 256 || a.WH > 256) {
    set_error("%s: window %d x %d exceeds the TMA box limit", fn, a.WW, a.WH);
    return UB_EUNSUPPORTED;
  }
  return P == 8 ? launch_bev_win<8>(a, value16, qproj, ld, (cudaStream_t)stream)
                : launch_bev_win<4>(a, value16, qproj, ld, (cudaStream_t)stream);
}

extern "C" int ub_build_hits(const uint8_t* mask, int* hit_idx, int* hit_cnt, float* inv_cnt, int B, int N, int Nq,
                             ub_stream_t stream) {
  UB_REQUIRE(mask && hit_idx && hit_cnt && inv_cnt, "ub_build_hits: null pointer");
  UB_REQUIRE(B > 0 && N > 0 && N <= 32 && Nq > 0, "ub_build_hits: bad dimension (B=%d N=%d Nq=%d)", B, N, Nq);
  int blocks = (int)(((int64_t)B * Nq + 1023) / 1024);
  if (blocks < N) blocks = N;
  if (blocks > kNumSMs) blocks = kNumSMs;
  build_hits_kernel<<<blocks, 1024, 0, (cudaStream_t)stream>>>(mask, hit_idx, hit_cnt, inv_cnt, B, N, Nq);
  return check_launch("ub_build_hits");
}

extern "C" int ub_img_sample_win_fwd(const void* value16, const float* qproj, const float* ref_cam, const int* hit_idx,
                                     const int* hit_cnt, const float* inv_cnt, float* out, int B, int N, int bev_h,
                                     int bev_w, int fH, int fW, int H, int Dh, int P, int D, int ld, int off_col,
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
  if (Dh != 32 || (P != 4 && P != 8) || ld % 2 != 0 || off_col % 2 != 0 || fW + 2 > 256 || fH + 2 > 256 ||
      (int64_t)(fH + 2) * (fW + 2) > 65535 || (int64_t)B * N * H > (1 << 20)) {
    set_error("%s: shape not covered by the window kernels (Dh=%d P=%d fH=%d fW=%d)", fn, Dh, P, fH, fW);
    return UB_EUNSUPPORTED;
  }
  ImgWinArgs a;
  a.qproj = qproj, a.ref_cam = ref_cam, a.inv_cnt = inv_cnt, a.hit_idx = hit_idx, a.hit_cnt = hit_cnt, a.out = out;
  a.B = B, a.N = N, a.Nq = bev_h * bev_w, a.fH = fH, a.fW = fW, a.H = H, a.P = P, a.D = D;
  a.ld = ld, a.off_col = off_col, a.logit_col = logit_col;
  a.WW = fW + 2, a.WH = fH + 2;
  return P == 8 ? launch_img_win<8>(a, value16, (cudaStream_t)stream) : launch_img_win<4>(a, value16, (cudaStream_t)stream);
}
