// Fused per-modality deformable sampling for the UniBEV BEV encoder (one feature level).
//
//   ub_bev_sample_fwd  : BEV self-attention / LiDAR cross-attention sampling              [R4]
//   ub_img_sample_fwd  : camera cross-attention sampling, summed over cameras / count     [R3]
//
// The kernels never see a sampling_locations / attention_weights tensor: they read the raw outputs of
// the sampling_offsets / attention_weights linears (one fused GEMM), build the reference point in-kernel,
// normalise the offsets, run the softmax over the P logits, gather bilinearly and reduce.
//
// Persistent CTAs (a multiple of the 148 SMs), 256 threads.  A work unit is a tile of 64 BEV queries x HC
// heads; every CTA walks a contiguous range of units (tile fastest, so consecutive units are neighbouring
// tiles of the same heads and share their halo in L1).  Per unit:
//   phase 1  one thread per SAMPLE (query, head, point): its offset pair and logit were prefetched into
//            registers while the previous unit was being gathered (HBM latency hidden behind phase 2);
//            softmax across the P adjacent lanes with shuffles, location -> a sample descriptor in shared
//            memory: four corner weights (attention weight folded in, zero where the reference pads with
//            zeros) and the index of a 2x2 pixel block that always lies inside the map, so phase 2 loads
//            unconditionally.  Samples that miss the map are compacted to the end of their item's list.
//            Scalar work is done once per sample instead of once per channel lane.
//   phase 2  one group of LPG = Dh/4 lanes per ITEM (query, head), four channels per lane, two items
//            interleaved for memory-level parallelism: per point one broadcast LDS.128 (weights), a quarter
//            LDS.128 (indices), four LDG.128 corner fetches (each a fully used 16*LPG-byte segment: 128 B
//            at Dh = 32), 16 FFMAs.  Items are ordered head-major so all groups of the CTA gather from one
//            head's 128-byte column of a compact patch of the value map at a time (L1-resident).
// Descriptor slots are XOR-swizzled so the groups of a warp hit different 16-byte bank groups.
// Query-side streams (offsets/logits in, output out) bypass L1 (no-allocate).
#include "ub_common.cuh"

namespace ub {

constexpr int kThreads = 256;
constexpr int kTileQ = 64;  // BEV queries per tile (tile_w x tile_h, tile_w a power of two <= 64)

struct Tuning {
  int tile_w_log2 = 3;  // 8 x 8
  int ctas_per_sm = 0;  // 0 = kernel default
};
static Tuning g_bev_tuning, g_img_tuning;

// Sample descriptor, produced once per (query, head, point) in phase 1:
//   weights {w00, w01, w10, w11} = attention weight x bilinear corner weights of the 2 x 2 pixel block whose
//   top-left pixel is `index`; the block is shifted to lie inside the map ([0, fH-2] x [0, fW-2]) and the weights
//   of the corners the reference treats as zero padding are 0.
//   index < 0: the sample misses the map entirely (mmcv: h_im <= -1 || w_im <= -1 || h_im >= H || w_im >= W).
struct SampleDesc {
  float4 w;
  int index;
};

__device__ __forceinline__ SampleDesc make_desc(float h_im, float w_im, float aw, int fH, int fW) {
  SampleDesc d;
  d.w = make_float4(0.f, 0.f, 0.f, 0.f);
  d.index = -1;
  if (h_im > -1.f && w_im > -1.f && h_im < (float)fH && w_im < (float)fW) {
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h0 = (int)hf, w0 = (int)wf;
    const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
    float wt = hh, wb = lh, wl = hw, wr = lw;
    int yb = h0, xb = w0;
    if (h0 < 0) wt = lh, wb = 0.f, yb = 0;                 // only pixel row 0 (the reference's bottom corners)
    else if (h0 > fH - 2) wt = 0.f, wb = hh, yb = fH - 2;  // only pixel row fH-1 (the reference's top corners)
    if (w0 < 0) wl = lw, wr = 0.f, xb = 0;
    else if (w0 > fW - 2) wl = 0.f, wr = hw, xb = fW - 2;
    d.w = make_float4(aw * (wt * wl), aw * (wt * wr), aw * (wb * wl), aw * (wb * wr));
    d.index = yb * fW + xb;
  }
  return d;
}

// One sampling point of one item: four unconditional corner fetches -> acc += sum_k w_k * v_k.
// ROW > 0: floats per value token known at compile time (immediate load offsets).
template <int ROW>
__device__ __forceinline__ void gather_point(float4& acc, const float4 w, int index, const float* __restrict__ vbase,
                                             int row_rt, int fW) {
  const int row = ROW > 0 ? ROW : row_rt;
  const float* p0 = vbase + (int64_t)index * row;
  const float* p1 = vbase + (int64_t)(index + fW) * row;
  const float4 v1 = ldg4(p0), v2 = ldg4(p0 + row), v3 = ldg4(p1), v4 = ldg4(p1 + row);
  fma4(acc, w.x, v1);
  fma4(acc, w.y, v2);
  fma4(acc, w.z, v3);
  fma4(acc, w.w, v4);
}

template <int PP>
__device__ __forceinline__ int w_slot(int item, int p) {
  return item * PP + (p ^ (item & (PP - 1)));
}
template <int PP>
__device__ __forceinline__ int i_slot(int item, int p) {
  constexpr int Q = PP >= 4 ? PP / 4 : 1;  // int4 chunks per item
  if (PP < 4) return item * PP + p;
  return (item * Q + ((p >> 2) ^ (item & (Q - 1)))) * 4 + (p & 3);
}

// softmax weight of this lane's logit across the PP adjacent lanes of its (query, head)
template <int PP>
__device__ __forceinline__ float group_softmax(float logit, bool ok) {
  float mx = logit;
#pragma unroll
  for (int o = PP / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = ok ? __expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = PP / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  return ok ? __fdividef(e, sum) : 0.f;
}

// Writes one sample's descriptor: in-map samples of the item first (in point order), the rest after them with
// zero weights and the index of an in-map sample of the same item (so phase 2 only ever touches pixels the
// reference touches).  Returns the number of in-map samples of the item.
template <int PP>
__device__ __forceinline__ int store_desc(float4* s_w, int* s_i, int item, int p, const SampleDesc& d, int lane32) {
  const int grp_base = lane32 & ~(PP - 1);
  const unsigned full = PP >= 32 ? 0xffffffffu : ((1u << PP) - 1u);
  const unsigned in_map = (__ballot_sync(0xffffffffu, d.index >= 0) >> grp_base) & full;
  const int first = in_map ? __ffs(in_map) - 1 : 0;
  const int fill = __shfl_sync(0xffffffffu, d.index, grp_base + first);
  const unsigned below = (1u << p) - 1u;
  const int n_in = __popc(in_map);
  const bool in = d.index >= 0;
  const int slot = in ? __popc(in_map & below) : n_in + __popc(~in_map & below);
  s_w[w_slot<PP>(item, slot)] = d.w;
  s_i[i_slot<PP>(item, slot)] = in ? d.index : max(fill, 0);
  return n_in;
}

struct SampleArgs {
  const float* value;
  const float* qproj;
  float* out;
  const float* ref_cam;   // camera mode only
  const uint8_t* mask;    // camera mode only
  int N, D;               // cameras, Z-anchors (camera mode)
  int B, bev_h, bev_w, fH, fW, H, P, ld, off_col, logit_col;
  int tile_w_log2, tiles_x, n_tiles, n_chunks, n_units;
  float sx, sy;           // fW / bev_w, fH / bev_h
  int vec2_ok;            // offsets readable as float2
};

// compile-time geometry of a work unit
template <int LPG, int PP>
struct Geo {
  static constexpr int Dh = LPG * 4;
  static constexpr int n_groups = kThreads / LPG;
  static constexpr int hc0 = 8 / PP > n_groups / 32 ? 8 / PP : n_groups / 32;
  static constexpr int HC = hc0 > 1 ? hc0 : 1;      // heads per unit: >= 2 samples / thread, >= 2 items / group
  static constexpr int n_items = kTileQ * HC, n_samples = n_items * PP;
  static constexpr int SPT = n_samples / kThreads;  // samples per thread in phase 1
  static constexpr int IPG = n_items / n_groups;    // items per group in phase 2
  static constexpr size_t smem = (size_t)n_samples * (sizeof(float4) + sizeof(int)) + n_items * sizeof(int);
  static_assert(n_samples % kThreads == 0 && n_items % n_groups == 0 && IPG % 2 == 0, "unit geometry");
};

struct Unit {
  int b, h0, tx0, ty0, tile, chunk;
};
template <int HC>
__device__ __forceinline__ Unit decode_unit(const SampleArgs& a, int u) {
  Unit w;
  w.tile = u % a.n_tiles;
  const int rest = u / a.n_tiles;
  w.chunk = rest % a.n_chunks;
  w.h0 = w.chunk * HC;
  w.b = rest / a.n_chunks;
  w.tx0 = (w.tile % a.tiles_x) << a.tile_w_log2;
  w.ty0 = (w.tile / a.tiles_x) * (kTileQ >> a.tile_w_log2);
  return w;
}
// unit u -> unit u + 1 without divisions (tile fastest, then head chunk, then batch item)
template <int HC>
__device__ __forceinline__ void next_unit(const SampleArgs& a, Unit& w) {
  ++w.tile;
  w.tx0 += 1 << a.tile_w_log2;
  if (w.tx0 >= (a.tiles_x << a.tile_w_log2)) w.tx0 = 0, w.ty0 += kTileQ >> a.tile_w_log2;
  if (w.tile == a.n_tiles) {
    w.tile = 0, w.tx0 = 0, w.ty0 = 0, ++w.chunk, w.h0 += HC;
    if (w.chunk == a.n_chunks) w.chunk = 0, w.h0 = 0, ++w.b;
  }
}

// NI items of one group, interleaved point by point.  COUNTED: n_pts[j] in-map samples sit at the front of item
// j's list; chunks of four points beyond them are skipped.
template <int PP, int ROW, int NI, bool COUNTED>
__device__ __forceinline__ void gather_items(float4 (&acc)[NI], const float4* __restrict__ s_w,
                                             const int* __restrict__ s_i, const int (&item)[NI],
                                             const float* const (&vbase)[NI], const int (&n_pts)[NI], int row, int fW) {
  if (PP >= 4) {
#pragma unroll
    for (int p4 = 0; p4 < PP / 4; ++p4) {
      int ii[NI][4];
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int4 t = *reinterpret_cast<const int4*>(s_i + i_slot<PP>(item[j], p4 * 4));
        ii[j][0] = t.x, ii[j][1] = t.y, ii[j][2] = t.z, ii[j][3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          if (!COUNTED || p4 * 4 < n_pts[j])
            gather_point<ROW>(acc[j], s_w[w_slot<PP>(item[j], p4 * 4 + k)], ii[j][k], vbase[j], row, fW);
        }
      }
    }
  } else {
#pragma unroll
    for (int p = 0; p < PP; ++p) {
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        if (!COUNTED || p < n_pts[j])
          gather_point<ROW>(acc[j], s_w[w_slot<PP>(item[j], p)], s_i[i_slot<PP>(item[j], p)], vbase[j], row, fW);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// BEV-grid mode: reference point = cell centre, one value map (B, fH*fW, H*Dh).
// HF > 0: number of heads known at compile time.
template <int LPG, int PP, int HF, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) bev_sample_kernel(const SampleArgs a) {
  using G = Geo<LPG, PP>;
  constexpr int Dh = G::Dh, HC = G::HC, n_groups = G::n_groups, SPT = G::SPT, IPG = G::IPG;
  constexpr int ROW = HF * Dh;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_w = reinterpret_cast<float4*>(smem_raw);
  int* s_i = reinterpret_cast<int*>(smem_raw + (size_t)G::n_samples * sizeof(float4));
  int* s_cnt = s_i + G::n_samples;  // in-map samples per item

  const int tid = threadIdx.x;
  const int tw_mask = (1 << a.tile_w_log2) - 1;
  const int H = HF > 0 ? HF : a.H;
  const int Nq = a.bev_h * a.bev_w, row = H * Dh;
  const int lane = tid % LPG, group = tid / LPG;

  const int per = a.n_units / gridDim.x, rem = a.n_units % gridDim.x;
  int u = blockIdx.x * per + min((int)blockIdx.x, rem);
  const int u_end = u + per + ((int)blockIdx.x < rem ? 1 : 0);
  if (u >= u_end) return;

  float ox[SPT], oy[SPT], lg[SPT];
  auto prefetch = [&](const Unit& w) {
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int s = r * kThreads + tid;
      const int p = s % PP, item = s / PP;
      const int ql = item % kTileQ, hl = item / kTileQ;
      const int qx = w.tx0 + (ql & tw_mask), qy = w.ty0 + (ql >> a.tile_w_log2), h = w.h0 + hl;
      ox[r] = 0.f, oy[r] = 0.f, lg[r] = -INFINITY;
      if (qx < a.bev_w && qy < a.bev_h && h < H && p < a.P) {
        const float* rowp = a.qproj + ((int64_t)w.b * Nq + qy * a.bev_w + qx) * a.ld;
        const float* op = rowp + a.off_col + (h * a.P + p) * 2;
        if (a.vec2_ok) {
          const float2 t = ld_stream2(op);
          ox[r] = t.x, oy[r] = t.y;
        } else {
          ox[r] = ld_stream1(op), oy[r] = ld_stream1(op + 1);
        }
        lg[r] = ld_stream1(rowp + a.logit_col + h * a.P + p);
      }
    }
  };
  Unit w = decode_unit<HC>(a, u), w_next = w;
  prefetch(w);

  for (; u < u_end; ++u, w = w_next) {
    // ---- phase 1: descriptors from the prefetched registers
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int s = r * kThreads + tid;
      const int p = s % PP, item = s / PP;
      const int ql = item % kTileQ, hl = item / kTileQ;
      const int qx = w.tx0 + (ql & tw_mask), qy = w.ty0 + (ql >> a.tile_w_log2), h = w.h0 + hl;
      const bool ok = qx < a.bev_w && qy < a.bev_h && h < H && p < a.P;
      const float aw = group_softmax<PP>(lg[r], ok);
      SampleDesc d;
      d.w = make_float4(0.f, 0.f, 0.f, 0.f), d.index = -1;
      // pixel = ((q + .5) / bev + off / f) * f - .5  ==  (q + .5) * (f / bev) + off - .5
      if (ok)
        d = make_desc(fmaf((float)qy + 0.5f, a.sy, oy[r] - 0.5f), fmaf((float)qx + 0.5f, a.sx, ox[r] - 0.5f), aw, a.fH,
                      a.fW);
      const int n_in = store_desc<PP>(s_w, s_i, item, p, d, tid & 31);
      if (p == 0) s_cnt[item] = n_in;
    }
    __syncthreads();
    next_unit<HC>(a, w_next);
    if (u + 1 < u_end) prefetch(w_next);

    // ---- phase 2: gather + reduce, two items at a time
    const float* vb = a.value + (int64_t)w.b * a.fH * a.fW * row + lane * 4;
#pragma unroll
    for (int k = 0; k < IPG; k += 2) {
      const int item[2] = {group + k * n_groups, group + (k + 1) * n_groups};
      const int hl[2] = {item[0] / kTileQ, item[1] / kTileQ};
      const float* const vbase[2] = {vb + (w.h0 + hl[0]) * Dh, vb + (w.h0 + hl[1]) * Dh};
      const int n_pts[2] = {s_cnt[item[0]], s_cnt[item[1]]};
      float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
      if (n_pts[0] > 0 && n_pts[1] > 0)  // the common case: every slot holds an in-bounds index, no branches inside
        gather_items<PP, ROW, 2, false>(acc, s_w, s_i, item, vbase, n_pts, row, a.fW);
      else
        gather_items<PP, ROW, 2, true>(acc, s_w, s_i, item, vbase, n_pts, row, a.fW);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ql = item[j] % kTileQ;
        const int qx = w.tx0 + (ql & tw_mask), qy = w.ty0 + (ql >> a.tile_w_log2), h = w.h0 + hl[j];
        if (qx < a.bev_w && qy < a.bev_h && h < H)
          st_stream4(a.out + ((int64_t)w.b * Nq + qy * a.bev_w + qx) * row + h * Dh + lane * 4, acc[j]);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Camera mode: value (B, N, fH*fW, H*Dh); reference points / visibility from ub_project_points.
// A camera contributes to (b, q) iff batch item 0 sees q in it (reference quirk, sca_img:142); the sum over
// cameras is divided by max(1, #cameras that see (b, q)) (sca_img:209-212).
// Rounds: round r handles, for every query of the tile, its r-th contributing camera (most queries have one,
// frustum overlaps two), so a round is dense in queries whatever the camera layout.  Samples that miss the
// image are compacted away in phase 1, phase 2 skips them four at a time.
// Every group keeps its items' accumulators in registers across rounds (same item -> same group).
__device__ __forceinline__ int rth_camera(unsigned hit, int r) {
  for (int i = 0; i < r; ++i) hit &= hit - 1;
  return hit ? __ffs(hit) - 1 : -1;
}
// the first four contributing cameras of a query, one signed byte each (-1 = none)
__device__ __forceinline__ unsigned pack_cameras(unsigned hit) {
  unsigned packed = 0u;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = hit ? __ffs(hit) - 1 : -1;
    hit &= hit - 1;
    packed |= ((unsigned)n & 0xffu) << (8 * r);
  }
  return packed;
}
__device__ __forceinline__ int camera_of(unsigned packed, unsigned hit, int r) {
  if (r < 4) return ((int)(packed << (24 - 8 * r))) >> 24;
  return rth_camera(hit, r);
}

template <int LPG, int PP, int HF, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) img_sample_kernel(const SampleArgs a) {
  using G = Geo<LPG, PP>;
  constexpr int Dh = G::Dh, HC = G::HC, n_groups = G::n_groups, SPT = G::SPT, IPG = G::IPG;
  constexpr int ROW = HF * Dh;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_w = reinterpret_cast<float4*>(smem_raw);
  int* s_i = reinterpret_cast<int*>(smem_raw + (size_t)G::n_samples * sizeof(float4));
  int* s_cnt = s_i + G::n_samples;       // in-map samples of the item in this round
  __shared__ unsigned s_hit[2][kTileQ];  // cameras (bit n) that contribute to the query   (double-buffered per unit)
  __shared__ unsigned s_cams[2][kTileQ]; // the same as a packed list (pack_cameras)
  __shared__ float s_count[2][kTileQ];   // divisor
  __shared__ int s_rounds[2];

  const int tid = threadIdx.x;
  const int tw_mask = (1 << a.tile_w_log2) - 1;
  const int H = HF > 0 ? HF : a.H;
  const int Nq = a.bev_h * a.bev_w, row = H * Dh;
  const int cam_stride = a.fH * a.fW * row;  // < 2^31 (checked on the host)
  const int lane = tid % LPG, group = tid / LPG;

  const int per = a.n_units / gridDim.x, rem = a.n_units % gridDim.x;
  int u = blockIdx.x * per + min((int)blockIdx.x, rem);
  const int u_end = u + per + ((int)blockIdx.x < rem ? 1 : 0);
  if (u >= u_end) return;

  auto load_tile_info = [&](const Unit& w, int buf) {  // threads 0..63; s_rounds[buf] was zeroed a barrier earlier
    if (tid < kTileQ) {
      const int qx = w.tx0 + (tid & tw_mask), qy = w.ty0 + (tid >> a.tile_w_log2);
      unsigned hit = 0u;
      int count = 0;
      if (qx < a.bev_w && qy < a.bev_h) {
        const int q = qy * a.bev_w + qx;
        for (int n = 0; n < a.N; ++n) {
          hit |= (a.mask[(int64_t)q * a.N + n] != 0 ? 1u : 0u) << n;
          count += a.mask[((int64_t)w.b * Nq + q) * a.N + n] != 0 ? 1 : 0;
        }
      }
      s_hit[buf][tid] = hit;
      s_cams[buf][tid] = pack_cameras(hit);
      s_count[buf][tid] = (float)max(count, 1);
      atomicMax(&s_rounds[buf], max(__popc(hit), 1));
    }
  };

  float ox[SPT], oy[SPT], lg[SPT], rx[SPT], ry[SPT];
  auto prefetch = [&](const Unit& w, int buf, int r_cam) {
#pragma unroll
    for (int r = 0; r < SPT; ++r) {
      const int s = r * kThreads + tid;
      const int p = s % PP, item = s / PP;
      const int ql = item % kTileQ, hl = item / kTileQ;
      const int qx = w.tx0 + (ql & tw_mask), qy = w.ty0 + (ql >> a.tile_w_log2), h = w.h0 + hl;
      const int n = camera_of(s_cams[buf][ql], s_hit[buf][ql], r_cam);
      ox[r] = 0.f, oy[r] = 0.f, lg[r] = -INFINITY, rx[r] = 0.f, ry[r] = 0.f;
      if (n >= 0 && h < H && p < a.P) {
        const int64_t bq = (int64_t)w.b * Nq + qy * a.bev_w + qx;
        const float* rowp = a.qproj + bq * a.ld;
        const float* op = rowp + a.off_col + (h * a.P + p) * 2;
        if (a.vec2_ok) {
          const float2 t = ld_stream2(op);
          ox[r] = t.x, oy[r] = t.y;
        } else {
          ox[r] = ld_stream1(op), oy[r] = ld_stream1(op + 1);
        }
        lg[r] = ld_stream1(rowp + a.logit_col + h * a.P + p);
        const float2 t = __ldg(reinterpret_cast<const float2*>(a.ref_cam) + (bq * a.N + n) * a.D + (p % a.D));
        rx[r] = t.x, ry[r] = t.y;
      }
    }
  };

  if (tid < 2) s_rounds[tid] = 0;
  __syncthreads();
  Unit w = decode_unit<HC>(a, u), w_next = w;
  next_unit<HC>(a, w_next);
  load_tile_info(w, 0);
  __syncthreads();
  prefetch(w, 0, 0);

  for (int it = 0; u < u_end; ++u, ++it, w = w_next, next_unit<HC>(a, w_next)) {
    const int cur = it & 1;
    const int rounds = s_rounds[cur];
    float4 acc[IPG];
#pragma unroll
    for (int k = 0; k < IPG; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int rc = 0; rc < rounds; ++rc) {
      // ---- phase 1: descriptors of round rc from the prefetched registers
#pragma unroll
      for (int r = 0; r < SPT; ++r) {
        const int s = r * kThreads + tid;
        const int p = s % PP, item = s / PP;
        const int ql = item % kTileQ, hl = item / kTileQ;
        const bool ok = camera_of(s_cams[cur][ql], s_hit[cur][ql], rc) >= 0 && w.h0 + hl < H && p < a.P;
        const float aw = group_softmax<PP>(lg[r], ok);
        SampleDesc d;
        d.w = make_float4(0.f, 0.f, 0.f, 0.f), d.index = -1;
        if (ok)
          d = make_desc(fmaf(ry[r], (float)a.fH, oy[r] - 0.5f), fmaf(rx[r], (float)a.fW, ox[r] - 0.5f), aw, a.fH, a.fW);
        const int n_in = store_desc<PP>(s_w, s_i, item, p, d, tid & 31);
        if (p == 0) s_cnt[item] = n_in;
      }
      if (rc == 0 && u + 1 < u_end && tid == 0) s_rounds[cur ^ 1] = 0;
      __syncthreads();
      if (rc == 0 && u + 1 < u_end) load_tile_info(w_next, cur ^ 1);  // visible after the next barrier
      if (rc + 1 < rounds) prefetch(w, cur, rc + 1);

      // ---- phase 2
#pragma unroll
      for (int k = 0; k < IPG; k += 2) {
        const int item[2] = {group + k * n_groups, group + (k + 1) * n_groups};
        int n_pts[2];
        const float* vb2[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ql = item[j] % kTileQ, hl = item[j] / kTileQ;
          const int n = camera_of(s_cams[cur][ql], s_hit[cur][ql], rc);
          n_pts[j] = (n >= 0 && w.h0 + hl < H) ? s_cnt[item[j]] : 0;
          vb2[j] = a.value + (int64_t)(w.b * a.N + max(n, 0)) * cam_stride + ((w.h0 + hl) * Dh + lane * 4);
        }
        const float* const vbase[2] = {vb2[0], vb2[1]};
        float4 t[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        gather_items<PP, ROW, 2, true>(t, s_w, s_i, item, vbase, n_pts, row, a.fW);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          acc[k + j].x += t[j].x, acc[k + j].y += t[j].y, acc[k + j].z += t[j].z, acc[k + j].w += t[j].w;
      }
      __syncthreads();
      if (rc + 1 == rounds && u + 1 < u_end) prefetch(w_next, cur ^ 1, 0);  // its tile info landed two barriers ago
    }

#pragma unroll
    for (int k = 0; k < IPG; ++k) {
      const int item = group + k * n_groups;
      const int ql = item % kTileQ, hl = item / kTileQ;
      const int qx = w.tx0 + (ql & tw_mask), qy = w.ty0 + (ql >> a.tile_w_log2), h = w.h0 + hl;
      if (qx >= a.bev_w || qy >= a.bev_h || h >= H) continue;
      float4 o = acc[k];
      const float c = s_count[cur][ql];
      if (c != 1.f && s_hit[cur][ql]) o.x /= c, o.y /= c, o.z /= c, o.w /= c;
      st_stream4(a.out + ((int64_t)w.b * Nq + qy * a.bev_w + qx) * row + h * Dh + lane * 4, o);
    }
    // s_hit[cur] / s_count[cur] are next rewritten by load_tile_info(u + 2, cur), after the first barrier of unit u + 1
  }
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static int pad_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <typename K>
static int configure_smem(K kernel, size_t smem, const char* fn) {
  if (smem > 200 * 1024) {
    set_error("%s: this head dim / point count needs %zu bytes of shared memory", fn, smem);
    return UB_EINVAL;
  }
  return ensure_smem(kernel, smem, fn);
}

static int persistent_grid(int per_sm, int n_units, size_t smem) {
  const int by_smem = (int)((220 * 1024) / (smem + 2048));
  if (per_sm > by_smem) per_sm = by_smem > 0 ? by_smem : 1;
  const int grid = sm_count() * per_sm;
  return grid < n_units ? grid : n_units;
}

constexpr int kDefaultCtasPerSmBev = 4, kDefaultCtasPerSmImg = 3;  // measured best on B200 (tools/sweep_sampling.py)

template <int LPG, int PP, int HF, int MINB>
static int launch_bev_v(SampleArgs& a, cudaStream_t s) {
  using G = Geo<LPG, PP>;
  if (int rc = configure_smem(bev_sample_kernel<LPG, PP, HF, MINB>, G::smem, "ub_bev_sample_fwd")) return rc;
  a.n_chunks = (a.H + G::HC - 1) / G::HC;
  a.n_units = a.B * a.n_chunks * a.n_tiles;
  bev_sample_kernel<LPG, PP, HF, MINB><<<persistent_grid(MINB, a.n_units, G::smem), kThreads, G::smem, s>>>(a);
  return 0;
}
template <int LPG, int PP, int HF, int MINB>
static int launch_img_v(SampleArgs& a, cudaStream_t s) {
  using G = Geo<LPG, PP>;
  if (int rc = configure_smem(img_sample_kernel<LPG, PP, HF, MINB>, G::smem, "ub_img_sample_fwd")) return rc;
  a.n_chunks = (a.H + G::HC - 1) / G::HC;
  a.n_units = a.B * a.n_chunks * a.n_tiles;
  img_sample_kernel<LPG, PP, HF, MINB><<<persistent_grid(MINB, a.n_units, G::smem), kThreads, G::smem, s>>>(a);
  return 0;
}

// The shipped configurations (8 heads, 16 / 32 channels per head, 4 / 8 points) get the immediate-row-stride
// kernels, in a few occupancy variants selectable through ub_set_tuning; everything else one generic variant.
#define UB_LAUNCH_SHIPPED(FN, LPGv, PPv, DEFAULT)                                            \
  switch (t.ctas_per_sm > 0 ? t.ctas_per_sm : DEFAULT) {                           \
    case 3: return FN<LPGv, PPv, 8, 3>(a, s);                                                \
    default: return FN<LPGv, PPv, 8, 4>(a, s);                                               \
  }

template <int LPG, int PP>
static int launch_bev(SampleArgs& a, const Tuning& t, cudaStream_t s) {
  if constexpr ((LPG == 8 || LPG == 4) && (PP == 8 || PP == 4)) {
    if (a.H == 8) { UB_LAUNCH_SHIPPED(launch_bev_v, LPG, PP, kDefaultCtasPerSmBev) }
  }
  return launch_bev_v<LPG, PP, 0, 4>(a, s);
}
template <int LPG, int PP>
static int launch_img(SampleArgs& a, const Tuning& t, cudaStream_t s) {
  if constexpr ((LPG == 8 || LPG == 4) && PP == 8) {
    if (a.H == 8) { UB_LAUNCH_SHIPPED(launch_img_v, LPG, PP, kDefaultCtasPerSmImg) }
  }
  return launch_img_v<LPG, PP, 0, 4>(a, s);
}

}  // namespace ub

using namespace ub;

// which: 0 = ub_bev_sample_fwd, 1 = ub_img_sample_fwd.  tile_w: power of two <= 64 (tile = tile_w x 64/tile_w);
// ctas_per_sm: persistent CTAs per SM (0 = default).
extern "C" int ub_set_tuning(int which, int tile_w, int ctas_per_sm) {
  UB_REQUIRE(pow2(tile_w) && tile_w <= kTileQ, "ub_set_tuning: tile_w must be a power of two <= %d", kTileQ);
  UB_REQUIRE(ctas_per_sm >= 0 && ctas_per_sm <= 8, "ub_set_tuning: ctas_per_sm must be in 0..8");
  Tuning& t = which == 0 ? g_bev_tuning : g_img_tuning;
  int l = 0;
  while ((1 << l) < tile_w) ++l;
  t.tile_w_log2 = l;
  t.ctas_per_sm = ctas_per_sm;
  return UB_OK;
}

static int check_sample_args(const char* fn, int H, int Dh, int P, int ld, int off_col, int logit_col) {
  UB_REQUIRE(H > 0 && P > 0 && P <= 16, "%s: need H>0 and 0<P<=16 (got H=%d P=%d)", fn, H, P);
  UB_REQUIRE(Dh == 8 || Dh == 16 || Dh == 32 || Dh == 64, "%s: head dim %d unsupported (need 8, 16, 32 or 64)", fn, Dh);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small for off_col=%d logit_col=%d H=%d P=%d", fn, ld, off_col, logit_col, H,
             P);
  return UB_OK;
}

// Instantiated: head dims 8 / 16 / 32 / 64, up to 4 / 8 / 16 points.  Anything else is rejected (the plugin then
// takes the module path through ub_msda_fwd).
#define UB_DISPATCH_PP(FN, LPGv, PPv, ...)                       \
  switch (PPv) {                                                 \
    case 1: case 2: case 4: rc = FN<LPGv, 4>(__VA_ARGS__); break;  \
    case 8: rc = FN<LPGv, 8>(__VA_ARGS__); break;                \
    default: rc = FN<LPGv, 16>(__VA_ARGS__); break;              \
  }
#define UB_DISPATCH(FN, Dhv, PPv, ...)                                  \
  switch ((Dhv) / 4) {                                                  \
    case 2: UB_DISPATCH_PP(FN, 2, PPv, __VA_ARGS__); break;             \
    case 4: UB_DISPATCH_PP(FN, 4, PPv, __VA_ARGS__); break;             \
    case 8: UB_DISPATCH_PP(FN, 8, PPv, __VA_ARGS__); break;             \
    default: UB_DISPATCH_PP(FN, 16, PPv, __VA_ARGS__); break;           \
  }

static void fill_common(SampleArgs& a, const Tuning& t, const float* value, const float* qproj, float* out, int B,
                        int bev_h, int bev_w, int fH, int fW, int H, int P, int ld, int off_col, int logit_col) {
  a.value = value, a.qproj = qproj, a.out = out;
  a.ref_cam = nullptr, a.mask = nullptr, a.N = 0, a.D = 1;
  a.B = B, a.bev_h = bev_h, a.bev_w = bev_w, a.fH = fH, a.fW = fW, a.H = H, a.P = P;
  a.ld = ld, a.off_col = off_col, a.logit_col = logit_col;
  a.tile_w_log2 = t.tile_w_log2;
  a.tiles_x = (bev_w + (1 << t.tile_w_log2) - 1) >> t.tile_w_log2;
  const int tile_h = kTileQ >> t.tile_w_log2;
  a.n_tiles = a.tiles_x * ((bev_h + tile_h - 1) / tile_h);
  a.n_chunks = a.n_units = 0;
  a.sx = (float)fW / (float)bev_w, a.sy = (float)fH / (float)bev_h;
  a.vec2_ok = (ld % 2 == 0) && (off_col % 2 == 0) && (reinterpret_cast<uintptr_t>(qproj) & 7u) == 0;
}

extern "C" int ub_bev_sample_fwd(const float* value, const float* qproj, float* out, int B, int bev_h, int bev_w,
                                 int fH, int fW, int H, int Dh, int P, int ld, int off_col, int logit_col,
                                 ub_stream_t stream) {
  if (int rc = check_sample_args("ub_bev_sample_fwd", H, Dh, P, ld, off_col, logit_col)) return rc;
  UB_REQUIRE(value && qproj && out, "ub_bev_sample_fwd: null pointer");
  UB_REQUIRE(B > 0 && bev_h > 0 && bev_w > 0, "ub_bev_sample_fwd: non-positive dimension");
  UB_REQUIRE(fH >= 2 && fW >= 2 && (int64_t)fH * fW < (1 << 30), "ub_bev_sample_fwd: value map %d x %d unsupported",
             fH, fW);
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(out);
  SampleArgs a;
  fill_common(a, g_bev_tuning, value, qproj, out, B, bev_h, bev_w, fH, fW, H, P, ld, off_col, logit_col);
  int rc = 0;
  UB_DISPATCH(launch_bev, Dh, pad_pow2(P), a, g_bev_tuning, (cudaStream_t)stream);
  if (rc) return rc;
  return check_launch("ub_bev_sample_fwd");
}

extern "C" int ub_img_sample_fwd(const float* value, const float* qproj, const float* ref_cam, const uint8_t* mask,
                                 float* out, int B, int N, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P,
                                 int D, int ld, int off_col, int logit_col, ub_stream_t stream) {
  if (int rc = check_sample_args("ub_img_sample_fwd", H, Dh, P, ld, off_col, logit_col)) return rc;
  UB_REQUIRE(value && qproj && ref_cam && mask && out, "ub_img_sample_fwd: null pointer");
  UB_REQUIRE(B > 0 && N > 0 && N <= 32 && bev_h > 0 && bev_w > 0 && fH > 0 && fW > 0 && D > 0 && D <= 8,
             "ub_img_sample_fwd: bad dimension (B=%d N=%d D=%d)", B, N, D);
  UB_REQUIRE(P % D == 0, "ub_img_sample_fwd: num_points %d must be a multiple of the %d Z-anchors", P, D);
  UB_REQUIRE(fH >= 2 && fW >= 2 && (int64_t)fH * fW < (1 << 30), "ub_img_sample_fwd: value map %d x %d unsupported",
             fH, fW);
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(out);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(ref_cam) & 7u) == 0, "ub_img_sample_fwd: ref_cam not 8-byte aligned");
  UB_REQUIRE((int64_t)fH * fW * H * Dh < (1ll << 31) && (int64_t)B * N < (1 << 20),
             "ub_img_sample_fwd: per-camera value map too large");
  SampleArgs a;
  fill_common(a, g_img_tuning, value, qproj, out, B, bev_h, bev_w, fH, fW, H, P, ld, off_col, logit_col);
  a.ref_cam = ref_cam, a.mask = mask, a.N = N, a.D = D;
  int rc = 0;
  UB_DISPATCH(launch_img, Dh, pad_pow2(P), a, g_img_tuning, (cudaStream_t)stream);
  if (rc) return rc;
  return check_launch("ub_img_sample_fwd");
}
