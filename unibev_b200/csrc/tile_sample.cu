// Fused per-modality deformable sampling for the UniBEV BEV encoder (one feature level).
//
//   ub_bev_sample_fwd  : BEV self-attention / LiDAR cross-attention sampling              [R4]
//   ub_img_sample_fwd  : camera cross-attention sampling, summed over cameras / count     [R3]
//
// The kernels never see a sampling_locations / attention_weights tensor: they read the raw outputs of
// the sampling_offsets / attention_weights linears (one fused GEMM), build the reference point in-kernel,
// normalise the offsets, run the softmax over the P logits, gather bilinearly and reduce.
//
// Two phases per CTA round, 256 threads, a tile of 64 BEV queries x HC heads:
//   phase 1  one thread per SAMPLE (query, head, point): coalesced read of its offset pair and logit,
//            softmax across the P adjacent lanes with shuffles, location -> a 16-byte sample descriptor
//            {attention weight, lw, lh, top-left pixel index | 4 corner-valid bits} in shared memory.
//            Scalar work is done once per sample instead of once per channel lane.
//   phase 2  one group of LPG = Dh/4 lanes per ITEM (query, head), four channels per lane: per point one
//            broadcast LDS.128 of the descriptor, up to four predicated LDG.128 corner fetches (each a
//            fully used 16*LPG-byte segment: 128 B at Dh = 32), 16 FFMAs.  Items are ordered head-major
//            so that all groups of the CTA gather from ONE head's 128-byte column of a compact patch of
//            the value map at a time -> the patch stays L1-resident.
// Descriptor slots are XOR-swizzled so the four groups of a warp hit four different 16-byte bank groups.
// Query-side streams (offsets/logits in, output out) bypass L1 (no-allocate).
#include "ub_common.cuh"

namespace ub {

constexpr int kThreads = 256;
constexpr int kTileQ = 64;  // BEV queries per tile (tile_w x tile_h, tile_w a power of two <= 64)

struct Tuning {
  int tile_w_log2 = 3;  // 8 x 8
  int min_ctas = 0;     // reserved
};
static Tuning g_bev_tuning, g_img_tuning;

// Sample descriptor, produced once per (query, head, point) in phase 1:
//   weights {w00, w01, w10, w11} = attention weight x bilinear corner weights of the 2 x 2 pixel block whose
//   top-left pixel is `index`; the block is shifted to lie inside the map ([0, fH-2] x [0, fW-2]) and the weights
//   of the corners the reference treats as zero padding are 0, so phase 2 loads unconditionally.
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
    // rows: (weight of block row 0, weight of block row 1, block row)
    float wt = hh, wb = lh, wl = hw, wr = lw;
    int yb = h0, xb = w0;
    if (h0 < 0) wt = lh, wb = 0.f, yb = 0;                 // only pixel row 0 (the reference's bottom corner)
    else if (h0 > fH - 2) wt = 0.f, wb = hh, yb = fH - 2;  // only pixel row fH-1 (the reference's top corner)
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

// Shared-memory descriptor store: weights as float4 per sample, indices as int per sample read back four at a
// time.  Slots are XOR-swizzled so that the (up to) four groups of a warp, which work on consecutive items and
// the same point, hit different 16-byte bank groups.
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
  const float e = ok ? expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = PP / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  return ok ? e / sum : 0.f;
}

struct SampleArgs {
  const float* value;
  const float* qproj;
  float* out;
  const float* ref_cam;   // camera mode only
  const uint8_t* mask;    // camera mode only
  int N, D;               // cameras, Z-anchors (camera mode)
  int bev_h, bev_w, fH, fW, H, P, ld, off_col, logit_col;
  int tile_w_log2, tiles_x;
  int vec2_ok;            // offsets readable as float2
};

// per-kernel compile-time geometry
template <int LPG, int PP, bool IMG>
struct Geo {
  static constexpr int Dh = LPG * 4;
  static constexpr int n_groups = kThreads / LPG;
  // heads per round.  BEV mode: 1024 samples; camera mode: every group owns exactly IPG items whose
  // accumulators stay in registers across the camera loop.
  static constexpr int HC = IMG ? (n_groups / 16 > 0 ? n_groups / 16 : 1) : ((16 / PP) > 0 ? (16 / PP) : 1);
  static constexpr int n_items = kTileQ * HC, n_samples = n_items * PP;
  static constexpr int IPG = (n_items + n_groups - 1) / n_groups;
  static constexpr size_t smem = (size_t)n_samples * (sizeof(float4) + sizeof(int));
};

// phase 2 for one item: PP points, descriptors from shared memory
template <int PP, int ROW>
__device__ __forceinline__ void gather_item(float4& acc, const float4* __restrict__ s_w, const int* __restrict__ s_i,
                                            int item, const float* __restrict__ vbase, int row, int fW, bool skip_miss) {
  if (PP >= 4) {
#pragma unroll
    for (int p4 = 0; p4 < PP / 4; ++p4) {
      const int4 idx = *reinterpret_cast<const int4*>(s_i + i_slot<PP>(item, p4 * 4));
      const int ii[4] = {idx.x, idx.y, idx.z, idx.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int p = p4 * 4 + k;
        if (skip_miss) {
          if (ii[k] >= 0) gather_point<ROW>(acc, s_w[w_slot<PP>(item, p)], ii[k], vbase, row, fW);
        } else {
          gather_point<ROW>(acc, s_w[w_slot<PP>(item, p)], max(ii[k], 0), vbase, row, fW);
        }
      }
    }
  } else {
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      const int i = s_i[i_slot<PP>(item, p)];
      if (i >= 0) gather_point<ROW>(acc, s_w[w_slot<PP>(item, p)], i, vbase, row, fW);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// BEV-grid mode: reference point = cell centre, one value map (B, fH*fW, H*Dh).
// grid = (tiles, ceil(H / HC), B).  HF > 0: number of heads known at compile time.
template <int LPG, int PP, int HF>
__global__ void __launch_bounds__(kThreads) bev_sample_kernel(const SampleArgs a) {
  using G = Geo<LPG, PP, false>;
  constexpr int Dh = G::Dh, HC = G::HC, n_items = G::n_items, n_samples = G::n_samples, n_groups = G::n_groups;
  constexpr int ROW = HF * Dh;
  __shared__ float4 s_w[n_samples];
  __shared__ __align__(16) int s_i[n_samples];

  const int tid = threadIdx.x;
  const int tw_mask = (1 << a.tile_w_log2) - 1;
  const int tx0 = (blockIdx.x % a.tiles_x) << a.tile_w_log2;
  const int ty0 = (blockIdx.x / a.tiles_x) * (kTileQ >> a.tile_w_log2);
  const int h0 = blockIdx.y * HC, b = blockIdx.z;
  const int H = HF > 0 ? HF : a.H;
  const int Nq = a.bev_h * a.bev_w, row = H * Dh;

  // ---- phase 1: descriptors
#pragma unroll
  for (int s0 = 0; s0 < n_samples; s0 += kThreads) {
    const int s = s0 + tid;
    const bool live = s < n_samples;
    const int p = s % PP, item = s / PP;
    const int ql = item % kTileQ, hl = item / kTileQ;
    const int qx = tx0 + (ql & tw_mask), qy = ty0 + (ql >> a.tile_w_log2), h = h0 + hl;
    const bool ok = live && qx < a.bev_w && qy < a.bev_h && h < H && p < a.P;
    float ox = 0.f, oy = 0.f, logit = -INFINITY;
    if (ok) {
      const float* rowp = a.qproj + ((int64_t)b * Nq + qy * a.bev_w + qx) * a.ld;
      const float* op = rowp + a.off_col + (h * a.P + p) * 2;
      if (a.vec2_ok) {
        const float2 t = ld_stream2(op);
        ox = t.x, oy = t.y;
      } else {
        ox = ld_stream1(op), oy = ld_stream1(op + 1);
      }
      logit = ld_stream1(rowp + a.logit_col + h * a.P + p);
    }
    const float aw = group_softmax<PP>(logit, ok);
    if (live) {
      SampleDesc d;
      d.w = make_float4(0.f, 0.f, 0.f, 0.f), d.index = -1;
      if (ok) {
        const float rx = ((float)qx + 0.5f) / (float)a.bev_w, ry = ((float)qy + 0.5f) / (float)a.bev_h;
        const float lx = rx + ox / (float)a.fW, ly = ry + oy / (float)a.fH;
        d = make_desc(ly * a.fH - 0.5f, lx * a.fW - 0.5f, aw, a.fH, a.fW);
      }
      s_w[w_slot<PP>(item, p)] = d.w;
      s_i[i_slot<PP>(item, p)] = d.index;
    }
  }
  __syncthreads();

  // ---- phase 2: gather + reduce
  const int lane = tid % LPG, group = tid / LPG;
  const float* vb = a.value + (int64_t)b * a.fH * a.fW * row + lane * 4;
  for (int item = group; item < n_items; item += n_groups) {
    const int ql = item % kTileQ, hl = item / kTileQ;
    const int qx = tx0 + (ql & tw_mask), qy = ty0 + (ql >> a.tile_w_log2), h = h0 + hl;
    if (qx >= a.bev_w || qy >= a.bev_h || h >= H) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_item<PP, ROW>(acc, s_w, s_i, item, vb + h * Dh, row, a.fW, false);
    st_stream4(a.out + ((int64_t)b * Nq + qy * a.bev_w + qx) * row + h * Dh + lane * 4, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Camera mode: value (B, N, fH*fW, H*Dh); reference points / visibility from ub_project_points.
// A camera contributes to (b, q) iff batch item 0 sees q in it (reference quirk, sca_img:142); the sum over
// cameras is divided by max(1, #cameras that see (b, q)) (sca_img:209-212).
// Every group keeps its IPG items' accumulators in registers across the camera loop (same item -> same group).
template <int LPG, int PP, int HF>
__global__ void __launch_bounds__(kThreads) img_sample_kernel(const SampleArgs a) {
  using G = Geo<LPG, PP, true>;
  constexpr int Dh = G::Dh, HC = G::HC, n_items = G::n_items, n_samples = G::n_samples, n_groups = G::n_groups;
  constexpr int IPG = G::IPG, ROW = HF * Dh;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_w = reinterpret_cast<float4*>(smem_raw);
  int* s_i = reinterpret_cast<int*>(smem_raw + (size_t)n_samples * sizeof(float4));
  __shared__ unsigned s_hit[kTileQ];    // cameras (bit n) that contribute to the query
  __shared__ float s_count[kTileQ];     // divisor
  __shared__ unsigned s_tile_hit;

  const int tid = threadIdx.x;
  const int tw_mask = (1 << a.tile_w_log2) - 1;
  const int tx0 = (blockIdx.x % a.tiles_x) << a.tile_w_log2;
  const int ty0 = (blockIdx.x / a.tiles_x) * (kTileQ >> a.tile_w_log2);
  const int h0 = blockIdx.y * HC, b = blockIdx.z;
  const int H = HF > 0 ? HF : a.H;
  const int Nq = a.bev_h * a.bev_w, row = H * Dh;
  const int64_t cam_stride = (int64_t)a.fH * a.fW * row;

  if (tid == 0) s_tile_hit = 0u;
  __syncthreads();
  if (tid < kTileQ) {
    const int qx = tx0 + (tid & tw_mask), qy = ty0 + (tid >> a.tile_w_log2);
    unsigned hit = 0u;
    int count = 0;
    if (qx < a.bev_w && qy < a.bev_h) {
      const int q = qy * a.bev_w + qx;
      for (int n = 0; n < a.N; ++n) {
        hit |= (a.mask[(int64_t)q * a.N + n] != 0 ? 1u : 0u) << n;
        count += a.mask[((int64_t)b * Nq + q) * a.N + n] != 0 ? 1 : 0;
      }
    }
    s_hit[tid] = hit;
    s_count[tid] = (float)max(count, 1);
    if (hit) atomicOr(&s_tile_hit, hit);
  }
  __syncthreads();
  unsigned cams = s_tile_hit;

  const int lane = tid % LPG, group = tid / LPG;
  float4 acc[IPG];
#pragma unroll
  for (int k = 0; k < IPG; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  while (cams) {
    const int n = __ffs(cams) - 1;
    cams &= cams - 1;
    // ---- phase 1: descriptors of camera n
    for (int s0 = 0; s0 < n_samples; s0 += kThreads) {
      const int s = s0 + tid;
      const bool live = s < n_samples;
      const int p = s % PP, item = s / PP;
      const int ql = item % kTileQ, hl = item / kTileQ;
      const int qx = tx0 + (ql & tw_mask), qy = ty0 + (ql >> a.tile_w_log2), h = h0 + hl;
      const bool ok = live && h < H && p < a.P && ((s_hit[ql] >> n) & 1u);
      float ox = 0.f, oy = 0.f, logit = -INFINITY;
      float2 r = make_float2(0.f, 0.f);
      if (ok) {
        const int64_t bq = (int64_t)b * Nq + qy * a.bev_w + qx;
        const float* rowp = a.qproj + bq * a.ld;
        const float* op = rowp + a.off_col + (h * a.P + p) * 2;
        if (a.vec2_ok) {
          const float2 t = ld_stream2(op);
          ox = t.x, oy = t.y;
        } else {
          ox = ld_stream1(op), oy = ld_stream1(op + 1);
        }
        logit = ld_stream1(rowp + a.logit_col + h * a.P + p);
        r = __ldg(reinterpret_cast<const float2*>(a.ref_cam) + (bq * a.N + n) * a.D + (p % a.D));
      }
      const float aw = group_softmax<PP>(logit, ok);
      if (live) {
        SampleDesc d;
        d.w = make_float4(0.f, 0.f, 0.f, 0.f), d.index = -1;
        if (ok) {
          const float lx = r.x + ox / (float)a.fW, ly = r.y + oy / (float)a.fH;
          d = make_desc(ly * a.fH - 0.5f, lx * a.fW - 0.5f, aw, a.fH, a.fW);
        }
        s_w[w_slot<PP>(item, p)] = d.w;
        s_i[i_slot<PP>(item, p)] = d.index;
      }
    }
    __syncthreads();
    // ---- phase 2
    const float* vb = a.value + ((int64_t)b * a.N + n) * cam_stride + lane * 4;
#pragma unroll
    for (int k = 0; k < IPG; ++k) {
      const int item = group + k * n_groups;
      if (item < n_items) {
        const int ql = item % kTileQ, hl = item / kTileQ;
        if (h0 + hl < H && ((s_hit[ql] >> n) & 1u)) {
          float4 cam_acc = make_float4(0.f, 0.f, 0.f, 0.f);
          gather_item<PP, ROW>(cam_acc, s_w, s_i, item, vb + (h0 + hl) * Dh, row, a.fW, true);
          acc[k].x += cam_acc.x, acc[k].y += cam_acc.y, acc[k].z += cam_acc.z, acc[k].w += cam_acc.w;
        }
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int k = 0; k < IPG; ++k) {
    const int item = group + k * n_groups;
    if (item >= n_items) continue;
    const int ql = item % kTileQ, hl = item / kTileQ;
    const int qx = tx0 + (ql & tw_mask), qy = ty0 + (ql >> a.tile_w_log2), h = h0 + hl;
    if (qx >= a.bev_w || qy >= a.bev_h || h >= H) continue;
    float4 o = acc[k];
    if (s_hit[ql]) {
      const float c = s_count[ql];
      o.x /= c, o.y /= c, o.z /= c, o.w /= c;
    }
    st_stream4(a.out + ((int64_t)b * Nq + qy * a.bev_w + qx) * row + h * Dh + lane * 4, o);
  }
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static int pad_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int LPG, int PP, int HF>
static int launch_bev_hf(const SampleArgs& a, int B, cudaStream_t s) {
  using G = Geo<LPG, PP, false>;
  const int tile_h = kTileQ >> a.tile_w_log2;
  dim3 grid(a.tiles_x * ((a.bev_h + tile_h - 1) / tile_h), (a.H + G::HC - 1) / G::HC, B);
  bev_sample_kernel<LPG, PP, HF><<<grid, kThreads, 0, s>>>(a);
  return 0;
}
template <int LPG, int PP>
static int launch_bev(const SampleArgs& a, int B, cudaStream_t s) {
  if ((LPG == 8 || LPG == 4) && (PP == 8 || PP == 4) && a.H == 8)  // the shipped configs: immediate row stride
    return launch_bev_hf<LPG, PP, ((LPG == 8 || LPG == 4) && (PP == 8 || PP == 4)) ? 8 : 0>(a, B, s);
  return launch_bev_hf<LPG, PP, 0>(a, B, s);
}

template <int LPG, int PP, int HF>
static int launch_img_hf(const SampleArgs& a, int B, cudaStream_t s) {
  using G = Geo<LPG, PP, true>;
  constexpr size_t smem = G::smem;
  if (smem > 200 * 1024) {
    set_error("ub_img_sample_fwd: head dim %d with %d points needs %zu bytes of shared memory", LPG * 4, PP, smem);
    return UB_EINVAL;
  }
  static bool configured = false;  // per instantiation
  if (!configured) {
    if (smem > 48 * 1024 && cudaFuncSetAttribute(img_sample_kernel<LPG, PP, HF>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("ub_img_sample_fwd: cannot reserve %zu bytes of shared memory", smem);
      return UB_ECUDA;
    }
    configured = true;
  }
  const int tile_h = kTileQ >> a.tile_w_log2;
  dim3 grid(a.tiles_x * ((a.bev_h + tile_h - 1) / tile_h), (a.H + G::HC - 1) / G::HC, B);
  img_sample_kernel<LPG, PP, HF><<<grid, kThreads, smem, s>>>(a);
  return 0;
}
template <int LPG, int PP>
static int launch_img(const SampleArgs& a, int B, cudaStream_t s) {
  if ((LPG == 8 || LPG == 4) && (PP == 8 || PP == 4) && a.H == 8)
    return launch_img_hf<LPG, PP, ((LPG == 8 || LPG == 4) && (PP == 8 || PP == 4)) ? 8 : 0>(a, B, s);
  return launch_img_hf<LPG, PP, 0>(a, B, s);
}

}  // namespace ub

using namespace ub;

// which: 0 = ub_bev_sample_fwd, 1 = ub_img_sample_fwd.  tile_w must be a power of two <= 64 (tile = tile_w x 64/tile_w).
extern "C" int ub_set_tuning(int which, int tile_w, int tile_h, int heads_per_cta, int threads) {
  (void)tile_h, (void)heads_per_cta, (void)threads;
  UB_REQUIRE(pow2(tile_w) && tile_w <= kTileQ, "ub_set_tuning: tile_w must be a power of two <= %d", kTileQ);
  Tuning& t = which == 0 ? g_bev_tuning : g_img_tuning;
  int l = 0;
  while ((1 << l) < tile_w) ++l;
  t.tile_w_log2 = l;
  return UB_OK;
}

static int check_sample_args(const char* fn, int H, int Dh, int P, int ld, int off_col, int logit_col) {
  UB_REQUIRE(H > 0 && P > 0 && P <= 16, "%s: need H>0 and 0<P<=16 (got H=%d P=%d)", fn, H, P);
  UB_REQUIRE(Dh % 4 == 0 && pow2(Dh / 4) && Dh <= 128, "%s: head dim %d unsupported (need 4*2^k <= 128)", fn, Dh);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small for off_col=%d logit_col=%d H=%d P=%d", fn, ld, off_col, logit_col, H,
             P);
  return UB_OK;
}

#define UB_DISPATCH_PP(FN, LPGv, PPv, ...)                       \
  switch (PPv) {                                                 \
    case 1: rc = FN<LPGv, 1>(__VA_ARGS__); break;                \
    case 2: rc = FN<LPGv, 2>(__VA_ARGS__); break;                \
    case 4: rc = FN<LPGv, 4>(__VA_ARGS__); break;                \
    case 8: rc = FN<LPGv, 8>(__VA_ARGS__); break;                \
    default: rc = FN<LPGv, 16>(__VA_ARGS__); break;              \
  }
#define UB_DISPATCH(FN, Dhv, PPv, ...)                                  \
  switch ((Dhv) / 4) {                                                  \
    case 1: UB_DISPATCH_PP(FN, 1, PPv, __VA_ARGS__); break;             \
    case 2: UB_DISPATCH_PP(FN, 2, PPv, __VA_ARGS__); break;             \
    case 4: UB_DISPATCH_PP(FN, 4, PPv, __VA_ARGS__); break;             \
    case 8: UB_DISPATCH_PP(FN, 8, PPv, __VA_ARGS__); break;             \
    case 16: UB_DISPATCH_PP(FN, 16, PPv, __VA_ARGS__); break;           \
    default: UB_DISPATCH_PP(FN, 32, PPv, __VA_ARGS__); break;           \
  }

static void fill_common(SampleArgs& a, const Tuning& t, const float* value, const float* qproj, float* out, int bev_h,
                        int bev_w, int fH, int fW, int H, int P, int ld, int off_col, int logit_col) {
  a.value = value, a.qproj = qproj, a.out = out;
  a.ref_cam = nullptr, a.mask = nullptr, a.N = 0, a.D = 1;
  a.bev_h = bev_h, a.bev_w = bev_w, a.fH = fH, a.fW = fW, a.H = H, a.P = P;
  a.ld = ld, a.off_col = off_col, a.logit_col = logit_col;
  a.tile_w_log2 = t.tile_w_log2;
  a.tiles_x = (bev_w + (1 << t.tile_w_log2) - 1) >> t.tile_w_log2;
  a.vec2_ok = (ld % 2 == 0) && (off_col % 2 == 0) && (reinterpret_cast<uintptr_t>(qproj) & 7u) == 0;
}

extern "C" int ub_bev_sample_fwd(const float* value, const float* qproj, float* out, int B, int bev_h, int bev_w,
                                 int fH, int fW, int H, int Dh, int P, int ld, int off_col, int logit_col,
                                 ub_stream_t stream) {
  if (int rc = check_sample_args("ub_bev_sample_fwd", H, Dh, P, ld, off_col, logit_col)) return rc;
  UB_REQUIRE(value && qproj && out, "ub_bev_sample_fwd: null pointer");
  UB_REQUIRE(B > 0 && bev_h > 0 && bev_w > 0, "ub_bev_sample_fwd: non-positive dimension");
  UB_REQUIRE(fH >= 2 && fW >= 2 && (int64_t)fH * fW < (1 << 30), "ub_bev_sample_fwd: value map %d x %d unsupported", fH, fW);
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(out);
  SampleArgs a;
  fill_common(a, g_bev_tuning, value, qproj, out, bev_h, bev_w, fH, fW, H, P, ld, off_col, logit_col);
  int rc = 0;
  UB_DISPATCH(launch_bev, Dh, pad_pow2(P), a, B, (cudaStream_t)stream);
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
  UB_REQUIRE(fH >= 2 && fW >= 2 && (int64_t)fH * fW < (1 << 30), "ub_img_sample_fwd: value map %d x %d unsupported", fH, fW);
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(out);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(ref_cam) & 7u) == 0, "ub_img_sample_fwd: ref_cam not 8-byte aligned");
  SampleArgs a;
  fill_common(a, g_img_tuning, value, qproj, out, bev_h, bev_w, fH, fW, H, P, ld, off_col, logit_col);
  a.ref_cam = ref_cam, a.mask = mask, a.N = N, a.D = D;
  int rc = 0;
  UB_DISPATCH(launch_img, Dh, pad_pow2(P), a, B, (cudaStream_t)stream);
  if (rc) return rc;
  return check_launch("ub_img_sample_fwd");
}
