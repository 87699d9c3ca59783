// Fused per-modality deformable sampling for the UniBEV BEV encoder (one feature level).
//
//   ub_project_points  : pillar reference points -> camera planes + visibility bits      [R2]
//   ub_bev_sample_fwd  : BEV self-attention / LiDAR cross-attention sampling              [R4]
//   ub_img_sample_fwd  : camera cross-attention sampling, summed over cameras / count     [R3]
//
// The sampling kernels never see a sampling_locations / attention_weights tensor: they read the raw
// outputs of the sampling_offsets / attention_weights linears (one fused GEMM), build the reference point
// in-kernel, normalise the offsets, run the softmax over the P logits in registers, gather and reduce.
//
// Decomposition: a CTA owns a 2-D tile of BEV queries and a chunk of heads, and walks heads in the outer
// loop, so at any moment its gathers touch one head's 4*LPG... = Dh*4-byte slice of a compact patch of the
// value map -> the patch stays L1-resident while 32 corner fetches per (query, head) hit it.  A work item
// (q, h) is owned by LPG = Dh/4 adjacent lanes, 4 channels each: every corner fetch is one fully used
// 16*LPG-byte segment.  Query-side streams (offsets/logits in, output out) bypass L1.
#include "ub_common.cuh"

namespace ub {

struct Tuning {
  int tile_w = 16, tile_h = 8, heads_per_cta = 0 /* 0 = all */, threads = 256;
};
static Tuning g_bev_tuning, g_img_tuning;

struct ProjParams {
  float zs[8];
  float sx, sy, sz, x0, y0, z0;
  float img_h, img_w;
};

__global__ void __launch_bounds__(256) project_points_kernel(const float* __restrict__ lidar2img, ProjParams pp,
                                                             float* __restrict__ ref_cam, uint8_t* __restrict__ mask,
                                                             int B, int N, int bev_h, int bev_w, int D) {
  const int Nq = bev_h * bev_w;
  const int64_t total = (int64_t)B * Nq * N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int cam = (int)(idx % N);
    const int64_t bq = idx / N;
    const int q = (int)(bq % Nq), b = (int)(bq / Nq);
    const int qx = q % bev_w, qy = q / bev_w;
    // normalised cell centre -> metres; mul and add rounded separately like the reference's tensor ops
    const float x = __fadd_rn(__fmul_rn(((float)qx + 0.5f) / (float)bev_w, pp.sx), pp.x0);
    const float y = __fadd_rn(__fmul_rn(((float)qy + 0.5f) / (float)bev_h, pp.sy), pp.y0);
    const float* m = lidar2img + ((int64_t)b * N + cam) * 16;
    float mm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) mm[i] = __ldg(m + i);
    unsigned bits = 0;
    for (int d = 0; d < D; ++d) {
      const float z = __fadd_rn(__fmul_rn(pp.zs[d], pp.sz), pp.z0);
      const float cx = fmaf(mm[3], 1.f, fmaf(mm[2], z, fmaf(mm[1], y, mm[0] * x)));
      const float cy = fmaf(mm[7], 1.f, fmaf(mm[6], z, fmaf(mm[5], y, mm[4] * x)));
      const float cz = fmaf(mm[11], 1.f, fmaf(mm[10], z, fmaf(mm[9], y, mm[8] * x)));
      const float eps = 1e-5f;
      const float zc = fmaxf(cz, eps);
      const float u = (cx / zc) / pp.img_w;
      const float v = (cy / zc) / pp.img_h;
      const bool vis = (cz > eps) && (v > 0.f) && (v < 1.f) && (u < 1.f) && (u > 0.f);
      bits |= (vis ? 1u : 0u) << d;
      reinterpret_cast<float2*>(ref_cam)[idx * D + d] = make_float2(u, v);
    }
    mask[idx] = (uint8_t)bits;
  }
}

// softmax over P logits + offsets of head h of one query row, all in registers
template <int PMAX>
struct HeadParams {
  float ox[PMAX], oy[PMAX], w[PMAX];
};

template <int PMAX>
__device__ __forceinline__ void load_head_params(HeadParams<PMAX>& hp, const float* __restrict__ row, int off_col,
                                                 int logit_col, int h, int P, bool vec) {
  const float* op = row + off_col + h * P * 2;
  const float* lp = row + logit_col + h * P;
  if (vec && P == PMAX) {
#pragma unroll
    for (int i = 0; i < PMAX / 2; ++i) {
      const float4 t = ld_stream4(op + 4 * i);
      hp.ox[2 * i] = t.x, hp.oy[2 * i] = t.y, hp.ox[2 * i + 1] = t.z, hp.oy[2 * i + 1] = t.w;
    }
#pragma unroll
    for (int i = 0; i < PMAX / 4; ++i) {
      const float4 t = ld_stream4(lp + 4 * i);
      hp.w[4 * i] = t.x, hp.w[4 * i + 1] = t.y, hp.w[4 * i + 2] = t.z, hp.w[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
      hp.ox[p] = p < P ? ld_stream1(op + 2 * p) : 0.f;
      hp.oy[p] = p < P ? ld_stream1(op + 2 * p + 1) : 0.f;
      hp.w[p] = p < P ? ld_stream1(lp + p) : -INFINITY;
    }
  }
  float mx = hp.w[0];
#pragma unroll
  for (int p = 1; p < PMAX; ++p) mx = fmaxf(mx, hp.w[p]);
  float sum = 0.f;
#pragma unroll
  for (int p = 0; p < PMAX; ++p) {
    hp.w[p] = expf(hp.w[p] - mx);
    sum += hp.w[p];
  }
#pragma unroll
  for (int p = 0; p < PMAX; ++p) hp.w[p] = hp.w[p] / sum;
}

template <int LPG, int PMAX>
__global__ void __launch_bounds__(1024)
    bev_sample_kernel(const float* __restrict__ value, const float* __restrict__ qproj, float* __restrict__ out,
                      int bev_h, int bev_w, int fH, int fW, int H, int P, int ld, int off_col, int logit_col,
                      int tile_w, int tile_h, int heads_per_cta, int vec_ok) {
  constexpr int Dh = LPG * 4;
  const int lane = threadIdx.x % LPG, group = threadIdx.x / LPG, n_groups = blockDim.x / LPG;
  const int tiles_x = (bev_w + tile_w - 1) / tile_w;
  const int tx0 = (blockIdx.x % tiles_x) * tile_w, ty0 = (blockIdx.x / tiles_x) * tile_h;
  const int b = blockIdx.z, Nq = bev_h * bev_w, row = H * Dh;
  const int h_begin = blockIdx.y * heads_per_cta, h_end = min(H, h_begin + heads_per_cta);
  const float* vbase = value + (int64_t)b * fH * fW * row + lane * 4;
  for (int h = h_begin; h < h_end; ++h) {
    for (int t = group; t < tile_w * tile_h; t += n_groups) {
      const int qx = tx0 + t % tile_w, qy = ty0 + t / tile_w;
      if (qx >= bev_w || qy >= bev_h) continue;
      const int64_t bq = (int64_t)b * Nq + qy * bev_w + qx;
      HeadParams<PMAX> hp;
      load_head_params<PMAX>(hp, qproj + bq * ld, off_col, logit_col, h, P, vec_ok);
      const float rx = ((float)qx + 0.5f) / (float)bev_w, ry = ((float)qy + 0.5f) / (float)bev_h;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < PMAX; ++p) {
        if (p < P) {
          const float lx = rx + hp.ox[p] / (float)fW, ly = ry + hp.oy[p] / (float)fH;
          bilinear_acc4(acc, vbase + h * Dh, fH, fW, row, ly * fH - 0.5f, lx * fW - 0.5f, hp.w[p]);
        }
      }
      st_stream4(out + bq * row + h * Dh + lane * 4, acc);
    }
  }
}

template <int LPG, int PMAX>
__global__ void __launch_bounds__(1024)
    img_sample_kernel(const float* __restrict__ value, const float* __restrict__ qproj,
                      const float* __restrict__ ref_cam, const uint8_t* __restrict__ mask, float* __restrict__ out,
                      int N, int bev_h, int bev_w, int fH, int fW, int H, int P, int D, int ld, int off_col,
                      int logit_col, int tile_w, int tile_h, int heads_per_cta, int vec_ok) {
  constexpr int Dh = LPG * 4;
  const int lane = threadIdx.x % LPG, group = threadIdx.x / LPG, n_groups = blockDim.x / LPG;
  const int tiles_x = (bev_w + tile_w - 1) / tile_w;
  const int tx0 = (blockIdx.x % tiles_x) * tile_w, ty0 = (blockIdx.x / tiles_x) * tile_h;
  const int b = blockIdx.z, Nq = bev_h * bev_w, row = H * Dh;
  const int h_begin = blockIdx.y * heads_per_cta, h_end = min(H, h_begin + heads_per_cta);
  const int64_t cam_stride = (int64_t)fH * fW * row;
  for (int h = h_begin; h < h_end; ++h) {
    for (int t = group; t < tile_w * tile_h; t += n_groups) {
      const int qx = tx0 + t % tile_w, qy = ty0 + t / tile_w;
      if (qx >= bev_w || qy >= bev_h) continue;
      const int q = qy * bev_w + qx;
      const int64_t bq = (int64_t)b * Nq + q;
      // cameras that contribute: batch item 0's visibility (reference quirk); divisor: this item's own
      unsigned hit = 0;
      int count = 0;
      for (int n = 0; n < N; ++n) {
        hit |= (mask[(int64_t)q * N + n] != 0 ? 1u : 0u) << n;
        count += mask[bq * N + n] != 0 ? 1 : 0;
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hit) {
        HeadParams<PMAX> hp;
        load_head_params<PMAX>(hp, qproj + bq * ld, off_col, logit_col, h, P, vec_ok);
        while (hit) {
          const int n = __ffs(hit) - 1;
          hit &= hit - 1;
          const float2* refs = reinterpret_cast<const float2*>(ref_cam) + (bq * N + n) * D;
          const float* vb = value + ((int64_t)b * N + n) * cam_stride + h * Dh + lane * 4;
          float4 cam_acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int p = 0; p < PMAX; ++p) {
            if (p < P) {
              const float2 r = __ldg(refs + (p % D));
              const float lx = r.x + hp.ox[p] / (float)fW, ly = r.y + hp.oy[p] / (float)fH;
              bilinear_acc4(cam_acc, vb, fH, fW, row, ly * fH - 0.5f, lx * fW - 0.5f, hp.w[p]);
            }
          }
          acc.x += cam_acc.x, acc.y += cam_acc.y, acc.z += cam_acc.z, acc.w += cam_acc.w;
        }
        const float c = (float)max(count, 1);
        acc.x /= c, acc.y /= c, acc.z /= c, acc.w /= c;
      }
      st_stream4(out + bq * row + h * Dh + lane * 4, acc);
    }
  }
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace ub

using namespace ub;

extern "C" int ub_set_tuning(int which, int tile_w, int tile_h, int heads_per_cta, int threads) {
  UB_REQUIRE(tile_w > 0 && tile_h > 0 && heads_per_cta >= 0 && threads >= 32 && threads <= 1024 && threads % 32 == 0,
             "ub_set_tuning: bad values");
  Tuning& t = which == 0 ? g_bev_tuning : g_img_tuning;
  t.tile_w = tile_w, t.tile_h = tile_h, t.heads_per_cta = heads_per_cta, t.threads = threads;
  return UB_OK;
}

extern "C" int ub_project_points(const float* lidar2img, const float* zs_host, const float* pc_range_host,
                                 float img_h, float img_w, float* ref_cam, uint8_t* mask, int B, int N, int bev_h,
                                 int bev_w, int D, ub_stream_t stream) {
  UB_REQUIRE(lidar2img && zs_host && pc_range_host && ref_cam && mask, "ub_project_points: null pointer");
  UB_REQUIRE(B > 0 && N > 0 && N <= 32 && bev_h > 0 && bev_w > 0 && D > 0 && D <= 8,
             "ub_project_points: need B>0, 0<N<=32, bev dims>0, 0<D<=8 (got B=%d N=%d %dx%d D=%d)", B, N, bev_h, bev_w,
             D);
  UB_REQUIRE(img_h > 0.f && img_w > 0.f, "ub_project_points: image size must be positive");
  ProjParams pp;
  for (int d = 0; d < 8; ++d) pp.zs[d] = d < D ? zs_host[d] : 0.f;
  // (pc_range[3] - pc_range[0]) is evaluated in Python doubles by the reference, then rounded once
  pp.sx = (float)((double)pc_range_host[3] - (double)pc_range_host[0]);
  pp.sy = (float)((double)pc_range_host[4] - (double)pc_range_host[1]);
  pp.sz = (float)((double)pc_range_host[5] - (double)pc_range_host[2]);
  pp.x0 = pc_range_host[0], pp.y0 = pc_range_host[1], pp.z0 = pc_range_host[2];
  pp.img_h = img_h, pp.img_w = img_w;
  const int64_t total = (int64_t)B * bev_h * bev_w * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  project_points_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lidar2img, pp, ref_cam, mask, B, N, bev_h, bev_w, D);
  return check_launch("ub_project_points");
}

static int check_sample_args(const char* fn, int H, int Dh, int P, int ld, int off_col, int logit_col) {
  UB_REQUIRE(H > 0 && P > 0 && P <= 16, "%s: need H>0 and 0<P<=16 (got H=%d P=%d)", fn, H, P);
  UB_REQUIRE(Dh % 4 == 0 && pow2(Dh / 4) && Dh <= 128, "%s: head dim %d unsupported (need 4*2^k <= 128)", fn, Dh);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small for off_col=%d logit_col=%d H=%d P=%d", fn, ld, off_col, logit_col, H,
             P);
  return UB_OK;
}

#define UB_DISPATCH_LPG_P(KERNEL, LPGv, Pv, ...)                       \
  do {                                                                 \
    if (Pv <= 4) { KERNEL(LPGv, 4, __VA_ARGS__); }                     \
    else if (Pv <= 8) { KERNEL(LPGv, 8, __VA_ARGS__); }                \
    else { KERNEL(LPGv, 16, __VA_ARGS__); }                            \
  } while (0)

extern "C" int ub_bev_sample_fwd(const float* value, const float* qproj, float* out, int B, int bev_h, int bev_w,
                                 int fH, int fW, int H, int Dh, int P, int ld, int off_col, int logit_col,
                                 ub_stream_t stream) {
  if (int rc = check_sample_args("ub_bev_sample_fwd", H, Dh, P, ld, off_col, logit_col)) return rc;
  UB_REQUIRE(value && qproj && out, "ub_bev_sample_fwd: null pointer");
  UB_REQUIRE(B > 0 && bev_h > 0 && bev_w > 0 && fH > 0 && fW > 0, "ub_bev_sample_fwd: non-positive dimension");
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(out);
  const Tuning t = g_bev_tuning;
  const int hpc = t.heads_per_cta > 0 ? min(t.heads_per_cta, H) : H;
  const int vec_ok = (ld % 4 == 0) && (off_col % 4 == 0) && (logit_col % 4 == 0) && (P % 4 == 0) &&
                     (reinterpret_cast<uintptr_t>(qproj) & 15u) == 0;
  dim3 grid(((bev_w + t.tile_w - 1) / t.tile_w) * ((bev_h + t.tile_h - 1) / t.tile_h), (H + hpc - 1) / hpc, B);
  cudaStream_t s = (cudaStream_t)stream;
#define UB_LAUNCH(LPGv, PMAXv, dummy)                                                                            \
  bev_sample_kernel<LPGv, PMAXv><<<grid, t.threads, 0, s>>>(value, qproj, out, bev_h, bev_w, fH, fW, H, P, ld,   \
                                                            off_col, logit_col, t.tile_w, t.tile_h, hpc, vec_ok)
  switch (Dh / 4) {
    case 1: UB_DISPATCH_LPG_P(UB_LAUNCH, 1, P, 0); break;
    case 2: UB_DISPATCH_LPG_P(UB_LAUNCH, 2, P, 0); break;
    case 4: UB_DISPATCH_LPG_P(UB_LAUNCH, 4, P, 0); break;
    case 8: UB_DISPATCH_LPG_P(UB_LAUNCH, 8, P, 0); break;
    case 16: UB_DISPATCH_LPG_P(UB_LAUNCH, 16, P, 0); break;
    default: UB_DISPATCH_LPG_P(UB_LAUNCH, 32, P, 0); break;
  }
#undef UB_LAUNCH
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
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(out);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(ref_cam) & 7u) == 0, "ub_img_sample_fwd: ref_cam not 8-byte aligned");
  const Tuning t = g_img_tuning;
  const int hpc = t.heads_per_cta > 0 ? min(t.heads_per_cta, H) : H;
  const int vec_ok = (ld % 4 == 0) && (off_col % 4 == 0) && (logit_col % 4 == 0) && (P % 4 == 0) &&
                     (reinterpret_cast<uintptr_t>(qproj) & 15u) == 0;
  dim3 grid(((bev_w + t.tile_w - 1) / t.tile_w) * ((bev_h + t.tile_h - 1) / t.tile_h), (H + hpc - 1) / hpc, B);
  cudaStream_t s = (cudaStream_t)stream;
#define UB_LAUNCH(LPGv, PMAXv, dummy)                                                                             \
  img_sample_kernel<LPGv, PMAXv><<<grid, t.threads, 0, s>>>(value, qproj, ref_cam, mask, out, N, bev_h, bev_w, fH, \
                                                            fW, H, P, D, ld, off_col, logit_col, t.tile_w, t.tile_h, \
                                                            hpc, vec_ok)
  switch (Dh / 4) {
    case 1: UB_DISPATCH_LPG_P(UB_LAUNCH, 1, P, 0); break;
    case 2: UB_DISPATCH_LPG_P(UB_LAUNCH, 2, P, 0); break;
    case 4: UB_DISPATCH_LPG_P(UB_LAUNCH, 4, P, 0); break;
    case 8: UB_DISPATCH_LPG_P(UB_LAUNCH, 8, P, 0); break;
    case 16: UB_DISPATCH_LPG_P(UB_LAUNCH, 16, P, 0); break;
    default: UB_DISPATCH_LPG_P(UB_LAUNCH, 32, P, 0); break;
  }
#undef UB_LAUNCH
  return check_launch("ub_img_sample_fwd");
}
