// Backward twins of the fused deformable sampling kernels (training step, BASELINE configs[4]).
//
//   ub_bev_sample_bwd  : backward of ub_bev_sample_fwd  (BEV self-attention / LiDAR cross-attention)        [R4]
//   ub_img_sample_bwd  : backward of ub_img_sample_fwd  (camera cross-attention, summed over cameras / count) [R3]
//
// Like the forward kernels they never see a sampling_locations / attention_weights tensor: they read the RAW outputs of
// the sampling_offsets / attention_weights linears (qproj rows), rebuild reference point, normalised offset, softmax and
// bilinear weights in registers, and return the gradient with respect to those raw rows -- the softmax backward
// (g_logit[p] = a[p] (g_a[p] - sum_j a[j] g_a[j])) and the offset normalisation (d pixel / d offset = 1: the 1 / (fW, fH) of
// the normaliser cancels against the pixel scale) are folded in.  What mmcv's op + autograd do in five kernels and four
// (B, Nq, H, L, P, 2)-sized temporaries per attention (reference: spatial_cross_attention_img.py:390-419 forward,
// ms_deform_attn_backward + softmax / add / div backward) is one launch here.
//
// Work item = one (b, q, h).  A group of LPG = Dh / 4 adjacent lanes owns it; each lane carries four channels, so every
// corner fetch / gradient reduction of the group is one fully used 16 * LPG-byte segment (128 B at Dh = 32).
//   grad_value : red.global.add.v4.f32 per corner and lane (the caller zero-fills it)
//   grad_qproj : per-lane partial sums over the lane's channels -- accumulated over cameras in camera mode -- folded
//                across the group's lanes with shuffles once per point, written with plain stores (each (b, q, h, p) is
//                owned by exactly one group: no atomics).  Columns of the row outside [off_col, off_col + 2 H P) and
//                [logit_col, logit_col + H P) are not touched.
#include "ub_common.cuh"

namespace ub {

__device__ __forceinline__ void red_add4g(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float dot4g(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float4 scale4g(float s, const float4& v) { return make_float4(s * v.x, s * v.y, s * v.z, s * v.w); }

// One sample of one item: scatters a * w_c * go into the four corners of grad_value and accumulates this lane's share
// of the gradients with respect to the attention weight (g_a) and the pixel coordinates (g_x, g_y; already times a).
__device__ __forceinline__ void sample_bwd(const float* __restrict__ vbase, float* gbase, int fH, int fW, int row, float h_im,
                                           float w_im, float a, const float4& go, float& g_a, float& g_x, float& g_y) {
  if (!(h_im > -1.f && w_im > -1.f && h_im < (float)fH && w_im < (float)fW)) return;
  const float hf = floorf(h_im), wf = floorf(w_im);
  const int h0 = (int)hf, w0 = (int)wf;
  const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
  const bool top = h0 >= 0, bot = h0 + 1 <= fH - 1, left = w0 >= 0, right = w0 + 1 <= fW - 1;
  const int o00 = (h0 * fW + w0) * row, o10 = o00 + fW * row;   // (one map < 2^31 floats: checked on the host; 32-bit address math)
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 v1 = (top && left) ? ldg4(vbase + o00) : z4;
  const float4 v2 = (top && right) ? ldg4(vbase + o00 + row) : z4;
  const float4 v3 = (bot && left) ? ldg4(vbase + o10) : z4;
  const float4 v4 = (bot && right) ? ldg4(vbase + o10 + row) : z4;
  const float4 tg = scale4g(a, go);
  if (top && left) red_add4g(gbase + o00, scale4g(hh * hw, tg));
  if (top && right) red_add4g(gbase + o00 + row, scale4g(hh * lw, tg));
  if (bot && left) red_add4g(gbase + o10, scale4g(lh * hw, tg));
  if (bot && right) red_add4g(gbase + o10 + row, scale4g(lh * lw, tg));
  const float d1 = dot4g(go, v1), d2 = dot4g(go, v2), d3 = dot4g(go, v3), d4 = dot4g(go, v4);
  g_a += hh * hw * d1 + hh * lw * d2 + lh * hw * d3 + lh * lw * d4;
  g_y += a * (-hw * d1 - lw * d2 + hw * d3 + lw * d4);
  g_x += a * (-hh * d1 + hh * d2 - lh * d3 + lh * d4);
}

struct SampleBwdArgs {
  const float* value;      // BEV: (B, fH*fW, H*Dh); camera: (B, N, fH*fW, H*Dh)
  const float* qproj;      // (B, Nq, ld)
  const float* grad_out;   // (B, Nq, H*Dh)
  float* grad_value;       // like value, zero-filled by the caller
  float* grad_qproj;       // (B, Nq, ld)
  const float* ref_cam;    // camera mode: (B, Nq, N, D, 2)
  const uint8_t* mask;     // camera mode: (B, Nq, N)
  int N, D;
  int B, bev_h, bev_w, fH, fW, H, P, ld, off_col, logit_col;
  float sx, sy;
};

// CAM = false: BEV-grid mode (reference point = cell centre); CAM = true: camera mode (projected anchors, hit cameras of
// batch item 0, divisor from the item's own visibility).  PP >= P: compile-time bound of the point loop.
// Register diet (the kernel is latency-bound: four 128-bit loads and four reductions per sample and lane, nothing to hide
// them behind but other warps): only the softmax weights and the attention-weight gradients of the P points stay in
// registers across the point loop; offsets are read per point and the location gradients are folded across the group's
// lanes and written as soon as the point is done (camera mode: cameras are the INNER loop for that reason).
template <int LPG, int PP, bool CAM>
__global__ void __launch_bounds__(256, CAM ? 3 : 4) sample_bwd_kernel(const SampleBwdArgs a) {
  constexpr int Dh = LPG * 4;
  const int lane = threadIdx.x % LPG;
  const int Nq = a.bev_h * a.bev_w, row = a.H * Dh, P = a.P;
  const int64_t n_items = (int64_t)a.B * Nq * a.H;
  const int64_t groups_per_grid = (int64_t)gridDim.x * (blockDim.x / LPG);
  const int64_t n_iter = (n_items + groups_per_grid - 1) / groups_per_grid;   // lock-step: every lane joins the shuffles
  int64_t item = (int64_t)blockIdx.x * (blockDim.x / LPG) + threadIdx.x / LPG;
  const int64_t map_floats = (int64_t)a.fH * a.fW * row;
  for (int64_t it = 0; it < n_iter; ++it, item += groups_per_grid) {
    const bool live = item < n_items;
    const int64_t item_c = live ? item : 0;
    const int h = (int)(item_c % a.H);
    const int64_t bq = item_c / a.H;
    const int b = (int)(bq / Nq), q = (int)(bq - (int64_t)b * Nq);
    const float* rowp = a.qproj + bq * a.ld;
    const float* offp = rowp + a.off_col + h * P * 2;
    float* grow = a.grad_qproj + bq * a.ld;
    float4 go = live ? ldg4(a.grad_out + bq * row + h * Dh + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    // softmax over the item's P logits (every lane of the group computes it: P <= 16 scalar loads served by one line)
    float aw[PP], g_a[PP];
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      aw[p] = p < P ? __ldg(rowp + a.logit_col + h * P + p) : -INFINITY;
      mx = fmaxf(mx, aw[p]);
    }
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      aw[p] = p < P ? expf(aw[p] - mx) : 0.f;
      sum += aw[p];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int p = 0; p < PP; ++p) aw[p] *= inv, g_a[p] = 0.f;

    unsigned hit = 1u;           // BEV mode: one "camera"
    if (CAM) {
      hit = 0u;
      int count = 0;
      for (int n = 0; n < a.N; ++n) {
        hit |= (__ldg(a.mask + (int64_t)q * a.N + n) != 0 ? 1u : 0u) << n;          // batch item 0 decides who contributes
        count += __ldg(a.mask + bq * a.N + n) != 0 ? 1 : 0;                           // the item's own visibility divides
      }
      go = scale4g(1.f / (float)max(count, 1), go);
    }
    if (!live) hit = 0u;
    const int64_t lane_off = h * Dh + lane * 4;

#pragma unroll
    for (int p = 0; p < PP; ++p) {
      float g_x = 0.f, g_y = 0.f;
      if (p < P) {
        const float ox = __ldg(offp + 2 * p), oy = __ldg(offp + 2 * p + 1);
        if (!CAM) {
          if (hit) {
            // pixel = ((q + .5) / bev + off / f) * f - .5  ==  (q + .5) * (f / bev) + off - .5   (as the forward kernel)
            const int64_t base = (int64_t)b * map_floats + lane_off;
            sample_bwd(a.value + base, a.grad_value + base, a.fH, a.fW, row, fmaf((float)(q / a.bev_w) + 0.5f, a.sy, oy - 0.5f),
                       fmaf((float)(q % a.bev_w) + 0.5f, a.sx, ox - 0.5f), aw[p], go, g_a[p], g_x, g_y);
          }
        } else {
          unsigned m = hit;
          while (m) {
            const int n = __ffs(m) - 1;
            m &= m - 1;
            const int64_t base = ((int64_t)b * a.N + n) * map_floats + lane_off;
            const float2 r = __ldg(reinterpret_cast<const float2*>(a.ref_cam) + (bq * a.N + n) * a.D + (p % a.D));
            sample_bwd(a.value + base, a.grad_value + base, a.fH, a.fW, row, fmaf(r.y, (float)a.fH, oy - 0.5f),
                       fmaf(r.x, (float)a.fW, ox - 0.5f), aw[p], go, g_a[p], g_x, g_y);
          }
        }
      }
      // fold the lanes' channel shares of this point's location gradient; one lane writes it
#pragma unroll
      for (int o = LPG / 2; o > 0; o >>= 1) {
        g_x += __shfl_xor_sync(0xffffffffu, g_x, o);
        g_y += __shfl_xor_sync(0xffffffffu, g_y, o);
      }
      if (live && p < P && lane == p % LPG) {
        grow[a.off_col + (h * P + p) * 2] = g_x;
        grow[a.off_col + (h * P + p) * 2 + 1] = g_y;
      }
    }
    // attention-weight gradients: fold, softmax backward, one lane per point writes
    float dot = 0.f;
#pragma unroll
    for (int p = 0; p < PP; ++p) {
#pragma unroll
      for (int o = LPG / 2; o > 0; o >>= 1) g_a[p] += __shfl_xor_sync(0xffffffffu, g_a[p], o);
      dot = fmaf(aw[p], g_a[p], dot);
    }
    if (live) {
#pragma unroll
      for (int p = 0; p < PP; ++p)
        if (p < P && lane == p % LPG) grow[a.logit_col + h * P + p] = aw[p] * (g_a[p] - dot);
    }
  }
}

template <bool CAM>
static int launch_sample_bwd(const char* fn, const SampleBwdArgs& a, int Dh, cudaStream_t s) {
  const int64_t items = (int64_t)a.B * a.bev_h * a.bev_w * a.H;
  const int lpg = Dh / 4;
  int64_t blocks = (items + 256 / lpg - 1) / (256 / lpg);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
#define UB_SB(LPG, PP) sample_bwd_kernel<LPG, PP, CAM><<<(int)blocks, 256, 0, s>>>(a)
#define UB_SB_P(LPG)        \
  do {                      \
    if (a.P <= 4)           \
      UB_SB(LPG, 4);        \
    else if (a.P <= 8)      \
      UB_SB(LPG, 8);        \
    else                    \
      UB_SB(LPG, 16);       \
  } while (0)
  switch (lpg) {
    case 1: UB_SB_P(1); break;
    case 2: UB_SB_P(2); break;
    case 4: UB_SB_P(4); break;
    case 8: UB_SB_P(8); break;
    case 16: UB_SB_P(16); break;
    default: UB_SB_P(32); break;
  }
#undef UB_SB_P
#undef UB_SB
  return check_launch(fn);
}

static int check_common(const char* fn, const void* value, const void* qproj, const void* grad_out, const void* grad_value,
                        const void* grad_qproj, int B, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P, int ld,
                        int off_col, int logit_col) {
  UB_REQUIRE(value && qproj && grad_out && grad_value && grad_qproj, "%s: null pointer", fn);
  UB_REQUIRE(B > 0 && bev_h > 0 && bev_w > 0 && fH > 0 && fW > 0 && H > 0 && Dh > 0 && P > 0, "%s: non-positive dimension", fn);
  UB_REQUIRE(off_col >= 0 && logit_col >= 0 && ld >= off_col + H * P * 2 && ld >= logit_col + H * P,
             "%s: qproj row stride %d too small", fn, ld);
  UB_REQUIRE_ALIGNED16(value);
  UB_REQUIRE_ALIGNED16(grad_out);
  UB_REQUIRE_ALIGNED16(grad_value);
  if (Dh % 4 != 0 || ((Dh / 4) & (Dh / 4 - 1)) != 0 || Dh > 128 || P > 16 || (int64_t)fH * fW * H * Dh >= (1ll << 31)) {
    set_error("%s: shape not covered (Dh=%d P=%d): head size must be 4 x a power of two <= 128, at most 16 points", fn, Dh, P);
    return ub::unsupported();
  }
  return UB_OK;
}

}  // namespace ub

using namespace ub;

extern "C" int ub_bev_sample_bwd(const float* value, const float* qproj, const float* grad_out, float* grad_value,
                                 float* grad_qproj, int B, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P, int ld,
                                 int off_col, int logit_col, ub_stream_t stream) {
  const char* fn = "ub_bev_sample_bwd";
  if (int rc = check_common(fn, value, qproj, grad_out, grad_value, grad_qproj, B, bev_h, bev_w, fH, fW, H, Dh, P, ld, off_col,
                            logit_col))
    return rc;
  SampleBwdArgs a = {};
  a.value = value, a.qproj = qproj, a.grad_out = grad_out, a.grad_value = grad_value, a.grad_qproj = grad_qproj;
  a.B = B, a.bev_h = bev_h, a.bev_w = bev_w, a.fH = fH, a.fW = fW, a.H = H, a.P = P, a.ld = ld;
  a.off_col = off_col, a.logit_col = logit_col;
  a.sx = (float)fW / (float)bev_w, a.sy = (float)fH / (float)bev_h;
  return launch_sample_bwd<false>(fn, a, Dh, (cudaStream_t)stream);
}

extern "C" int ub_img_sample_bwd(const float* value, const float* qproj, const float* ref_cam, const uint8_t* mask,
                                 const float* grad_out, float* grad_value, float* grad_qproj, int B, int N, int bev_h,
                                 int bev_w, int fH, int fW, int H, int Dh, int P, int D, int ld, int off_col, int logit_col,
                                 ub_stream_t stream) {
  const char* fn = "ub_img_sample_bwd";
  if (int rc = check_common(fn, value, qproj, grad_out, grad_value, grad_qproj, B, bev_h, bev_w, fH, fW, H, Dh, P, ld, off_col,
                            logit_col))
    return rc;
  UB_REQUIRE(ref_cam && mask, "%s: null pointer", fn);
  UB_REQUIRE(N > 0 && N <= 32 && D > 0 && D <= 8, "%s: bad dimension (N=%d D=%d)", fn, N, D);
  UB_REQUIRE(P % D == 0, "%s: num_points %d must be a multiple of the %d Z-anchors", fn, P, D);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(ref_cam) & 7u) == 0, "%s: ref_cam must be 8-byte aligned", fn);
  SampleBwdArgs a = {};
  a.value = value, a.qproj = qproj, a.grad_out = grad_out, a.grad_value = grad_value, a.grad_qproj = grad_qproj;
  a.ref_cam = ref_cam, a.mask = mask, a.N = N, a.D = D;
  a.B = B, a.bev_h = bev_h, a.bev_w = bev_w, a.fH = fH, a.fW = fW, a.H = H, a.P = P, a.ld = ld;
  a.off_col = off_col, a.logit_col = logit_col;
  a.sx = (float)fW / (float)bev_w, a.sy = (float)fH / (float)bev_h;
  return launch_sample_bwd<true>(fn, a, Dh, (cudaStream_t)stream);
}
