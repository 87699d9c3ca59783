// Pillar reference points -> camera planes + visibility bits                                   [R2]
// (ImgEncoder.get_reference_points + point_sampling, encoder_unibev_detr_img.py:45-187).
#include "ub_common.cuh"

namespace ub {

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

}  // namespace ub

using namespace ub;

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
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  project_points_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lidar2img, pp, ref_cam, mask, B, N, bev_h, bev_w, D);
  return check_launch("ub_project_points");
}

