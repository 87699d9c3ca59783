// LiDAR hard voxelisation, the index path upstream of the LiDAR backbone                                  [R8]
// (UniBEV.voxelize, unibev_detector.py:151-175 -> mmcv / mmdet3d hard_voxelize_forward).
//
// Deterministic and bit-exact with the sequential CPU algorithm (voxel ids in order of first occurrence, the
// first max_points points of a voxel in point order, voxels beyond max_voxels dropped), but data-parallel:
//   1. key(i) = linear cell index of point i (fp32 floor((p - min) / size), like the op) or INVALID
//   2. radix sort of (key << 32 | i): the points of a voxel become one run, ordered by point index
//      (cub::DeviceRadixSort, the CUDA toolkit's sort primitive)
//   3. run heads -> first point of every voxel; exclusive scan of the first-point flags in point order
//      = voxel id in order of first occurrence; inclusive max-scan of the head positions = run start of every
//      sorted element (rank inside the voxel = position - run start)
//   4. scatter the kept points into voxels / coors / num_points_per_voxel
// No atomics, no dense 1440 x 1440 x 41 lookup table (the op's CPU path allocates one per call).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ub_common.cuh"

namespace ub {

constexpr uint32_t kInvalidKey = 0x7FFFFFFFu;

struct VoxParams {
  float vx, vy, vz, x0, y0, z0;
  int gx, gy, gz;
};

__global__ void __launch_bounds__(256) vox_keys_kernel(const float* __restrict__ pts, int N, int C, VoxParams p,
                                                       unsigned long long* __restrict__ keys, int* __restrict__ first_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float x = pts[(int64_t)i * C], y = pts[(int64_t)i * C + 1], z = pts[(int64_t)i * C + 2];
  const float fx = floorf(__fdiv_rn(__fsub_rn(x, p.x0), p.vx));
  const float fy = floorf(__fdiv_rn(__fsub_rn(y, p.y0), p.vy));
  const float fz = floorf(__fdiv_rn(__fsub_rn(z, p.z0), p.vz));
  const bool ok = fx >= 0.f && fx < (float)p.gx && fy >= 0.f && fy < (float)p.gy && fz >= 0.f && fz < (float)p.gz;
  uint32_t key = kInvalidKey;
  if (ok) key = (uint32_t)(((int)fz * p.gy + (int)fy) * p.gx + (int)fx);
  keys[i] = ((unsigned long long)key << 32) | (unsigned)i;
  first_flag[i] = 0;
}

// sorted element j: head of a run? -> mark its point as the first of a voxel, remember the run start
__global__ void __launch_bounds__(256) vox_heads_kernel(const unsigned long long* __restrict__ sorted, int N,
                                                        int* __restrict__ first_flag, int* __restrict__ run_start) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const uint32_t key = (uint32_t)(sorted[j] >> 32);
  const bool head = key != kInvalidKey && (j == 0 || (uint32_t)(sorted[j - 1] >> 32) != key);
  run_start[j] = head ? j : 0;
  if (head) first_flag[(uint32_t)sorted[j]] = 1;
}

struct MaxOp {
  __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

__global__ void __launch_bounds__(256)
    vox_assign_kernel(const float* __restrict__ pts, const unsigned long long* __restrict__ sorted,
                      const int* __restrict__ run_start, const int* __restrict__ vid_of_point,
                      const int* __restrict__ first_flag, int N, int C, VoxParams p, int max_points, int max_voxels,
                      float* __restrict__ voxels, int* __restrict__ coors, int* __restrict__ num_points,
                      int* __restrict__ voxel_num) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) {   // number of voxels = number of first points (capped)
    const int total = vid_of_point[N - 1] + first_flag[N - 1];
    *voxel_num = total < max_voxels ? total : max_voxels;
  }
  if (j >= N) return;
  const uint32_t key = (uint32_t)(sorted[j] >> 32);
  if (key == kInvalidKey) return;
  const int i = (int)(uint32_t)sorted[j];
  const int start = run_start[j], rank = j - start;
  const int vid = vid_of_point[(uint32_t)sorted[start]];
  if (vid >= max_voxels) return;
  if (rank < max_points) {
    float* dst = voxels + ((int64_t)vid * max_points + rank) * C;
    for (int c = 0; c < C; ++c) dst[c] = pts[(int64_t)i * C + c];
  }
  if (rank == 0) {
    const int x = key % p.gx, y = (key / p.gx) % p.gy, z = key / (p.gx * p.gy);
    coors[vid * 3] = z, coors[vid * 3 + 1] = y, coors[vid * 3 + 2] = x;
  }
  const bool last = j + 1 == N || (uint32_t)(sorted[j + 1] >> 32) != key;
  if (last) num_points[vid] = rank + 1 < max_points ? rank + 1 : max_points;
}

// mean of the points of every voxel (HardSimpleVFE)
__global__ void __launch_bounds__(256) vox_mean_kernel(const float* __restrict__ voxels, const int* __restrict__ num_points,
                                                       int M, int max_points, int C, int F, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * F) return;
  const int v = t / F, f = t % F;
  float s = 0.f;
  for (int k = 0; k < max_points; ++k) s += voxels[((int64_t)v * max_points + k) * C + f];   // empty slots hold zeros
  out[t] = s / (float)num_points[v];
}

struct VoxWorkspace {
  size_t keys, sorted, first, vid, run, cub, total;
};
static VoxWorkspace vox_layout(int N) {
  VoxWorkspace w;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t sort_b = 0, scan_b = 0, scan2_b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, N);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (int*)nullptr, (int*)nullptr, N);
  cub::DeviceScan::InclusiveScan(nullptr, scan2_b, (int*)nullptr, (int*)nullptr, MaxOp(), N);
  size_t cub_b = sort_b > scan_b ? sort_b : scan_b;
  if (scan2_b > cub_b) cub_b = scan2_b;
  size_t off = 0;
  w.keys = off, off += al((size_t)N * 8);
  w.sorted = off, off += al((size_t)N * 8);
  w.first = off, off += al((size_t)N * 4);
  w.vid = off, off += al((size_t)N * 4);
  w.run = off, off += al((size_t)N * 4);
  w.cub = off, off += al(cub_b);
  w.total = off;
  return w;
}

}  // namespace ub

using namespace ub;

extern "C" int ub_voxelize_workspace_bytes(int num_points, size_t* bytes) {
  UB_REQUIRE(num_points > 0 && bytes, "ub_voxelize_workspace_bytes: need num_points > 0 and an output pointer");
  *bytes = vox_layout(num_points).total;
  return UB_OK;
}

extern "C" int ub_hard_voxelize(const float* points, int N, int C, const float* voxel_size_host,
                                const float* pc_range_host, int max_points, int max_voxels, float* voxels, int* coors,
                                int* num_points_per_voxel, int* voxel_num, void* workspace, size_t workspace_bytes,
                                ub_stream_t stream) {
  const char* fn = "ub_hard_voxelize";
  UB_REQUIRE(points && voxel_size_host && pc_range_host && voxels && coors && num_points_per_voxel && voxel_num && workspace,
             "%s: null pointer", fn);
  UB_REQUIRE(N > 0 && C >= 3 && max_points > 0 && max_voxels > 0, "%s: need N>0, C>=3, max_points>0, max_voxels>0", fn);
  VoxParams p;
  p.vx = voxel_size_host[0], p.vy = voxel_size_host[1], p.vz = voxel_size_host[2];
  p.x0 = pc_range_host[0], p.y0 = pc_range_host[1], p.z0 = pc_range_host[2];
  UB_REQUIRE(p.vx > 0.f && p.vy > 0.f && p.vz > 0.f, "%s: voxel sizes must be positive", fn);
  // grid_size = round((max - min) / size) in fp32, like the op's Python wrapper
  p.gx = (int)nearbyintf((pc_range_host[3] - pc_range_host[0]) / p.vx);
  p.gy = (int)nearbyintf((pc_range_host[4] - pc_range_host[1]) / p.vy);
  p.gz = (int)nearbyintf((pc_range_host[5] - pc_range_host[2]) / p.vz);
  UB_REQUIRE(p.gx > 0 && p.gy > 0 && p.gz > 0 && (int64_t)p.gx * p.gy * p.gz < (int64_t)kInvalidKey,
             "%s: grid %d x %d x %d unsupported", fn, p.gx, p.gy, p.gz);
  const VoxWorkspace w = vox_layout(N);
  UB_REQUIRE(workspace_bytes >= w.total, "%s: workspace too small (%zu < %zu bytes)", fn, workspace_bytes, w.total);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "%s: workspace must be 256-byte aligned", fn);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  auto* keys = reinterpret_cast<unsigned long long*>(ws + w.keys);
  auto* sorted = reinterpret_cast<unsigned long long*>(ws + w.sorted);
  int* first = reinterpret_cast<int*>(ws + w.first);
  int* vid = reinterpret_cast<int*>(ws + w.vid);
  int* run = reinterpret_cast<int*>(ws + w.run);
  size_t cub_b = w.total - w.cub;
  const int blocks = (N + 255) / 256;
  cudaMemsetAsync(voxels, 0, (size_t)max_voxels * max_points * C * sizeof(float), s);
  cudaMemsetAsync(num_points_per_voxel, 0, (size_t)max_voxels * sizeof(int), s);
  vox_keys_kernel<<<blocks, 256, 0, s>>>(points, N, C, p, keys, first);
  cub::DeviceRadixSort::SortKeys(ws + w.cub, cub_b, keys, sorted, N, 0, 63, s);
  vox_heads_kernel<<<blocks, 256, 0, s>>>(sorted, N, first, run);
  cub::DeviceScan::ExclusiveSum(ws + w.cub, cub_b, first, vid, N, s);
  cub::DeviceScan::InclusiveScan(ws + w.cub, cub_b, run, run, MaxOp(), N, s);
  vox_assign_kernel<<<blocks, 256, 0, s>>>(points, sorted, run, vid, first, N, C, p, max_points, max_voxels, voxels, coors,
                                           num_points_per_voxel, voxel_num);
  return check_launch(fn);
}

extern "C" int ub_voxel_mean(const float* voxels, const int* num_points_per_voxel, int M, int max_points, int C,
                             int num_features, float* out, ub_stream_t stream) {
  UB_REQUIRE(voxels && num_points_per_voxel && out, "ub_voxel_mean: null pointer");
  UB_REQUIRE(M >= 0 && max_points > 0 && C > 0 && num_features > 0 && num_features <= C, "ub_voxel_mean: bad shape");
  if (M == 0) return UB_OK;
  vox_mean_kernel<<<(M * num_features + 255) / 256, 256, 0, (cudaStream_t)stream>>>(voxels, num_points_per_voxel, M,
                                                                                    max_points, C, num_features, out);
  return check_launch("ub_voxel_mean");
}
