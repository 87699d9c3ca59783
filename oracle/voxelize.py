"""CPU restatement of the LiDAR hard-voxelisation index path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference call site: ``UniBEV.voxelize`` (projects/UniBEV/unibev_plugin/models/detectors/unibev_detector.py:151-175)
loops ``self.pts_voxel_layer(res)`` over the samples of a batch and prepends the batch index to the voxel
coordinates (:170-174); the layer is ``Voxelization(max_num_points=10, voxel_size=[0.075, 0.075, 0.2],
point_cloud_range=[-54, -54, -5, 54, 54, 3], max_voxels=(90000, 120000))`` (configs/unibev/
unibev_nus_LC_cnw_256_modality_dropout.py:186-190), followed by ``HardSimpleVFE`` (:191-193).

The op itself lives in the un-vendored mmcv-full / mmdet3d (``hard_voxelize_forward``): parity UNPINNED by
reference tests.  Restated from its published CPU algorithm (``hard_voxelize_kernel``), which the deterministic
CUDA path reproduces:

    for each point, in order:
        c[j] = floor((p[j] - range_min[j]) / voxel_size[j])   j = x, y, z (fp32 arithmetic)
        skip the point if any c[j] < 0 or >= grid_size[j]
        voxel = (c_z, c_y, c_x); a new voxel gets the next id, unless max_voxels ids are taken (point dropped)
        append the point to its voxel unless the voxel already holds max_points points

with grid_size = round((range_max - range_min) / voxel_size).
"""
import numpy as np


def grid_size(voxel_size, pc_range):
    vs = np.asarray(voxel_size, dtype=np.float32)
    r = np.asarray(pc_range, dtype=np.float32)
    return np.round((r[3:] - r[:3]) / vs).astype(np.int64)          # (x, y, z)


def point_coors(points, voxel_size, pc_range):
    """-> (N, 3) int64 (x, y, z) cell of every point and (N,) validity, fp32 arithmetic like the op."""
    vs = np.asarray(voxel_size, dtype=np.float32)
    r = np.asarray(pc_range, dtype=np.float32)
    c = np.floor((points[:, :3].astype(np.float32) - r[:3]) / vs)
    g = grid_size(voxel_size, pc_range)
    ok = np.all((c >= 0) & (c < g), axis=1) & np.all(np.isfinite(c), axis=1)
    return np.where(ok[:, None], c, 0).astype(np.int64), ok


def hard_voxelize_loop(points, voxel_size, pc_range, max_points, max_voxels):
    """Literal per-point loop (small inputs only)."""
    N, C = points.shape
    coors_xyz, ok = point_coors(points, voxel_size, pc_range)
    table = {}
    voxels = np.zeros((max_voxels, max_points, C), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    n_vox = 0
    for i in range(N):
        if not ok[i]:
            continue
        key = (int(coors_xyz[i, 2]), int(coors_xyz[i, 1]), int(coors_xyz[i, 0]))     # (z, y, x)
        v = table.get(key, -1)
        if v == -1:
            if n_vox >= max_voxels:
                continue
            v = n_vox
            n_vox += 1
            table[key] = v
            coors[v] = key
        if num[v] < max_points:
            voxels[v, num[v]] = points[i]
            num[v] += 1
    return voxels[:n_vox], coors[:n_vox], num[:n_vox]


def hard_voxelize(points, voxel_size, pc_range, max_points, max_voxels):
    """Vectorised, same result as ``hard_voxelize_loop``."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    N, C = points.shape
    coors_xyz, ok = point_coors(points, voxel_size, pc_range)
    g = grid_size(voxel_size, pc_range)
    idx = np.nonzero(ok)[0]
    if idx.size == 0:
        return (np.zeros((0, max_points, C), np.float32), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32))
    c = coors_xyz[idx]
    key = (c[:, 2] * g[1] + c[:, 1]) * g[0] + c[:, 0]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')                  # voxels in order of first occurrence
    vid_of_uniq = np.empty_like(order)
    vid_of_uniq[order] = np.arange(order.size)
    vid = vid_of_uniq[inv]                                    # voxel id of every valid point
    # rank of the point inside its voxel (points are already in index order)
    srt = np.argsort(vid, kind='stable')
    vs_sorted = vid[srt]
    starts = np.r_[0, np.nonzero(np.diff(vs_sorted))[0] + 1]
    seg_start = np.repeat(starts, np.diff(np.r_[starts, vs_sorted.size]))
    rank = np.empty(vid.size, dtype=np.int64)
    rank[srt] = np.arange(vid.size) - seg_start
    n_vox = min(order.size, max_voxels)
    keep = (vid < max_voxels) & (rank < max_points)
    voxels = np.zeros((n_vox, max_points, C), dtype=np.float32)
    voxels[vid[keep], rank[keep]] = points[idx[keep]]
    num = np.bincount(vid[keep], minlength=n_vox).astype(np.int32)[:n_vox]
    fc = c[first[order[:n_vox]]]
    coors = np.stack((fc[:, 2], fc[:, 1], fc[:, 0]), 1).astype(np.int32)
    return voxels, coors, num


def voxelize_batch(points_list, voxel_size, pc_range, max_points, max_voxels):
    """``UniBEV.voxelize`` (unibev_detector.py:151-175): -> voxels (sum M, max_points, C), num_points (sum M),
    coors_batch (sum M, 4) with the sample index in column 0."""
    vs, cs, ns = [], [], []
    for b, pts in enumerate(points_list):
        v, c, n = hard_voxelize(pts, voxel_size, pc_range, max_points, max_voxels)
        vs.append(v)
        ns.append(n)
        cs.append(np.concatenate((np.full((c.shape[0], 1), b, np.int32), c), 1))
    return np.concatenate(vs, 0), np.concatenate(ns, 0), np.concatenate(cs, 0)


def hard_simple_vfe(voxels, num_points, num_features):
    """mmdet3d ``HardSimpleVFE``: mean of the points of a voxel over the first ``num_features`` channels."""
    s = voxels[:, :, :num_features].sum(1)
    return s / num_points.astype(np.float32)[:, None]
