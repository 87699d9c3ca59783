"""LiDAR hard voxelisation behind the reference's interface.

``Voxelization`` mirrors the layer the reference builds from ``pts_voxel_layer=dict(max_num_points=10,
voxel_size=[0.075, 0.075, 0.2], max_voxels=(90000, 120000), point_cloud_range=...)``
(configs/unibev/unibev_nus_LC_cnw_256_modality_dropout.py:186-190; constructed by mmdet3d's
``MVXTwoStageDetector`` as ``Voxelization(**pts_voxel_layer)``): same constructor keywords, same call
(``layer(points) -> voxels, coors, num_points_per_voxel``), ``max_voxels[0]`` while training and ``[1]`` in eval mode.
``voxelize`` mirrors ``UniBEV.voxelize`` (unibev_detector.py:151-175) and ``HardSimpleVFE`` the voxel encoder of
:191-193.  The arithmetic is ``ub_hard_voxelize`` / ``ub_voxel_mean`` (csrc/voxelize.cu); there is no CPU path.
"""
import torch
from torch import nn

from .. import ops


class Voxelization(nn.Module):
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True):
        super().__init__()
        if max_num_points == -1 or max_voxels == -1:
            raise NotImplementedError('dynamic voxelisation (max_num_points=-1) is not on the UniBEV path')
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        if len(self.voxel_size) != 3 or len(self.point_cloud_range) != 6:
            raise ValueError('voxel_size needs 3 entries and point_cloud_range 6')
        self.max_num_points = int(max_num_points)
        self.max_voxels = tuple(max_voxels) if isinstance(max_voxels, (tuple, list)) else (int(max_voxels),) * 2
        self.deterministic = deterministic        # this implementation is always deterministic
        r, v = torch.tensor(self.point_cloud_range), torch.tensor(self.voxel_size)
        grid = torch.round((r[3:] - r[:3]) / v).long()
        self.grid_size = grid
        self.pcd_shape = [*grid[:2].tolist(), 1][::-1]

    def padded(self, points):
        """Device-resident result without a host sync: rows >= voxel_num[0] are padding."""
        mv = self.max_voxels[0] if self.training else self.max_voxels[1]
        return ops.hard_voxelize(points, self.voxel_size, self.point_cloud_range, self.max_num_points, mv)

    def forward(self, points):
        voxels, coors, num, voxel_num = self.padded(points)
        m = int(voxel_num.item())                 # the reference's return convention needs exact shapes
        return voxels[:m], coors[:m], num[:m]

    def __repr__(self):
        return (f'{type(self).__name__}(voxel_size={self.voxel_size}, point_cloud_range={self.point_cloud_range}, '
                f'max_num_points={self.max_num_points}, max_voxels={self.max_voxels})')


@torch.no_grad()
def voxelize(pts_voxel_layer, points):
    """``UniBEV.voxelize``: list of per-sample (N_i, C) clouds -> voxels (sum M, T, C), num_points (sum M),
    coors_batch (sum M, 4) int32 with the sample index in column 0."""
    voxels, coors, num_points = [], [], []
    for i, res in enumerate(points):
        v, c, n = pts_voxel_layer(res)
        voxels.append(v)
        num_points.append(n)
        coors.append(torch.nn.functional.pad(c, (1, 0), mode='constant', value=i))
    return torch.cat(voxels, 0), torch.cat(num_points, 0), torch.cat(coors, 0)


class HardSimpleVFE(nn.Module):
    """mmdet3d ``HardSimpleVFE(num_features=5)``: mean of the points of each voxel."""

    def __init__(self, num_features=4):
        super().__init__()
        self.num_features = num_features

    def forward(self, features, num_points, coors=None):
        return ops.voxel_mean(features, num_points.int(), self.num_features)
