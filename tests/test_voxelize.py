"""LiDAR hard voxelisation: CPU oracle self-checks and bit-exact GPU parity through the C ABI.

The op lives in un-vendored mmcv / mmdet3d, the reference ships no fixture for it: the oracle restates the published
sequential algorithm (oracle/voxelize.py) and is pinned here by a hand-computed known-answer case."""
import numpy as np
import pytest
import torch

from oracle import voxelize as ov
from unibev_b200 import synth

VS, RANGE = [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 4.0, 4.0, 2.0]
KAT_POINTS = np.array([[0.5, 0.5, 0.5, 10.0],     # cell x0 y0 z0 -> voxel 0
                       [3.9, 0.1, 1.2, 20.0],     # cell x3 y0 z1 -> voxel 1
                       [0.7, 0.2, 0.1, 30.0],     # voxel 0 again
                       [4.0, 1.0, 1.0, 40.0],     # x == range max -> cell 4 == grid size: dropped
                       [-0.1, 1.0, 1.0, 50.0],    # below range: dropped
                       [1.0, 2.0, 1.999, 60.0],   # cell x1 y2 z1 -> voxel 2
                       [0.1, 0.9, 0.9, 70.0]],    # voxel 0, third point
                      dtype=np.float32)


def test_oracle_known_answer():
    for fn in (ov.hard_voxelize_loop, ov.hard_voxelize):
        voxels, coors, num = fn(KAT_POINTS, VS, RANGE, max_points=2, max_voxels=10)
        assert coors.tolist() == [[0, 0, 0], [1, 0, 3], [1, 2, 1]]          # (z, y, x), first-occurrence order
        assert num.tolist() == [2, 1, 1]                                     # the third point of voxel 0 is dropped
        assert voxels[0, :, 3].tolist() == [10.0, 30.0] and voxels[1, :, 3].tolist() == [20.0, 0.0]
        assert voxels[2, 0].tolist() == KAT_POINTS[5].tolist()
        voxels, coors, num = fn(KAT_POINTS, VS, RANGE, max_points=5, max_voxels=2)
        assert coors.tolist() == [[0, 0, 0], [1, 0, 3]] and num.tolist() == [3, 1]   # voxel 2 dropped, voxel 0 still fills


@pytest.mark.parametrize('n,max_points,max_voxels', [(1, 3, 5), (500, 1, 10), (4000, 3, 60), (4000, 10, 100000)])
def test_oracle_vectorised_equals_loop(n, max_points, max_voxels):
    pts = synth.make_cloud(n, seed=n)
    vs, r = [0.8, 0.8, 1.0], [-20, -20, -5, 20, 20, 3]
    a = ov.hard_voxelize_loop(pts, vs, r, max_points, max_voxels)
    b = ov.hard_voxelize(pts, vs, r, max_points, max_voxels)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y)


def test_oracle_batch_and_mean():
    clouds = [synth.make_cloud(700, seed=s) for s in (1, 2)]
    vs, r = [2.0, 2.0, 4.0], [-20, -20, -5, 20, 20, 3]
    voxels, num, coors = ov.voxelize_batch(clouds, vs, r, 4, 1000)
    assert coors.shape[1] == 4 and set(coors[:, 0].tolist()) == {0, 1}
    assert voxels.shape[0] == num.shape[0] == coors.shape[0] and num.min() >= 1 and num.max() <= 4
    mean = ov.hard_simple_vfe(voxels, num, 5)
    k = int(np.argmax(num))
    np.testing.assert_allclose(mean[k], voxels[k, :num[k]].mean(0), rtol=1e-6)


# ------------------------------------------------------------------------------------------------ GPU parity
def _gpu(points, vs, r, max_points, max_voxels):
    from unibev_b200 import ops
    voxels, coors, num, m = ops.hard_voxelize(torch.from_numpy(points).cuda(), vs, r, max_points, max_voxels)
    m = int(m.item())
    assert (num[m:] == 0).all() and (voxels[m:] == 0).all()          # padding rows stay empty
    return voxels[:m].cpu().numpy(), coors[:m].cpu().numpy(), num[:m].cpu().numpy()


def _same(a, b):
    for x, y in zip(a, b):
        assert x.shape == y.shape and x.dtype == y.dtype
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_gpu_known_answer():
    _same(_gpu(KAT_POINTS, VS, RANGE, 2, 10), ov.hard_voxelize_loop(KAT_POINTS, VS, RANGE, 2, 10))
    _same(_gpu(KAT_POINTS, VS, RANGE, 5, 2), ov.hard_voxelize_loop(KAT_POINTS, VS, RANGE, 5, 2))


@pytest.mark.gpu
@pytest.mark.parametrize('n,max_points,max_voxels,real_like', [
    (1, 10, 100, True), (33, 1, 7, True), (5000, 3, 200, True), (5000, 10, 100000, False), (70001, 10, 2000, True)])
def test_gpu_bit_exact_small(n, max_points, max_voxels, real_like):
    pts = synth.make_cloud(n, seed=n, real_like=real_like)
    vs, r = [0.5, 0.5, 1.0], [-30, -30, -5, 30, 30, 3]
    _same(_gpu(pts, vs, r, max_points, max_voxels), ov.hard_voxelize(pts, vs, r, max_points, max_voxels))


@pytest.mark.gpu
def test_gpu_edge_cases():
    vs, r = synth.VOXEL_LAYER['voxel_size'], synth.VOXEL_LAYER['point_cloud_range']
    outside = np.full((100, 5), 1e3, np.float32)                      # nothing in range -> no voxels
    v, c, n = _gpu(outside, vs, r, 10, 50)
    assert v.shape == (0, 10, 5) and c.shape == (0, 3) and n.shape == (0,)
    weird = synth.make_cloud(2000, seed=5)
    weird[::13, 0] = np.nan
    weird[5::17, 1] = np.inf
    weird[7::19, 2] = -np.inf
    _same(_gpu(weird, vs, r, 10, 5000), ov.hard_voxelize(weird, vs, r, 10, 5000))
    same_cell = np.tile(np.array([[0.01, 0.01, 0.0, 1.0, 0.0]], np.float32), (300, 1))
    same_cell[:, 3] = np.arange(300)
    v, c, n = _gpu(same_cell, vs, r, 10, 50)
    assert n.tolist() == [10] and v[0, :, 3].tolist() == list(range(10))   # first ten points, in order
    # points exactly on cell borders / the range limits
    border = np.array([[-54.0, -54.0, -5.0, 0, 0], [54.0, 0, 0, 0, 0], [53.999996, 53.999996, 2.9999998, 0, 0],
                       [0.075, 0.15, 0.2, 0, 0], [-0.075, -0.15, -0.2, 0, 0]], np.float32)
    _same(_gpu(border, vs, r, 10, 50), ov.hard_voxelize_loop(border, vs, r, 10, 50))


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['train', 'test'])
def test_gpu_full_size_config2(mode):
    """BASELINE configs[1] cloud: 262 144 points, 0.075 m voxels, nuScenes range; bit-exact against the vectorised
    oracle plus size-independent properties."""
    L = synth.VOXEL_LAYER
    mv = L['max_voxels'][0 if mode == 'train' else 1]
    pts = synth.make_cloud(262144, seed=0)
    got = _gpu(pts, L['voxel_size'], L['point_cloud_range'], L['max_num_points'], mv)
    _same(got, ov.hard_voxelize(pts, L['voxel_size'], L['point_cloud_range'], L['max_num_points'], mv))
    voxels, coors, num = got
    assert coors.shape[0] <= mv and num.min() >= 1 and num.max() <= L['max_num_points']
    keys = (coors[:, 0].astype(np.int64) * 1440 + coors[:, 1]) * 1440 + coors[:, 2]
    assert np.unique(keys).size == keys.size                         # every voxel appears once
    # every kept point lies inside its voxel
    k = np.arange(voxels.shape[1])[None, :] < num[:, None]
    cell = np.floor((voxels[..., :3] - np.float32([-54, -54, -5])) / np.float32(L['voxel_size']))
    assert np.array_equal(cell[k][:, ::-1].astype(np.int32), np.repeat(coors, num, axis=0))


@pytest.mark.gpu
def test_gpu_plugin_voxelize_and_vfe():
    from unibev_b200.plugin.voxelize import HardSimpleVFE, Voxelization, voxelize
    L = synth.VOXEL_LAYER
    layer = Voxelization(**L).eval()
    clouds = [synth.make_cloud(30000 + 1000 * s, seed=10 + s) for s in range(3)]
    voxels, num, coors = voxelize(layer, [torch.from_numpy(c).cuda() for c in clouds])
    w_voxels, w_num, w_coors = ov.voxelize_batch(clouds, L['voxel_size'], L['point_cloud_range'], L['max_num_points'],
                                                 L['max_voxels'][1])
    assert coors.dtype == torch.int32 and np.array_equal(coors.cpu().numpy(), w_coors)
    assert np.array_equal(num.cpu().numpy(), w_num) and np.array_equal(voxels.cpu().numpy(), w_voxels)
    mean = HardSimpleVFE(num_features=5)(voxels, num, coors)
    np.testing.assert_allclose(mean.cpu().numpy(), ov.hard_simple_vfe(w_voxels, w_num, 5), rtol=1e-6, atol=1e-6)
    layer.train()
    assert layer(torch.from_numpy(clouds[0]).cuda())[1].shape[0] <= L['max_voxels'][0]
    with pytest.raises(RuntimeError):
        layer(torch.from_numpy(clouds[0]))                           # no CPU path
