"""Synthetic nuScenes-shaped workloads for tests and bench.py (no dataset, no checkpoint).

* ``transformer_cfg(...)``: the ``transformer=dict(type='UniBEVTransformer', ...)`` subtree of the
  reference configs (unibev_nus_LC_cnw_256_modality_dropout.py:255-292 etc.), decoder omitted.
* ``nominal_rig()``: a seeded, deterministic 6-camera pinhole rig with nuScenes-like yaw / focal
  lengths -> ``lidar2img`` matrices.
* ``make_inputs(...)``: backbone-output-shaped feature tensors, BEV queries, positional encoding.
* ``randomize_sampling_weights(model)``: default init zeroes the offset / attention linears
  (spatial_cross_attention_img.py:295,308), which would make every softmax uniform; redraw them so
  sampling is non-degenerate, keeping the directional offset bias.
"""
import math

import numpy as np
import torch

PC_RANGE = [-54, -54, -5, 54, 54, 3]


def transformer_cfg(embed_dims=256, fusion_method='linear', feature_norm='ChannelNormWeights', drop_modality=0.5,
                    num_layers=3, with_img=True, with_pts=True, cam_anchors=4, lidar_anchors=4, num_cams=6,
                    img_attn_type='MSDeformableAttention3DImg', dropout=0.1, **extra):
    C = embed_dims

    def layers(kind, inner):
        return dict(
            type=f'{kind}Layer',
            attn_cfgs=[dict(type='MultiScaleDeformableAttention', embed_dims=C, num_levels=1, dropout=dropout),
                       dict(type=f'SpatialCrossAttention{kind}', pc_range=PC_RANGE, num_cams=num_cams, dropout=dropout,
                            deformable_attention=dict(type=inner, embed_dims=C, num_points=8, num_levels=1),
                            embed_dims=C)],
            ffn_cfgs=dict(type='FFN', embed_dims=C), feedforward_channels=2 * C, ffn_dropout=dropout,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))

    cfg = dict(type='UniBEVTransformer', embed_dims=C, num_cams=num_cams, fusion_method=fusion_method,
               drop_modality=drop_modality, feature_norm=feature_norm, **extra)
    if with_img:
        cfg['img_encoder'] = dict(type='ImgEncoder', num_layers=num_layers, pc_range=PC_RANGE,
                                  num_points_in_pillar=cam_anchors, return_intermediate=False,
                                  transformerlayers=layers('Img', img_attn_type))
    if with_pts:
        cfg['pts_encoder'] = dict(type='PtsEncoder', num_layers=num_layers, pc_range=PC_RANGE,
                                  num_points_in_pillar_lidar=lidar_anchors, return_intermediate=False,
                                  transformerlayers=layers('Pts', 'MSDeformableAttention3DPts'))
    return cfg


# named workloads = BASELINE.json configs
WORKLOADS = {
    # unibev_nus_C.py: camera only, 6 x (3 x 256 x 704) -> stride-32 map 8 x 22, no CNW
    'unibev_nus_C': dict(cfg=dict(with_pts=False, feature_norm=None, drop_modality=None,
                                  img_attn_type='MSDeformableAttention3DUniQueryImg'),
                         img_hw=(256, 704), img_fhw=(8, 22), pts_fhw=None),
    'unibev_nus_L': dict(cfg=dict(with_img=False, feature_norm=None, drop_modality=None),
                         img_hw=None, img_fhw=None, pts_fhw=(180, 180)),
    'unibev_nus_LC_cnw_256': dict(cfg=dict(), img_hw=(928, 1600), img_fhw=(29, 50), pts_fhw=(180, 180)),
    'unibev_nus_LC_cat_128': dict(cfg=dict(embed_dims=128, fusion_method='cat', feature_norm=None),
                                  img_hw=(928, 1600), img_fhw=(29, 50), pts_fhw=(180, 180)),
}


def nominal_rig(img_hw=(928, 1600), num_cams=6):
    """lidar2img (num_cams, 4, 4) float64.  LiDAR frame: x forward, y left, z up; cameras 0.3 m below the
    LiDAR origin, 0.5 m out along their viewing direction.  Intrinsics are quoted at 1600 x 900 and scaled."""
    yaws = [0., -55., 55., 180., -110., 110.][:num_cams]
    sx = img_hw[1] / 1600.0
    mats = []
    for i, yaw in enumerate(yaws):
        f = (809.2 if i == 3 else 1266.4) * sx
        cx, cy = 816.3 * sx, 491.5 * sx
        a = math.radians(yaw)
        fwd = np.array([math.cos(a), math.sin(a), 0.0])
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        R = np.stack([right, -up, fwd], 0)                    # camera axes: x right, y down, z forward
        t = 0.5 * fwd + np.array([0.0, 0.0, -0.3])
        E = np.eye(4)
        E[:3, :3], E[:3, 3] = R, -R @ t
        K = np.eye(4)
        K[0, 0] = K[1, 1] = f
        K[0, 2], K[1, 2] = cx, cy
        mats.append(K @ E)
    return np.stack(mats, 0)


def img_metas(batch, img_hw=(928, 1600), num_cams=6):
    rig = nominal_rig(img_hw, num_cams)
    return [dict(lidar2img=[rig[n] for n in range(num_cams)], img_shape=[(img_hw[0], img_hw[1], 3)] * num_cams)
            for _ in range(batch)]


def make_inputs(workload, batch=1, bev_hw=(200, 200), seed=1, device='cpu', pin=False):
    """-> dict(img_feats, pts_feats, bev_queries, bev_pos, img_metas, bev_h, bev_w) for a named workload."""
    w = WORKLOADS[workload]
    C = w['cfg'].get('embed_dims', 256)
    g = torch.Generator().manual_seed(seed)

    def mk(*shape):
        t = torch.randn(*shape, generator=g)
        if pin:
            t = t.pin_memory()
        return t.to(device)
    out = dict(bev_h=bev_hw[0], bev_w=bev_hw[1], img_feats=None, pts_feats=None, img_metas=None)
    if w['img_fhw'] is not None:
        out['img_feats'] = [mk(batch, 6, C, *w['img_fhw'])]
        out['img_metas'] = img_metas(batch, w['img_hw'])
    if w['pts_fhw'] is not None:
        out['pts_feats'] = [mk(batch, C, *w['pts_fhw'])]
    out['bev_queries'] = mk(bev_hw[0] * bev_hw[1], C)
    out['bev_pos'] = mk(1, C, *bev_hw).expand(batch, -1, -1, -1).contiguous() if not pin else mk(batch, C, *bev_hw)
    return out


def randomize_sampling_weights(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('sampling_offsets.weight'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif name.endswith('attention_weights.weight'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith('attention_weights.bias'):
                p.copy_(torch.randn(p.shape, generator=g))
    return model


def build_model(workload, seed=0, **cfg_overrides):
    """UniBEVTransformer for a named workload: init_weights() under the seed, sampling linears re-drawn."""
    from .registry import build_transformer
    from . import plugin  # noqa: F401  (registers the modules)
    kw = dict(WORKLOADS[workload]['cfg'])
    kw.update(cfg_overrides)
    cfg = transformer_cfg(**kw)
    torch.manual_seed(seed)
    model = build_transformer(cfg)
    model.init_weights()
    randomize_sampling_weights(model, seed)
    return model, cfg


# LiDAR voxel layer of the L / LC configs (unibev_nus_LC_cnw_256_modality_dropout.py:186-190)
VOXEL_LAYER = dict(max_num_points=10, voxel_size=[0.075, 0.075, 0.2], max_voxels=(90000, 120000),
                   point_cloud_range=PC_RANGE)


def make_cloud(n=262144, seed=0, real_like=True):
    """Synthetic 10-sweep nuScenes-shaped cloud: (n, 5) fp32 = x, y, z, intensity, dt.
    ``real_like`` concentrates the points near the ego vehicle (range ~ exponential) so voxels hold several points, as
    sweeps do; otherwise x, y are uniform over +-54 m.  A few points fall outside the range on purpose."""
    g = np.random.default_rng(seed)
    if real_like:
        r = np.minimum(g.exponential(14.0, n) + 1.0, 60.0)
        th = g.uniform(0, 2 * math.pi, n)
        x, y = r * np.cos(th), r * np.sin(th)
        z = g.normal(-1.5, 0.9, n)
    else:
        x, y = g.uniform(-54, 54, n), g.uniform(-54, 54, n)
        z = g.uniform(-5, 3, n)
    inten = g.uniform(0, 255, n)
    dt = g.integers(0, 10, n) * 0.05
    return np.stack((x, y, z, inten, dt), 1).astype(np.float32)
