"""Shared helpers for the test-suite (golden loading, synthetic inputs)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

ENCODER_HALF_TAGS = ['lc_cnw_linear', 'lc_cat', 'lc_avg_spatial', 'c_only', 'l_only_cnw',
                     'lc_cnw_dropflags', 'lc_cnw_dropdict']


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    arrays, params = {}, {}
    for k in z.files:
        v = z[k]
        t = torch.from_numpy(v) if v.dtype.kind in 'fiub' and v.ndim > 0 else v
        if k.startswith('p.'):
            params[k[2:]] = t
        else:
            arrays[k] = t
    return arrays, params


def metas_from(arrays):
    """img_metas list[dict] in the reference's format (encoder_unibev_detr_img.py:115-118,166-167)."""
    l2i = arrays['lidar2img'].numpy()
    h, w = (int(v) for v in arrays['img_hw'])
    return [dict(lidar2img=[l2i[b, n] for n in range(l2i.shape[1])], img_shape=[(h, w, 3)] * l2i.shape[1])
            for b in range(l2i.shape[0])]


def encoder_half_inputs(arrays):
    cfg = json.loads(str(arrays['cfg_json']))
    img = [arrays['img_feats']] if 'img_feats' in arrays else None
    pts = [arrays['pts_feats']] if 'pts_feats' in arrays else None
    if cfg.get('dual_queries'):
        q = [arrays['bev_queries_img'], arrays['bev_queries_pts']]
    else:
        q = arrays['bev_queries']
    bev_h, bev_w = (int(v) for v in arrays['bev_hw'])
    return cfg, img, pts, q, bev_h, bev_w
