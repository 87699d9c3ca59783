"""Shared helpers for the test-suite (golden loading, synthetic inputs)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

ENCODER_HALF_TAGS = ['lc_cnw_linear', 'lc_cat', 'lc_avg_spatial', 'c_only', 'l_only_cnw',
                     'lc_cnw_dropflags', 'lc_cnw_dropdict', 'lc_mlp_cnw', 'lc_sigmoid_mlp_cnw_dropflags', 'lc_modproj_cat',
                     'lc_modproj_cat_dropflags', 'lc_cnw_modal_mlp']


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


def msda_quantised(value, shape, loc, aw, value_dtype=torch.float16, weight_dtype=torch.float16):
    """Single-level MSDA with explicit 4-corner gathers (mmcv kernel semantics: zero padding, each corner
    bounds-checked), emulating the window kernels' storage: value rounded to ``value_dtype``, the combined
    attention x bilinear weight rounded to ``weight_dtype``, products accumulated in fp32.
    value (B, Nv, H, D); loc (B, Nq, H, P, 2) normalised; aw (B, Nq, H, P) -> (B, Nq, H*D)."""
    B, Nv, H, D = value.shape
    _, Nq, _, P, _ = loc.shape
    h, w = shape
    v = value.to(value_dtype).float() if value_dtype is not None else value
    x = loc[..., 0] * w - 0.5
    y = loc[..., 1] * h - 0.5
    inside = (y > -1) & (x > -1) & (y < h) & (x < w)
    x0, y0 = torch.floor(x), torch.floor(y)
    lx, ly = x - x0, y - y0
    out = torch.zeros(B, Nq, H, D)
    bi = torch.arange(B).view(B, 1, 1, 1).expand(B, Nq, H, P)
    hi = torch.arange(H).view(1, 1, H, 1).expand(B, Nq, H, P)
    for dy, dx, wt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
        yy, xx = (y0 + dy), (x0 + dx)
        ok = inside & (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).long()
        wq = aw * wt
        if weight_dtype is not None:
            wq = wq.to(weight_dtype).float()
        wq = torch.where(ok, wq, torch.zeros_like(wq))
        out += (v[bi, idx, hi] * wq.unsqueeze(-1)).sum(3)
    return out.reshape(B, Nq, H * D)


def bev_loc_weights(qproj, bev_h, bev_w, fH, fW, H, P):
    """Raw offset|logit rows (B, Nq, H*P*3) -> sampling locations (B, Nq, H, P, 2) and softmax weights
    (B, Nq, H, P) of the BEV-grid attentions (reference point = cell centre; mmcv offset normaliser (W, H))."""
    B, Nq, _ = qproj.shape
    off = qproj[..., :H * P * 2].reshape(B, Nq, H, P, 2)
    aw = qproj[..., H * P * 2:].reshape(B, Nq, H, P).softmax(-1)
    ys, xs = torch.meshgrid(torch.arange(bev_h, dtype=torch.float32), torch.arange(bev_w, dtype=torch.float32),
                            indexing='ij')
    ref = torch.stack(((xs.reshape(-1) + 0.5) / bev_w, (ys.reshape(-1) + 0.5) / bev_h), -1)
    loc = ref.view(1, Nq, 1, 1, 2) + off / torch.tensor([float(fW), float(fH)])
    return loc, aw
