"""Restatement of the mmcv-full==1.3.17 pieces the UniBEV hot path leans on.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

mmcv is NOT vendored under /root/reference (docs/installation.md:6 pins
``mmcv-full==1.3.17``) and is not installable here, so the published algorithm is
restated.  Citations name the reference call site and the in-repo verbatim copy
of the mmcv code where one exists:

* ``msda_core``            <- mmcv.ops.multi_scale_deform_attn.multi_scale_deformable_attn_pytorch
                              (call sites: spatial_cross_attention_img.py:437-438,
                              spatial_cross_attention_pts.py:444-445, decoder.py:329-330)
* ``mmcv_msda_forward``    <- mmcv MultiScaleDeformableAttention.forward; verbatim in-repo copy at
                              projects/UniBEV/unibev_plugin/models/modules/decoder.py:230-338
* ``ffn_forward``          <- mmcv.cnn.bricks.transformer.FFN (Linear-ReLU-Drop-Linear-Drop, +identity)
* ``layer_norm``           <- nn.LayerNorm built by build_norm_layer(dict(type='LN'))
* ``learned_pos_encoding`` <- mmdet LearnedPositionalEncoding (unibev_head.py:180-182 is the caller)

All functions are eval-mode (every Dropout is the identity) and work on a flat
``params`` dict keyed by the reference's state-dict names.
"""
import torch
import torch.nn.functional as F


def msda_core(value, spatial_shapes, sampling_locations, attention_weights):
    """out[b,q,h*D+c] = sum_{l,p} w[b,q,h,l,p] * bilinear_zero_pad(value_l[b,:,h,c], loc[b,q,h,l,p])

    value (B, Nv, H, D); spatial_shapes list[(h, w)]; sampling_locations
    (B, Nq, H, L, P, 2) normalised (x, y); attention_weights (B, Nq, H, L, P).
    Pixel convention: x_pix = loc_x * W - 0.5 (grid_sample align_corners=False).
    """
    B, _, H, D = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    per_level = value.split([h * w for h, w in shapes], dim=1)
    grids = 2.0 * sampling_locations - 1.0
    sampled = []
    for lvl, (h, w) in enumerate(shapes):
        v = per_level[lvl].flatten(2).transpose(1, 2).reshape(B * H, D, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)          # (B*H, Nq, P, 2)
        sampled.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros',
                                     align_corners=False))              # (B*H, D, Nq, P)
    aw = attention_weights.transpose(1, 2).reshape(B * H, 1, Nq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1)
    return out.view(B, H * D, Nq).transpose(1, 2).contiguous()


def msda_core_scalar(value, spatial_shapes, sampling_locations, attention_weights):
    """Pure-python loop restatement of mmcv's CUDA kernel semantics
    (ms_deformable_im2col_gpu_kernel / ms_deform_attn_im2col_bilinear): each
    sample contributes iff h_im > -1 && w_im > -1 && h_im < H && w_im < W, and each
    of the four corners is bounds-checked on its own.  Small cases only.
    """
    import math
    B, _, H, D = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    out = torch.zeros(B, Nq, H * D, dtype=value.dtype)
    starts, acc = [], 0
    for h, w in spatial_shapes:
        starts.append(acc)
        acc += int(h) * int(w)
    for b in range(B):
        for q in range(Nq):
            for hd in range(H):
                tot = torch.zeros(D, dtype=value.dtype)
                for l, (hh, ww) in enumerate(spatial_shapes):
                    hh, ww = int(hh), int(ww)
                    for p in range(P):
                        x, y = sampling_locations[b, q, hd, l, p].tolist()
                        a = attention_weights[b, q, hd, l, p]
                        h_im = y * hh - 0.5
                        w_im = x * ww - 0.5
                        if not (h_im > -1 and w_im > -1 and h_im < hh and w_im < ww):
                            continue
                        h0, w0 = math.floor(h_im), math.floor(w_im)
                        lh, lw = h_im - h0, w_im - w0
                        for (yy, xx, wt) in ((h0, w0, (1 - lh) * (1 - lw)), (h0, w0 + 1, (1 - lh) * lw),
                                             (h0 + 1, w0, lh * (1 - lw)), (h0 + 1, w0 + 1, lh * lw)):
                            if 0 <= yy < hh and 0 <= xx < ww:
                                tot = tot + a * wt * value[b, starts[l] + yy * ww + xx, hd]
                out[b, q, hd * D:(hd + 1) * D] = tot
    return out


def linear(p, prefix, x):
    return F.linear(x, p[prefix + '.weight'], p[prefix + '.bias'])


def layer_norm(p, prefix, x, eps=1e-5):
    w = p[prefix + '.weight']
    return F.layer_norm(x, (w.numel(),), w, p[prefix + '.bias'], eps)


def ffn_forward(p, prefix, x, identity=None):
    """mmcv FFN with num_fcs=2, ReLU, add_identity=True; state-dict names
    ``layers.0.0`` (first Linear) and ``layers.1`` (second Linear)."""
    out = linear(p, prefix + '.layers.1', F.relu(linear(p, prefix + '.layers.0.0', x)))
    return (x if identity is None else identity) + out


def mmcv_msda_forward(p, prefix, query, value=None, identity=None, query_pos=None,
                      reference_points=None, spatial_shapes=None,
                      num_heads=8, num_levels=4, num_points=4, batch_first=False):
    """mmcv ``MultiScaleDeformableAttention.forward`` (decoder.py:230-338 is a
    verbatim in-repo copy).  Used as the BEV self-attention of every encoder
    layer (encoder_unibev_detr_img.py:417-430, config attn_cfgs[0])."""
    if value is None:
        value = query
    if identity is None:
        identity = query
    if query_pos is not None:
        query = query + query_pos
    if not batch_first:
        query = query.permute(1, 0, 2)
        value = value.permute(1, 0, 2)
    B, Nq, _ = query.shape
    _, Nv, _ = value.shape
    assert sum(int(h) * int(w) for h, w in spatial_shapes) == Nv
    value = linear(p, prefix + '.value_proj', value).view(B, Nv, num_heads, -1)
    off = linear(p, prefix + '.sampling_offsets', query).view(B, Nq, num_heads, num_levels, num_points, 2)
    aw = linear(p, prefix + '.attention_weights', query).view(B, Nq, num_heads, num_levels * num_points)
    aw = aw.softmax(-1).view(B, Nq, num_heads, num_levels, num_points)
    if reference_points.shape[-1] != 2:
        raise ValueError('oracle covers the 2-d reference-point form only')
    normalizer = torch.tensor([[float(w), float(h)] for h, w in spatial_shapes], dtype=query.dtype)
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    out = msda_core(value, spatial_shapes, loc, aw)
    out = linear(p, prefix + '.output_proj', out)
    if not batch_first:
        out = out.permute(1, 0, 2)
    return out + identity


def learned_pos_encoding(p, prefix, bs, h, w):
    """mmdet LearnedPositionalEncoding on an all-zero mask: (bs, 2*num_feats, h, w)."""
    col = p[prefix + '.col_embed.weight'][:w]
    row = p[prefix + '.row_embed.weight'][:h]
    pos = torch.cat((col.unsqueeze(0).repeat(h, 1, 1), row.unsqueeze(1).repeat(1, w, 1)), dim=-1)
    return pos.permute(2, 0, 1).unsqueeze(0).repeat(bs, 1, 1, 1)
