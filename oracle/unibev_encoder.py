"""Functional CPU restatement of the UniBEV uniform-BEV-encoder hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Eval mode: every Dropout
is the identity and modality dropout is off unless flags are forced.

Citations are relative to /root/reference/projects/UniBEV/unibev_plugin/models/modules/:
  fusion   = transformer_fusion.py         enc_img = encoder_unibev_detr_img.py
  sca_img  = spatial_cross_attention_img.py enc_pts = encoder_unibev_detr_pts.py
  sca_pts  = spatial_cross_attention_pts.py

``params`` is a flat dict of tensors keyed exactly like
``UniBEVTransformer.state_dict()`` (e.g.
``img_bev_encoder.layers.0.attentions.1.deformable_attention.value_proj.weight``);
``cfg`` is the ``transformer=dict(type='UniBEVTransformer', ...)`` subtree of a
reference config, unchanged.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .mmcv_semantics import (ffn_forward, layer_norm, linear, mmcv_msda_forward,
                             msda_core)


# --------------------------------------------------------------------------- #
# reference points                                                            #
# --------------------------------------------------------------------------- #
def pillar_points_3d(H, W, Z, D, bs, dtype=torch.float32):
    """enc_img:67-95 / enc_pts:48-88 -> (bs, D, H*W, 3), query index q = h*W + w,
    last dim (x, y, z) all normalised to (0, 1)."""
    zs = torch.linspace(0.5, Z - 0.5, D, dtype=dtype).view(D, 1, 1).expand(D, H, W) / Z
    xs = torch.linspace(0.5, W - 0.5, W, dtype=dtype).view(1, 1, W).expand(D, H, W) / W
    ys = torch.linspace(0.5, H - 0.5, H, dtype=dtype).view(1, H, 1).expand(D, H, W) / H
    ref = torch.stack((xs, ys, zs), -1).reshape(D, H * W, 3)
    return ref[None].repeat(bs, 1, 1, 1)


def grid_points_2d(H, W, bs, dtype=torch.float32):
    """enc_img:98-109 -> (bs, H*W, 1, 2) pixel centres, (x, y) normalised."""
    ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, dtype=dtype),
                            torch.linspace(0.5, W - 0.5, W, dtype=dtype), indexing='ij')
    ref = torch.stack((xs.reshape(-1) / W, ys.reshape(-1) / H), -1)
    return ref[None].repeat(bs, 1, 1).unsqueeze(2)


def project_to_cameras(ref_3d, pc_range, img_metas):
    """enc_img:112-187 ``ImgEncoder.point_sampling``.

    ref_3d (B, D, Nq, 3) -> ref_cam (num_cam, B, Nq, D, 2) normalised image (x, y),
    mask (num_cam, B, Nq, D) bool.  Points behind a camera keep z := 1e-5 and are
    only masked (enc_img:150-164)."""
    lidar2img = np.asarray([m['lidar2img'] for m in img_metas])            # (B, N, 4, 4)
    lidar2img = ref_3d.new_tensor(lidar2img)
    pts = ref_3d.clone()
    pts[..., 0:1] = pts[..., 0:1] * (pc_range[3] - pc_range[0]) + pc_range[0]
    pts[..., 1:2] = pts[..., 1:2] * (pc_range[4] - pc_range[1]) + pc_range[1]
    pts[..., 2:3] = pts[..., 2:3] * (pc_range[5] - pc_range[2]) + pc_range[2]
    pts = torch.cat((pts, torch.ones_like(pts[..., :1])), -1).permute(1, 0, 2, 3)   # (D, B, Nq, 4)
    D, B, Nq = pts.shape[:3]
    N = lidar2img.size(1)
    pts = pts.view(D, B, 1, Nq, 4).repeat(1, 1, N, 1, 1).unsqueeze(-1)
    mats = lidar2img.view(1, B, N, 1, 4, 4).repeat(D, 1, 1, Nq, 1, 1)
    cam = torch.matmul(mats.to(torch.float32), pts.to(torch.float32)).squeeze(-1)   # (D, B, N, Nq, 4)
    eps = 1e-5
    mask = cam[..., 2:3] > eps
    cam = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.ones_like(cam[..., 2:3]) * eps)
    cam[..., 0] /= img_metas[0]['img_shape'][0][1]
    cam[..., 1] /= img_metas[0]['img_shape'][0][0]
    mask = (mask & (cam[..., 1:2] > 0.0) & (cam[..., 1:2] < 1.0)
            & (cam[..., 0:1] < 1.0) & (cam[..., 0:1] > 0.0))
    mask = torch.nan_to_num(mask)
    return cam.permute(2, 1, 3, 0, 4), mask.permute(2, 1, 3, 0, 4).squeeze(-1)


def project_to_lidar(ref_3d):
    """enc_pts:106-127: keep normalised (x, y) only -> (D, B, Nq, 2); the mask the
    reference computes is discarded by its caller (enc_pts:169)."""
    return ref_3d.clone().permute(1, 0, 2, 3)[..., :2]


# --------------------------------------------------------------------------- #
# attention modules                                                           #
# --------------------------------------------------------------------------- #
def _attn_dims(acfg, default_points):
    return (acfg.get('num_heads', 8), acfg.get('num_levels', 4), acfg.get('num_points', default_points))


def msda3d_forward(p, prefix, acfg, query, value, reference_points, spatial_shapes):
    """sca_img:313-442 / sca_pts:306-449 ``MSDeformableAttention3D{Img,Pts}.forward``
    (batch_first=True, no output_proj): point k of the L*P sampling points uses
    Z-anchor k % D because of the ``view(..., P // D, D, 2)`` (sca_img:412-419)."""
    H, L, P = _attn_dims(acfg, 8)
    B, Nq, _ = query.shape
    _, Nv, _ = value.shape
    assert sum(int(h) * int(w) for h, w in spatial_shapes) == Nv
    value = linear(p, prefix + '.value_proj', value).view(B, Nv, H, -1)
    off = linear(p, prefix + '.sampling_offsets', query).view(B, Nq, H, L, P, 2)
    aw = linear(p, prefix + '.attention_weights', query).view(B, Nq, H, L * P).softmax(-1).view(B, Nq, H, L, P)
    if reference_points.shape[-1] != 2:
        raise ValueError('Last dim of reference_points must be 2')
    normalizer = torch.tensor([[float(w), float(h)] for h, w in spatial_shapes], dtype=query.dtype)
    D = reference_points.shape[2]
    ref = reference_points[:, :, None, None, None, :, :]
    off = (off / normalizer[None, None, None, :, None, :]).view(B, Nq, H, L, P // D, D, 2)
    loc = (ref + off).view(B, Nq, H, L, P, 2)
    return msda_core(value, spatial_shapes, loc, aw)


def sca_img_forward(p, prefix, acfg, query, key, value, ref_cam, bev_mask, spatial_shapes,
                    query_pos=None):
    """sca_img:67-215 ``SpatialCrossAttentionImg.forward`` incl. its quirks: hit
    indexes come from batch item 0's mask (sca_img:142) while the divisor counts
    cameras per batch item (sca_img:209-212); padded rebatch rows are computed and
    dropped; the residual is the query before ``query_pos`` (sca_img:120-124)."""
    residual = query
    slots = torch.zeros_like(query)
    if query_pos is not None:
        query = query + query_pos
    B, Nq, C = query.shape
    D = ref_cam.size(3)
    N = ref_cam.size(0)
    hit = [bev_mask[i][0].sum(-1).nonzero().squeeze(-1) for i in range(N)]
    max_len = max(len(h) for h in hit)
    q_re = query.new_zeros(B, N, max_len, C)
    r_re = ref_cam.new_zeros(B, N, max_len, D, 2)
    for j in range(B):
        for i in range(N):
            q_re[j, i, :len(hit[i])] = query[j, hit[i]]
            r_re[j, i, :len(hit[i])] = ref_cam[i][j, hit[i]]
    n_cam, l, bs, _ = key.shape
    value = value.permute(2, 0, 1, 3).reshape(bs * N, l, C)
    out = msda3d_forward(p, prefix + '.deformable_attention', acfg['deformable_attention'],
                         q_re.view(B * N, max_len, C), value,
                         r_re.view(B * N, max_len, D, 2), spatial_shapes).view(B, N, max_len, C)
    for j in range(B):
        for i in range(N):
            slots[j, hit[i]] += out[j, i, :len(hit[i])]
    count = (bev_mask.sum(-1) > 0).permute(1, 2, 0).sum(-1)
    count = torch.clamp(count, min=1.0)
    slots = slots / count[..., None]
    return linear(p, prefix + '.output_proj', slots) + residual


def sca_pts_forward(p, prefix, acfg, query, key, value, ref_lidar, spatial_shapes, query_pos=None):
    """sca_pts:65-206 ``SpatialCrossAttentionPts.forward``: every BEV query attends
    the one LiDAR BEV map; no rebatch, no mask, no count."""
    residual = query
    if query_pos is not None:
        query = query + query_pos
    B, Nq, C = query.shape
    value = value.permute(1, 0, 2)
    ref = ref_lidar.permute(1, 2, 0, 3)                                    # (B, Nq, D, 2)
    out = msda3d_forward(p, prefix + '.deformable_attention', acfg['deformable_attention'],
                         query, value, ref, spatial_shapes).view(B, -1, C)
    return linear(p, prefix + '.output_proj', out) + residual


# --------------------------------------------------------------------------- #
# encoder layers / encoders                                                   #
# --------------------------------------------------------------------------- #
def layer_forward(p, prefix, lcfg, modality, query, key, value, bev_pos, ref_2d, bev_h, bev_w,
                  spatial_shapes, ref_cam=None, bev_mask=None, ref_lidar=None):
    """enc_img:339-481 ``ImgLayer.forward`` / enc_pts:256-355 ``PtsLayer.forward`` with
    pre_norm=False: self-attn gets ``query_pos=bev_pos``; the cross-attn is
    attentions[1], so it gets ``query_pos=None`` (enc_img:457-463, enc_pts:337)."""
    attn_i = norm_i = ffn_i = 0
    acfgs = lcfg['attn_cfgs']
    for op in lcfg['operation_order']:
        if op == 'self_attn':
            H, L, P = _attn_dims(acfgs[attn_i], 4)
            query = mmcv_msda_forward(p, f'{prefix}.attentions.{attn_i}', query, query, None,
                                      query_pos=bev_pos, reference_points=ref_2d,
                                      spatial_shapes=[(bev_h, bev_w)], num_heads=H, num_levels=L,
                                      num_points=P, batch_first=True)
            attn_i += 1
        elif op == 'norm':
            query = layer_norm(p, f'{prefix}.norms.{norm_i}', query)
            norm_i += 1
        elif op == 'cross_attn':
            qp = bev_pos if attn_i == 0 and modality == 'img' else None
            if modality == 'img':
                query = sca_img_forward(p, f'{prefix}.attentions.{attn_i}', acfgs[attn_i], query, key, value,
                                        ref_cam, bev_mask, spatial_shapes, query_pos=qp)
            else:
                query = sca_pts_forward(p, f'{prefix}.attentions.{attn_i}', acfgs[attn_i], query, key, value,
                                        ref_lidar, spatial_shapes, query_pos=qp)
            attn_i += 1
        elif op == 'ffn':
            query = ffn_forward(p, f'{prefix}.ffns.{ffn_i}', query)
            ffn_i += 1
    return query


def encoder_forward(p, prefix, ecfg, modality, bev_query, feats, bev_h, bev_w, bev_pos,
                    spatial_shapes, img_metas=None):
    """enc_img:189-289 ``ImgEncoder.forward`` / enc_pts:129-209 ``PtsEncoder.forward``.
    bev_query (Nq, B, C) -> (B, Nq, C)."""
    bs = bev_query.size(1)
    pc = ecfg['pc_range']
    D = ecfg.get('num_points_in_pillar', 4) if modality == 'img' else ecfg.get('num_points_in_pillar_lidar', 1)
    ref_3d = pillar_points_3d(bev_h, bev_w, pc[5] - pc[2], D, bs, bev_query.dtype)
    ref_2d = grid_points_2d(bev_h, bev_w, bs, bev_query.dtype)
    extra = {}
    if modality == 'img':
        extra['ref_cam'], extra['bev_mask'] = project_to_cameras(ref_3d, pc, img_metas)
    else:
        extra['ref_lidar'] = project_to_lidar(ref_3d)
    q = bev_query.permute(1, 0, 2)
    pos = bev_pos.permute(1, 0, 2) if bev_pos is not None else None
    for lid in range(ecfg['num_layers']):
        q = layer_forward(p, f'{prefix}.layers.{lid}', ecfg['transformerlayers'], modality, q, feats, feats,
                          pos, ref_2d, bev_h, bev_w, spatial_shapes, **extra)
    return q


# --------------------------------------------------------------------------- #
# UniBEVTransformer, encoder half                                             #
# --------------------------------------------------------------------------- #
def flatten_img_feats(p, cfg, mlvl):
    """fusion:231-255 -> (num_cam, sum(hw), B, C), shapes list[(h, w)]."""
    flat, shapes = [], []
    for lvl, feat in enumerate(mlvl):
        bs, n, c, h, w = feat.shape
        f = feat.flatten(3).permute(1, 0, 3, 2)
        if cfg.get('use_cams_embeds', True):
            f = f + p['cams_embeds'][:, None, None, :]
        f = f + p['img_level_embeds'][None, None, lvl:lvl + 1, :]
        shapes.append((h, w))
        flat.append(f)
    return torch.cat(flat, 2).permute(0, 2, 1, 3), shapes


def flatten_pts_feats(p, cfg, mlvl):
    """fusion:257-278 -> (sum(hw), B, C).  Levels are concatenated on the CHANNEL
    axis there (fusion:272) -- only meaningful for one level."""
    flat, shapes = [], []
    for lvl, feat in enumerate(mlvl):
        bs, c, h, w = feat.shape
        f = feat.flatten(2).permute(0, 2, 1) + p['pts_level_embeds'][None, lvl:lvl + 1, :]
        shapes.append((h, w))
        flat.append(f)
    return torch.cat(flat, 2).permute(1, 0, 2), shapes


def channel_norm_weights(p, cfg, img, pts, c_flag, l_flag):
    """fusion:316-337 CNW: per-channel softmax over the two modality weights when
    both are present, softmax over a single row (== 1) otherwise; a missing
    modality becomes zeros."""
    if img is None:
        img = torch.zeros_like(pts)
    elif pts is None:
        pts = torch.zeros_like(img)
    if cfg.get('feature_norm') == 'ChannelNormWeights':
        w = torch.stack((p['img_channel_weights'], p['pts_channel_weights']), 0)
        if c_flag == 1 and l_flag == 1:
            n = w.softmax(0)
            wi, wp = n[0], n[1]
        else:
            wi, wp = w[0:1].softmax(0)[0], w[1:2].softmax(0)[0]
        img, pts = img * wi, pts * wp
    elif cfg.get('feature_norm') in _MLP_ACTS:          # fusion:345-366
        w = F.linear(torch.cat([img, pts], dim=1).permute(0, 2, 1), p['channel_weights_proj.0.weight'],
                     p['channel_weights_proj.0.bias'])
        w = _MLP_ACTS[cfg['feature_norm']](w)            # (bs, C, 2)
        if c_flag == 1 and l_flag == 1:
            n = F.softmax(w, dim=-1)
            wi, wp = n[:, :, 0], n[:, :, 1]
        else:
            wi, wp = F.softmax(w[:, :, :1], dim=-1).squeeze(-1), F.softmax(w[:, :, 1:], dim=-1).squeeze(-1)
        img, pts = img * wi[:, None, :], pts * wp[:, None, :]
    elif cfg.get('feature_norm') == 'ModalityProjection':   # fusion:26-47, 375-381
        def proj(prefix, x):
            h = F.relu(F.linear(x, p[prefix + '.net.0.weight'], p[prefix + '.net.0.bias']))
            w = p[prefix + '.net.2.weight']
            return x + F.layer_norm(h, (w.numel(),), w, p[prefix + '.net.2.bias'], 1e-5)
        img, pts = torch.cat([img, proj('l_modal_proj', img)], -1), torch.cat([proj('c_modal_proj', pts), pts], -1)
    elif cfg.get('feature_norm') is not None:
        raise NotImplementedError(cfg['feature_norm'])
    return img, pts


_MLP_ACTS = {'MLP_ChannelNormWeights': F.relu, 'Leaky_ReLU_MLP_ChannelNormWeights': F.leaky_relu,
             'ELU_MLP_ChannelNormWeights': F.elu, 'Sigmoid_MLP_ChannelNormWeights': torch.sigmoid}


def spatial_norm_weights(p, cfg, img, pts, c_flag, l_flag):
    """fusion:386-413."""
    if cfg.get('spatial_norm') == 'SpatialNormWeights':
        w = torch.stack((p['img_spatial_weights'], p['pts_spatial_weights']), 0)
        if c_flag == 1 and l_flag == 1:
            n = w.softmax(0)
            wi, wp = n[0], n[1]
        else:
            wi, wp = w[:1].softmax(0)[0], w[1:].softmax(0)[0]
        img, pts = img * wi[None, :, None], pts * wp[None, :, None]
    return img, pts


def fuse(p, cfg, img, pts, c_flag, l_flag):
    """fusion:280-314."""
    m = cfg.get('fusion_method', 'linear')
    if m == 'linear':
        out = c_flag * img + l_flag * pts
    elif m == 'avg':
        out = img * c_flag / (c_flag + l_flag) + pts * l_flag / (c_flag + l_flag)
    elif m == 'cat' and cfg.get('feature_norm') == 'ModalityProjection':    # fusion:287-300
        C = img.shape[-1] // 2
        img_flags = torch.cat((torch.full((C,), float(c_flag)), torch.full((C,), float(1 - l_flag))))
        pts_flags = torch.cat((torch.full((C,), float(1 - c_flag)), torch.full((C,), float(l_flag))))
        out = img * img_flags + pts * pts_flags
    elif m == 'cat':
        out = torch.cat((img * c_flag, pts * l_flag), -1)
    else:
        raise NotImplementedError(m)
    if cfg.get('use_modal_embeds') == 'MLP':            # fusion:304-307
        s = torch.tensor([float(c_flag), float(l_flag)])
        e = F.relu(F.linear(F.relu(F.linear(s, p['modal_embbeding_mlp.0.weight'], p['modal_embbeding_mlp.0.bias'])),
                            p['modal_embbeding_mlp.2.weight'], p['modal_embbeding_mlp.2.bias']))
        out = out + e
    elif cfg.get('use_modal_embeds') == 'Fixed':
        out = out + c_flag * p['modal_embbeding_C'] + l_flag * p['modal_embbeding_L']
    elif cfg.get('use_modal_embeds') is not None:
        raise NotImplementedError(cfg['use_modal_embeds'])
    return out


def encoder_half(p, cfg, img_mlvl_feats, pts_mlvl_feats, bev_queries, bev_h, bev_w, bev_pos=None,
                 img_metas=None, flags=None, return_parts=False):
    """fusion:463-538: everything ``UniBEVTransformer.forward`` does before the
    object-query decoder.  Returns fused_bev_embed (B, bev_h*bev_w, C*scale).
    ``flags=(c_flag, l_flag)`` forces a modality-dropout draw (fusion:474-477)."""
    c_flag, l_flag = (1, 1) if flags is None else flags
    if img_mlvl_feats is None:
        c_flag, bs = 0, pts_mlvl_feats[0].size(0)
    elif pts_mlvl_feats is None:
        l_flag, bs = 0, img_mlvl_feats[0].size(0)
    else:
        bs = img_mlvl_feats[0].size(0)
    if bev_pos is not None:
        bev_pos = bev_pos.flatten(2).permute(2, 0, 1)
    if cfg.get('dual_queries', False):
        q_img = bev_queries[0].unsqueeze(1).repeat(1, bs, 1)
        q_pts = bev_queries[1].unsqueeze(1).repeat(1, bs, 1)
    else:
        q_img = q_pts = bev_queries.unsqueeze(1).repeat(1, bs, 1)
    img = pts = None
    if img_mlvl_feats is not None:
        feats, shapes = flatten_img_feats(p, cfg, img_mlvl_feats)
        img = encoder_forward(p, 'img_bev_encoder', cfg['img_encoder'], 'img', q_img, feats, bev_h, bev_w,
                              bev_pos, shapes, img_metas)
    if pts_mlvl_feats is not None:
        feats, shapes = flatten_pts_feats(p, cfg, pts_mlvl_feats)
        pts = encoder_forward(p, 'pts_bev_encoder', cfg['pts_encoder'], 'pts', q_pts, feats, bev_h, bev_w,
                              bev_pos, shapes)
    parts = (img, pts)
    img, pts = channel_norm_weights(p, cfg, img, pts, c_flag, l_flag)
    img, pts = spatial_norm_weights(p, cfg, img, pts, c_flag, l_flag)
    fused = fuse(p, cfg, img, pts, c_flag, l_flag)
    return (fused, parts) if return_parts else fused
