"""BEV encoders and their layers -- plugin surface.

``ImgEncoder`` / ``ImgLayer`` / ``PtsEncoder`` / ``PtsLayer`` keep the reference's
registered names, constructor keywords, state-dict keys and call conventions
(encoder_unibev_detr_img.py:18-43,190-201,292-357; encoder_unibev_detr_pts.py:18-43,
130-143,212-278).  Layer construction restates mmcv 1.3.17 ``BaseTransformerLayer`` /
``TransformerLayerSequence`` / ``FFN`` (attentions -> ffns -> norms; ``batch_first``
injected into every attention cfg; deprecated ``feedforward_channels`` /
``ffn_dropout`` / ``ffn_num_fcs`` folded into ``ffn_cfgs``).
"""
import copy

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..registry import (FEEDFORWARD_NETWORK, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, build_attention,
                        build_feedforward_network, build_transformer_layer)


def _ffn_registered():
    try:
        return FEEDFORWARD_NETWORK.get('FFN') is not None
    except Exception:  # noqa: BLE001
        return False


class FFN(nn.Module):
    """Linear-ReLU-Dropout-Linear-Dropout with identity shortcut; parameter names
    ``layers.0.0`` / ``layers.1`` as in mmcv."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        if num_fcs != 2 or act_cfg.get('type', 'ReLU') != 'ReLU':
            raise NotImplementedError('FFN: only num_fcs=2 with ReLU is used by the UniBEV configs')
        self.embed_dims, self.feedforward_channels, self.add_identity = embed_dims, feedforward_channels, add_identity
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))
        self.dropout_layer = nn.Identity()

    def forward(self, x, identity=None, defer=False):
        (lin1, _relu, drop1), lin2, drop2 = self.layers[0], self.layers[1], self.layers[2]
        out = ops.linear_train(lin2, drop1(ops.linear_train(lin1, x, relu=True)))          # drop2(out) == self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(drop2(out))
        if isinstance(self.dropout_layer, nn.Identity):
            return ops.add_identity(out, x if identity is None else identity, defer, drop2)
        return ops.add_identity(self.dropout_layer(drop2(out)), x if identity is None else identity, defer)


if not _ffn_registered():
    FEEDFORWARD_NETWORK.register_module(module=FFN)


class _EncoderLayer(nn.Module):
    _DEPRECATED = dict(feedforward_channels='feedforward_channels', ffn_dropout='ffn_drop', ffn_num_fcs='num_fcs')

    def __init__(self, attn_cfgs, feedforward_channels=None, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), batch_first=True, ffn_num_fcs=2,
                 ffn_cfgs=None, init_cfg=None, **kwargs):
        super().__init__()
        if ffn_cfgs is None:
            ffn_cfgs = dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0.,
                            act_cfg=dict(type='ReLU', inplace=True))
        ffn_cfgs = copy.deepcopy(ffn_cfgs)
        legacy = dict(feedforward_channels=feedforward_channels, ffn_dropout=ffn_dropout, ffn_num_fcs=ffn_num_fcs)
        for old, new in self._DEPRECATED.items():
            if legacy[old] is not None:
                ffn_cfgs[new] = legacy[old]
        if set(operation_order) - {'self_attn', 'norm', 'ffn', 'cross_attn'}:
            raise ValueError(f'operation_order {operation_order} has unknown entries')
        self.batch_first = batch_first
        self.operation_order = tuple(operation_order)
        self.norm_cfg = norm_cfg
        self.pre_norm = operation_order[0] == 'norm'
        self.num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(self.num_attn)]
        if self.num_attn != len(attn_cfgs):
            raise ValueError(f'{len(attn_cfgs)} attn_cfgs for {self.num_attn} attentions in {operation_order}')
        self.attentions = nn.ModuleList()
        for i, op in enumerate(o for o in operation_order if o in ('self_attn', 'cross_attn')):
            cfg = copy.deepcopy(dict(attn_cfgs[i]))
            if 'batch_first' in cfg:
                assert cfg['batch_first'] == batch_first
            else:
                cfg['batch_first'] = batch_first
            att = build_attention(cfg)
            att.operation_name = op
            self.attentions.append(att)
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = nn.ModuleList()
        for _ in range(operation_order.count('ffn')):
            cfg = copy.deepcopy(ffn_cfgs)
            if 'embed_dims' not in cfg:
                cfg['embed_dims'] = self.embed_dims
            else:
                assert cfg['embed_dims'] == self.embed_dims
            self.ffns.append(build_feedforward_network(cfg, dict(type='FFN')))
        if norm_cfg.get('type', 'LN') != 'LN':
            raise NotImplementedError('only LayerNorm norms are used by the UniBEV configs')
        self.norms = nn.ModuleList(nn.LayerNorm(self.embed_dims) for _ in range(operation_order.count('norm')))
        self.fp16_enabled = False

    def _forward(self, modality, query, key, value, bev_pos, query_pos, key_pos, ref_2d, ref_3d, bev_h, bev_w,
                 spatial_shapes, level_start_index, extra, **kwargs):
        """Operation dispatcher (encoder_unibev_detr_img.py:413-479): the self-attention gets
        ``query_pos=bev_pos``; a cross-attention that is not attentions[0] gets the caller's
        ``query_pos`` (None in the encoders)."""
        attn_i = norm_i = ffn_i = 0
        identity = query
        order = self.operation_order
        for pos, op in enumerate(order):
            # post-norm layers: an attention / FFN directly followed by a 'norm' hands over (out, identity) unadded and the
            # norm kernel adds while normalising (ops.Deferred; only this package's modules are asked to)
            defer = (not self.pre_norm) and pos + 1 < len(order) and order[pos + 1] == 'norm'
            if defer:
                kwargs = dict(kwargs, ub_defer_add=True)
            else:
                kwargs = {k: v for k, v in kwargs.items() if k != 'ub_defer_add'}
            if op == 'self_attn':
                query = self.attentions[attn_i](
                    query, query, query, identity if self.pre_norm else None, query_pos=bev_pos, key_pos=bev_pos,
                    reference_points=ref_2d, spatial_shapes=ops.const_tensor([[bev_h, bev_w]], torch.long, query.device),
                    level_start_index=ops.const_tensor([0], torch.long, query.device), **kwargs)
                attn_i += 1
                identity = query
            elif op == 'norm':
                query = ops.layer_norm_train(self.norms[norm_i], query)
                norm_i += 1
            elif op == 'cross_attn':
                qp, kp = (bev_pos, bev_pos) if (modality == 'img' and attn_i == 0) else (query_pos, key_pos)
                query = self.attentions[attn_i](
                    query, key, value, identity if self.pre_norm else None, query_pos=qp, key_pos=kp,
                    reference_points=ref_3d, spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                    **extra, **kwargs)
                attn_i += 1
                identity = query
            elif op == 'ffn':
                ffn = self.ffns[ffn_i]
                if isinstance(ffn, FFN):
                    query = ffn(query, identity if self.pre_norm else None, defer=defer)
                else:
                    query = ffn(query, identity if self.pre_norm else None)
                ffn_i += 1
        return query.materialize() if isinstance(query, ops.Deferred) else query


@TRANSFORMER_LAYER.register_module()
class ImgLayer(_EncoderLayer):
    def forward(self, query, key=None, value=None, bev_pos=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, ref_2d=None, ref_3d=None, bev_h=None, bev_w=None,
                reference_points_cam=None, mask=None, spatial_shapes=None, level_start_index=None, **kwargs):
        return self._forward('img', query, key, value, bev_pos, query_pos, key_pos, ref_2d, ref_3d, bev_h, bev_w,
                             spatial_shapes, level_start_index, dict(reference_points_cam=reference_points_cam),
                             **kwargs)


@TRANSFORMER_LAYER.register_module()
class PtsLayer(_EncoderLayer):
    def forward(self, query, key=None, value=None, bev_pos=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, ref_2d=None, ref_3d=None, bev_h=None, bev_w=None,
                reference_points_lidar=None, mask=None, spatial_shapes=None, level_start_index=None, **kwargs):
        return self._forward('pts', query, key, value, bev_pos, query_pos, key_pos, ref_2d, ref_3d, bev_h, bev_w,
                             spatial_shapes, level_start_index, dict(reference_points_lidar=reference_points_lidar),
                             **kwargs)


def anchor_heights(Z, D, dtype=torch.float32):
    """Normalised heights of the D pillar anchors (encoder_unibev_detr_img.py:68-69)."""
    return torch.linspace(0.5, Z - 0.5, D, dtype=dtype) / Z


class _Encoder(nn.Module):
    def __init__(self, transformerlayers=None, num_layers=None, pc_range=None, return_intermediate=False,
                 dataset_type='nuscenes', init_cfg=None):
        super().__init__()
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        if not isinstance(transformerlayers, (list, tuple)) or len(transformerlayers) != num_layers:
            raise ValueError('transformerlayers must be a dict or a list of num_layers dicts')
        self.num_layers = num_layers
        self.layers = nn.ModuleList(build_transformer_layer(c) for c in transformerlayers)
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm
        self.return_intermediate = return_intermediate
        self.pc_range = pc_range
        self.fp16_enabled = False

    @staticmethod
    def get_reference_points(H, W, Z=8, num_points_in_pillar=4, dim='3d', bs=1, device='cuda', dtype=torch.float):
        """3d: (bs, D, H*W, 3) normalised pillar points; 2d: (bs, H*W, 1, 2) cell centres."""
        xs = (torch.arange(W, device=device, dtype=dtype) + 0.5) / W
        ys = (torch.arange(H, device=device, dtype=dtype) + 0.5) / H
        gx, gy = xs.repeat(H), ys.repeat_interleave(W)
        if dim == '2d':
            return torch.stack((gx, gy), -1)[None, :, None, :].repeat(bs, 1, 1, 1)
        zs = anchor_heights(Z, num_points_in_pillar, dtype).to(device)
        D = num_points_in_pillar
        ref = torch.stack((gx[None].expand(D, -1), gy[None].expand(D, -1), zs[:, None].expand(-1, H * W)), -1)
        return ref[None].repeat(bs, 1, 1, 1)

    def _grids(self, bev_h, bev_w, Z, D, bs, device, dtype):
        """(ref_3d, ref_2d) of ``get_reference_points``: constants of the grid, built once per shape and device (they cost a
        dozen small kernels and one host -> device copy per forward otherwise)."""
        key = (bev_h, bev_w, float(Z), D, bs, str(device), dtype)
        cache = self.__dict__.setdefault('_grid_cache', {})
        if key not in cache:
            cache.clear()
            cache[key] = (self.get_reference_points(bev_h, bev_w, Z, D, dim='3d', bs=bs, device=device, dtype=dtype),
                          self.get_reference_points(bev_h, bev_w, dim='2d', bs=bs, device=device, dtype=dtype))
        return cache[key]

    def _run_layers(self, bev_query, key, value, args, kwargs, **layer_kw):
        output, inter = bev_query, []
        for layer in self.layers:
            output = layer(output, key, value, *args, **layer_kw, **kwargs)
            if self.return_intermediate:
                inter.append(output)
        return torch.stack(inter) if self.return_intermediate else output


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class ImgEncoder(_Encoder):
    def __init__(self, *args, pc_range=None, num_points_in_pillar=4, return_intermediate=False,
                 dataset_type='nuscenes', **kwargs):
        super().__init__(*args, pc_range=pc_range, return_intermediate=return_intermediate, **kwargs)
        self.num_points_in_pillar = num_points_in_pillar

    def point_sampling(self, reference_points, pc_range, img_metas, bev_hw=None):
        """-> reference_points_cam (num_cam, B, Nq, D, 2), bev_mask (num_cam, B, Nq, D) bool, computed by
        ``ub_project_points`` from the cell grid directly (``reference_points`` is only read for its shape;
        it must be the regular pillar grid of ``get_reference_points``)."""
        ref, mask = self._project_raw(reference_points, pc_range, img_metas, bev_hw)
        D = reference_points.shape[1]
        bits = (mask[..., None] >> torch.arange(D, device=mask.device, dtype=torch.uint8)) & 1
        return ref.permute(2, 0, 1, 3, 4), bits.bool().permute(2, 0, 1, 3)

    def _project_raw(self, reference_points, pc_range, img_metas, bev_hw=None, lidar2img=None, img_shape=None):
        """``ub_project_points`` output as the fused kernels take it: ref (B, Nq, N, D, 2), mask (B, Nq, N) uint8 bits."""
        B, D, Nq, _ = reference_points.shape
        if lidar2img is not None:
            l2i = lidar2img
        else:
            l2i = np.asarray([m['lidar2img'] for m in img_metas], dtype=np.float32)
            l2i = torch.from_numpy(l2i).to(reference_points.device)
        if bev_hw is None:      # the reference signature carries no grid shape: recover W from the first grid row
            W = int((reference_points[0, 0, :, 1] == reference_points[0, 0, 0, 1]).sum())
            bev_hw = (Nq // W, W)
        H, W = bev_hw
        zs = anchor_heights(pc_range[5] - pc_range[2], D).tolist()
        ih, iw = img_shape if img_shape is not None else img_metas[0]['img_shape'][0][:2]
        return ops.project_points(l2i, zs, pc_range, ih, iw, H, W)             # (B,Nq,N,D,2), (B,Nq,N) bits

    def forward(self, bev_query, key, value, *args, bev_h=None, bev_w=None, bev_pos=None, spatial_shapes=None,
                level_start_index=None, valid_ratios=None, **kwargs):
        """bev_query (Nq, B, C); key/value (num_cam, sum(hw), B, C) -> (B, Nq, C)."""
        bs, dev, dt = bev_query.size(1), bev_query.device, bev_query.dtype
        ref_3d, ref_2d = self._grids(bev_h, bev_w, self.pc_range[5] - self.pc_range[2], self.num_points_in_pillar, bs, dev, dt)
        # ``lidar2img`` (B, N, 4, 4) fp32 on the device + ``img_shape`` (h, w) replace the per-forward host -> device copy of
        # img_metas[i]['lidar2img'] (encoder_unibev_detr_img.py:115-124) when the caller stages calibration itself
        raw_ref, raw_mask = self._project_raw(ref_3d, self.pc_range, kwargs.get('img_metas'), (bev_h, bev_w),
                                              kwargs.get('lidar2img'), kwargs.get('img_shape'))
        D = self.num_points_in_pillar
        bits = (raw_mask[..., None] >> torch.arange(D, device=raw_mask.device, dtype=torch.uint8)) & 1
        ref_cam, bev_mask = raw_ref.permute(2, 0, 1, 3, 4), bits.bool().permute(2, 0, 1, 3)     # == point_sampling(...)
        bev_query = bev_query.permute(1, 0, 2).contiguous()      # (one copy here instead of one per consumer in layer 0)
        if bev_pos is not None:
            bev_pos = bev_pos.permute(1, 0, 2)
        return self._run_layers(bev_query, key, value, args, kwargs, bev_pos=bev_pos, ref_2d=ref_2d, ref_3d=ref_3d,
                                bev_h=bev_h, bev_w=bev_w, spatial_shapes=spatial_shapes,
                                level_start_index=level_start_index, reference_points_cam=ref_cam, bev_mask=bev_mask,
                                # for the fused sampling kernels of the module path (plugin/attention.py)
                                ub_bev_grid=(bev_h, bev_w), ub_cam=(raw_ref, raw_mask))


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class PtsEncoder(_Encoder):
    def __init__(self, *args, pc_range=None, num_points_in_pillar_lidar=1, return_intermediate=False,
                 dataset_type='nuscenes', **kwargs):
        super().__init__(*args, pc_range=pc_range, return_intermediate=return_intermediate, **kwargs)
        self.num_points_in_pillar_lidar = num_points_in_pillar_lidar

    def point_sampling(self, reference_points):
        """(B, D, Nq, 3) -> (D, B, Nq, 2) normalised xy + the in-range mask (B, Nq, D) the reference discards."""
        ref = reference_points.permute(1, 0, 2, 3)[..., :2]
        mask = ((ref[..., 1:2] > 0.0) & (ref[..., 1:2] < 1.0) & (ref[..., 0:1] < 1.0) & (ref[..., 0:1] > 0.0))
        return ref, mask.permute(1, 2, 0, 3).squeeze(-1)

    def forward(self, bev_query, key, value, *args, bev_h=None, bev_w=None, bev_pos=None, spatial_shapes=None,
                level_start_index=None, valid_ratios=None, prev_bev=None, shift=0., **kwargs):
        """bev_query (Nq, B, C); key/value (sum(hw), B, C) -> (B, Nq, C)."""
        bs, dev, dt = bev_query.size(1), bev_query.device, bev_query.dtype
        ref_3d, ref_2d = self._grids(bev_h, bev_w, self.pc_range[5] - self.pc_range[2], self.num_points_in_pillar_lidar, bs,
                                     dev, dt)
        ref_lidar, _ = self.point_sampling(ref_3d)
        bev_query = bev_query.permute(1, 0, 2).contiguous()      # (one copy here instead of one per consumer in layer 0)
        if bev_pos is not None:
            bev_pos = bev_pos.permute(1, 0, 2)
        return self._run_layers(bev_query, key, value, args, kwargs, bev_pos=bev_pos, ref_2d=ref_2d, ref_3d=ref_3d,
                                bev_h=bev_h, bev_w=bev_w, spatial_shapes=spatial_shapes,
                                level_start_index=level_start_index, reference_points_lidar=ref_lidar,
                                ub_bev_grid=(bev_h, bev_w))
