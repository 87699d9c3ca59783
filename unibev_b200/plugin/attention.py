"""Attention modules of the UniBEV BEV encoder -- plugin surface.

Same registered names, constructor keywords and state-dict keys as the reference
(spatial_cross_attention_img.py:23-66,218-291; spatial_cross_attention_pts.py:23-63,
209-282; mmcv MultiScaleDeformableAttention, verbatim copy at decoder.py:131-226).

Each ``forward`` here is the module-level, autograd-capable path: projections are
torch linears, the deformable sampling itself is ``MultiScaleDeformableAttnFunction``
(libunibev_b200 ``ub_msda_fwd`` / ``ub_msda_bwd``).  In eval mode
``UniBEVTransformer`` bypasses these forwards and drives the fused kernels directly
(``unibev_b200/plugin/fused.py``), reading the parameters held here.
"""
import math
import os
import warnings

import torch
import torch.nn as nn

import torch.nn.functional as F

from ..ops import (BevSampleFunction, ImgSampleFunction, LinearFunction, MultiScaleDeformableAttnFunction, add_identity,
                   fused_sample_supported, linear_train)
from .. import ops as _ops
from ..registry import ATTENTION, HAVE_MMCV, build_attention

# Module (autograd) path: when the caller is one of this package's encoders -- which pass the BEV grid shape
# (``ub_bev_grid``) and, for the cameras, the raw output of ``ub_project_points`` (``ub_cam``) down to the attentions -- the
# softmax / offset normalisation / reference-point add / (camera) rebatch-scatter-count glue and mmcv's op are replaced by
# ONE fused kernel per direction (``ub_bev_sample_fwd/bwd``, ``ub_img_sample_fwd/bwd``) that reads the raw linear outputs.
# UB_FUSED_TRAIN=0 (or setting this flag to False) keeps the op-level path (``ub_msda_fwd/bwd`` + torch glue).
FUSED_TRAIN_SAMPLING = os.environ.get('UB_FUSED_TRAIN', '1') == '1'


def _xavier_uniform(linear, bias=0.):
    nn.init.xavier_uniform_(linear.weight)
    nn.init.constant_(linear.bias, bias)


def _directional_offset_bias(num_heads, num_levels, num_points):
    """sampling_offsets bias: head h looks along angle 2*pi*h/heads, point i at radius i+1
    (spatial_cross_attention_img.py:296-307)."""
    theta = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
    d = torch.stack([theta.cos(), theta.sin()], -1)
    d = (d / d.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2).repeat(1, num_levels, num_points, 1)
    for i in range(num_points):
        d[:, :, i, :] *= i + 1
    return d.view(-1)


def _check_head_split(embed_dims, num_heads):
    if embed_dims % num_heads != 0:
        raise ValueError(f'embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}')
    dh = embed_dims // num_heads
    if dh & (dh - 1):
        warnings.warn('embed_dims / num_heads should be a power of 2: other head sizes take the scalar kernel path.')


class _DeformAttnBase(nn.Module):
    """Parameters + init shared by the three deformable attention flavours."""

    def __init__(self, embed_dims, num_heads, num_levels, num_points, im2col_step, batch_first, with_output_proj,
                 dropout):
        super().__init__()
        _check_head_split(embed_dims, num_heads)
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.im2col_step = im2col_step
        self.batch_first = batch_first
        self.fp16_enabled = False
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims) if with_output_proj else None
        self.dropout = nn.Dropout(dropout) if with_output_proj else None
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.)
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(_directional_offset_bias(self.num_heads, self.num_levels, self.num_points))
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)
        _xavier_uniform(self.value_proj)
        if self.output_proj is not None:
            _xavier_uniform(self.output_proj)
        self._is_init = True

    def _raw_rows(self, query):
        """(B, Nq, 3 H L P): raw sampling offsets (H, L, P, 2) then raw attention logits (H, L, P), one GEMM."""
        w = torch.cat((self.sampling_offsets.weight, self.attention_weights.weight), 0)
        b = torch.cat((self.sampling_offsets.bias, self.attention_weights.bias), 0)
        if _ops.TRAIN_KERNELS and query.is_cuda and query.dtype == torch.float32 and torch.is_grad_enabled():
            return LinearFunction.apply(query, w, b)
        return F.linear(query, w, b)

    def _fused_ok(self, key_padding_mask):
        return (FUSED_TRAIN_SAMPLING and self.num_levels == 1 and key_padding_mask is None
                and fused_sample_supported(self.embed_dims // self.num_heads, self.num_points))

    def _project(self, query, value, key_padding_mask):
        bs, nq, _ = query.shape
        _, nv, _ = value.shape
        v = self.value_proj(value)
        if key_padding_mask is not None:
            v = v.masked_fill(key_padding_mask[..., None], 0.0)
        v = v.view(bs, nv, self.num_heads, -1)
        off = self.sampling_offsets(query).view(bs, nq, self.num_heads, self.num_levels, self.num_points, 2)
        aw = self.attention_weights(query).view(bs, nq, self.num_heads, self.num_levels * self.num_points)
        aw = aw.softmax(-1).view(bs, nq, self.num_heads, self.num_levels, self.num_points)
        return v, off, aw


# With mmcv installed its own ``MultiScaleDeformableAttention`` already owns that registry name (and serves other models,
# CPU fallback included): take the name over only when the deployment asks for it, otherwise register ours
# as ``UBMultiScaleDeformableAttention`` and leave mmcv's entry alone.
_TAKE_OVER = HAVE_MMCV and os.environ.get('UNIBEV_B200_OVERRIDE_MMCV_MSDA') == '1'


@ATTENTION.register_module(name=None if (not HAVE_MMCV or _TAKE_OVER) else 'UBMultiScaleDeformableAttention', force=_TAKE_OVER)
class MultiScaleDeformableAttention(_DeformAttnBase):
    """mmcv's module of the same name (BEV self-attention of every encoder layer,
    config ``attn_cfgs[0]``), backed by libunibev_b200."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64, dropout=0.1,
                 batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__(embed_dims, num_heads, num_levels, num_points, im2col_step, batch_first, True, dropout)
        self.norm_cfg = norm_cfg

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
        grid = kwargs.get('ub_bev_grid')
        if (grid is not None and reference_points.shape[-1] == 2 and value.shape[1] == grid[0] * grid[1]
                and query.shape[1] == grid[0] * grid[1] and self._fused_ok(key_padding_mask)):
            # BEV self-attention over the query grid itself: reference points are the cell centres the kernel generates
            H, P = self.num_heads, self.num_points
            out = BevSampleFunction.apply(linear_train(self.value_proj, value), self._raw_rows(query), grid[0], grid[1],
                                          grid[0], grid[1], H, P, 0, 2 * H * P)
            out = linear_train(self.output_proj, out)
            if not self.batch_first:
                out = out.permute(1, 0, 2)
            return add_identity(out, identity, kwargs.get('ub_defer_add', False), self.dropout)
        assert int((spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum()) == value.shape[1]
        v, off, aw = self._project(query, value, key_padding_mask)
        if reference_points.shape[-1] == 2:
            norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            loc = reference_points[:, :, None, :, None, :2] + \
                off / self.num_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(f'Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.')
        out = MultiScaleDeformableAttnFunction.apply(v, spatial_shapes, level_start_index, loc, aw, self.im2col_step)
        out = self.output_proj(out)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return self.dropout(out) + identity


class _MSDeformableAttention3D(_DeformAttnBase):
    """Inner attention of the spatial cross-attentions: no output projection, reference
    points carry D Z-anchors and sampling point k uses anchor k % D
    (spatial_cross_attention_img.py:404-419)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=8, im2col_step=64, dropout=0.1,
                 batch_first=True, norm_cfg=None, init_cfg=None):
        super().__init__(embed_dims, num_heads, num_levels, num_points, im2col_step, batch_first, False, dropout)
        self.norm_cfg = norm_cfg

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        if value is None:
            value = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
        bs, nq, _ = query.shape
        grid, fhw = kwargs.get('ub_bev_grid'), kwargs.get('ub_value_hw')
        if (grid is not None and fhw is not None and nq == grid[0] * grid[1] and value.shape[1] == fhw[0] * fhw[1]
                and reference_points.shape[-1] == 2 and self._fused_ok(key_padding_mask)):
            # LiDAR cross-attention: every pillar anchor projects to its cell centre, which the kernel generates
            H, P = self.num_heads, self.num_points
            out = BevSampleFunction.apply(linear_train(self.value_proj, value), self._raw_rows(query), grid[0], grid[1],
                                          fhw[0], fhw[1], H, P, 0, 2 * H * P)
            return out if self.batch_first else out.permute(1, 0, 2)
        assert int((spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum()) == value.shape[1]
        v, off, aw = self._project(query, value, key_padding_mask)
        if reference_points.shape[-1] != 2:
            raise ValueError(f'Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.')
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        n_anchor = reference_points.shape[2]
        assert self.num_points % n_anchor == 0
        off = (off / norm[None, None, None, :, None, :]).view(bs, nq, self.num_heads, self.num_levels,
                                                              self.num_points // n_anchor, n_anchor, 2)
        loc = (reference_points[:, :, None, None, None, :, :] + off).view(bs, nq, self.num_heads, self.num_levels,
                                                                          self.num_points, 2)
        out = MultiScaleDeformableAttnFunction.apply(v, spatial_shapes, level_start_index, loc, aw, self.im2col_step)
        return out if self.batch_first else out.permute(1, 0, 2)


@ATTENTION.register_module()
class MSDeformableAttention3DImg(_MSDeformableAttention3D):
    pass


@ATTENTION.register_module()
class MSDeformableAttention3DPts(_MSDeformableAttention3D):
    pass


# unibev_nus_C.py:206 names this type although the reference never registers it
ATTENTION.register_module(name='MSDeformableAttention3DUniQueryImg', module=MSDeformableAttention3DImg)


@ATTENTION.register_module()
class SpatialCrossAttentionImg(nn.Module):
    """Camera cross-attention: every camera attends only the BEV queries that project into it."""

    def __init__(self, embed_dims=256, num_cams=6, pc_range=None, dropout=0.1, init_cfg=None, batch_first=False,
                 deformable_attention=dict(type='MSDeformableAttention3DImg', embed_dims=256, num_levels=4), **kwargs):
        super().__init__()
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        self.fp16_enabled = False
        self.deformable_attention = build_attention(deformable_attention)
        self.embed_dims, self.num_cams, self.batch_first = embed_dims, num_cams, batch_first
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weight()

    def init_weight(self):
        _xavier_uniform(self.output_proj)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None, reference_points=None,
                spatial_shapes=None, reference_points_cam=None, bev_mask=None, level_start_index=None, flag='encoder',
                **kwargs):
        """query (B, Nq, C); key/value (num_cams, sum(hw), B, C); reference_points_cam (num_cams, B, Nq, D, 2);
        bev_mask (num_cams, B, Nq, D) bool.  Hit lists come from batch item 0 and the divisor from each item's
        own mask, as in the reference (spatial_cross_attention_img.py:142, 209-212)."""
        if key is None:
            key = query
        if value is None:
            value = key
        if residual is not None:
            raise NotImplementedError('residual must be None (the reference only works with pre_norm=False)')
        inp_residual = query
        if query_pos is not None:
            query = query + query_pos
        B, Nq, C = query.shape
        N, D = reference_points_cam.size(0), reference_points_cam.size(3)
        cam, grid, fhw = kwargs.get('ub_cam'), kwargs.get('ub_bev_grid'), kwargs.get('ub_value_hw')
        da = self.deformable_attention
        if (cam is not None and grid is not None and fhw is not None and Nq == grid[0] * grid[1]
                and value.shape[1] == fhw[0] * fhw[1] and da.batch_first and da.num_points % D == 0
                and da._fused_ok(key_padding_mask)):
            # one fused kernel per direction instead of rebatch -> inner attention -> scatter -> count division
            H, P = da.num_heads, da.num_points
            v = linear_train(da.value_proj, value.permute(2, 0, 1, 3))   # (B, N, hw, C)
            slots = ImgSampleFunction.apply(v, da._raw_rows(query), cam[0], cam[1], grid[0], grid[1], fhw[0], fhw[1],
                                            H, P, 0, 2 * H * P)
            return add_identity(linear_train(self.output_proj, slots), inp_residual, kwargs.get('ub_defer_add', False),
                                self.dropout)
        hit0 = bev_mask[:, 0].any(-1)                                  # (N, Nq)
        lens = hit0.sum(1)
        max_len = int(lens.max())                                      # one host sync (training path only)
        if max_len == 0:                                               # no query sees any camera
            return self.dropout(self.output_proj(torch.zeros_like(query))) + inp_residual
        order = torch.sort((~hit0).to(torch.uint8), dim=1, stable=True)[1][:, :max_len]     # hit indexes first, ascending
        valid = torch.arange(max_len, device=query.device)[None] < lens[:, None]              # (N, max_len)
        q_re = query[:, order] * valid[None, :, :, None]                                      # (B, N, max_len, C)
        ref = reference_points_cam.permute(1, 0, 2, 3, 4)                                     # (B, N, Nq, D, 2)
        r_re = torch.gather(ref, 2, order[None, :, :, None, None].expand(B, N, max_len, D, 2)) * valid[None, :, :, None, None]
        l = value.shape[1]
        v = value.permute(2, 0, 1, 3).reshape(B * N, l, C)
        out = self.deformable_attention(query=q_re.reshape(B * N, max_len, C), key=v, value=v,
                                        reference_points=r_re.reshape(B * N, max_len, D, 2),
                                        spatial_shapes=spatial_shapes, level_start_index=level_start_index)
        out = out.view(B, N, max_len, C) * valid[None, :, :, None]
        slots = torch.zeros_like(query)
        for n in range(N):                                             # ascending camera order == reference sum order
            slots = slots.index_add(1, order[n], out[:, n])
        count = bev_mask.any(-1).permute(1, 2, 0).sum(-1).clamp(min=1.0)
        slots = self.output_proj(slots / count[..., None])
        return self.dropout(slots) + inp_residual


@ATTENTION.register_module()
class SpatialCrossAttentionPts(nn.Module):
    """LiDAR cross-attention: all BEV queries attend the one BEV-shaped LiDAR map."""

    def __init__(self, embed_dims=256, num_cams=6, pc_range=None, dropout=0.1, init_cfg=None, batch_first=False,
                 deformable_attention=None, **kwargs):
        super().__init__()
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        self.fp16_enabled = False
        self.deformable_attention = build_attention(deformable_attention)
        self.embed_dims, self.num_cams, self.batch_first = embed_dims, num_cams, batch_first
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        _xavier_uniform(self.output_proj)

    def forward(self, query, key, value, residual=None, query_pos=None, spatial_shapes=None,
                reference_points_lidar=None, bev_mask=None, level_start_index=None, **kwargs):
        """query (B, Nq, C); key/value (sum(hw), B, C); reference_points_lidar (D, B, Nq, 2)."""
        if key is None:
            key = query
        if value is None:
            value = key
        if residual is not None:
            raise NotImplementedError('residual must be None (the reference only works with pre_norm=False)')
        inp_residual = query
        if query_pos is not None:
            query = query + query_pos
        B, Nq, C = query.shape
        value = value.permute(1, 0, 2)
        out = self.deformable_attention(query=query, key=value, value=value,
                                        reference_points=reference_points_lidar.permute(1, 2, 0, 3),
                                        spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                                        ub_bev_grid=kwargs.get('ub_bev_grid'), ub_value_hw=kwargs.get('ub_value_hw'))
        out = linear_train(self.output_proj, out.view(B, -1, C))
        return add_identity(out, inp_residual, kwargs.get('ub_defer_add', False), self.dropout)
