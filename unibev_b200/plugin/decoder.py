"""Object-query decoder of UniBEV: the immediate consumer of ``fused_bev_embed``.

Registered under the reference's names so the ``decoder=dict(type='DetectionTransformerDecoder', ...)`` subtree of the
configs (unibev_nus_LC_cnw_256_modality_dropout.py:325-352) builds unchanged:

* ``DetectionTransformerDecoder`` (TRANSFORMER_LAYER_SEQUENCE) -- decoder.py:51-128: layer loop with iterative
  reference-point refinement through the head's ``reg_branches`` (``inverse_sigmoid``, decoder.py:33-48);
* ``CustomMSDeformableAttention`` (ATTENTION) -- decoder.py:131-338: deformable cross-attention of the 900 object
  queries into the 200 x 200 BEV map; the sampling is ``ub_msda_fwd`` / ``ub_msda_bwd`` (Nq = 900, Nv = 40 000);
* ``DetrTransformerDecoderLayer`` (TRANSFORMER_LAYER) and ``MultiheadAttention`` (ATTENTION): the mmcv 1.3.17 classes the
  config names (un-vendored dependency; restated: ``BaseTransformerLayer`` operation dispatch, ``nn.MultiheadAttention``
  with ``query_pos`` added to query and key, dropout + identity).

State-dict keys equal the reference's (``decoder.layers.{i}.attentions.0.attn.in_proj_weight``,
``...attentions.1.sampling_offsets.weight``, ``...ffns.0.layers.0.0.weight``, ``...norms.{k}.weight``).
"""
import copy

import torch
import torch.nn as nn

from ..registry import (ATTENTION, HAVE_MMCV, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, build_attention,
                        build_feedforward_network, build_transformer_layer)
from .attention import MultiScaleDeformableAttention


def inverse_sigmoid(x, eps=1e-5):
    """decoder.py:33-48."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


@ATTENTION.register_module()
class CustomMSDeformableAttention(MultiScaleDeformableAttention):
    """decoder.py:131-338 -- the same arithmetic as mmcv's MultiScaleDeformableAttention (the reference file is a copy
    of it with an extra ``flag`` keyword), so it shares the implementation backed by ``ub_msda_fwd``."""

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, flag='decoder', **kwargs):
        return super().forward(query, key, value, identity, query_pos, key_padding_mask, reference_points,
                               spatial_shapes, level_start_index, **kwargs)


class MultiheadAttention(nn.Module):
    """mmcv ``MultiheadAttention``: ``nn.MultiheadAttention`` + positional encodings on query / key + dropout + identity."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=None, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__()
        drop_prob = 0.
        if dropout_layer:
            drop_prob = dropout_layer.get('drop_prob', 0.)
        if 'dropout' in kwargs:            # deprecated spelling used by the UniBEV configs (dropout=0.1)
            attn_drop = kwargs.pop('dropout')
            drop_prob = attn_drop
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Dropout(drop_prob) if drop_prob > 0 else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


if not HAVE_MMCV:
    ATTENTION.register_module(module=MultiheadAttention)


class DetrTransformerDecoderLayer(nn.Module):
    """mmcv/mmdet ``DetrTransformerDecoderLayer`` = ``BaseTransformerLayer`` with six operations."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), ffn_num_fcs=2, batch_first=False,
                 ffn_cfgs=None, init_cfg=None, **kwargs):
        super().__init__()
        if len(operation_order) != 6 or set(operation_order) != {'self_attn', 'norm', 'cross_attn', 'ffn'}:
            raise ValueError(f'DetrTransformerDecoderLayer needs the six operations self_attn / norm / cross_attn / ffn, '
                             f'got {operation_order}')
        if norm_cfg.get('type', 'LN') != 'LN':
            raise NotImplementedError('only LayerNorm norms are used by the UniBEV configs')
        ffn_cfgs = copy.deepcopy(ffn_cfgs) if ffn_cfgs else dict(type='FFN', embed_dims=256, feedforward_channels=1024,
                                                                 num_fcs=2, ffn_drop=0., act_cfg=act_cfg)
        ffn_cfgs.update(feedforward_channels=feedforward_channels, ffn_drop=ffn_dropout, num_fcs=ffn_num_fcs)
        self.batch_first, self.operation_order = batch_first, tuple(operation_order)
        self.pre_norm = operation_order[0] == 'norm'
        self.num_attn = 2
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(2)]
        if len(attn_cfgs) != 2:
            raise ValueError('two attn_cfgs (self, cross) are required')
        self.attentions = nn.ModuleList()
        for i, op in enumerate(o for o in operation_order if o in ('self_attn', 'cross_attn')):
            cfg = copy.deepcopy(dict(attn_cfgs[i]))
            cfg.setdefault('batch_first', batch_first)
            att = build_attention(cfg)
            att.operation_name = op
            self.attentions.append(att)
        self.embed_dims = self.attentions[0].embed_dims
        ffn_cfgs['embed_dims'] = self.embed_dims
        self.ffns = nn.ModuleList([build_feedforward_network(ffn_cfgs, dict(type='FFN'))])
        self.norms = nn.ModuleList(nn.LayerNorm(self.embed_dims) for _ in range(operation_order.count('norm')))

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_i = attn_i = ffn_i = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None] * self.num_attn
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
        for op in self.operation_order:
            if op == 'self_attn':
                query = self.attentions[attn_i](query, query, query, identity if self.pre_norm else None,
                                                query_pos=query_pos, key_pos=query_pos, attn_mask=attn_masks[attn_i],
                                                key_padding_mask=query_key_padding_mask, **kwargs)
                attn_i += 1
                identity = query
            elif op == 'norm':
                query = self.norms[norm_i](query)
                norm_i += 1
            elif op == 'cross_attn':
                query = self.attentions[attn_i](query, key, value, identity if self.pre_norm else None,
                                                query_pos=query_pos, key_pos=key_pos, attn_mask=attn_masks[attn_i],
                                                key_padding_mask=key_padding_mask, **kwargs)
                attn_i += 1
                identity = query
            else:
                query = self.ffns[ffn_i](query, identity if self.pre_norm else None)
                ffn_i += 1
        return query


if not HAVE_MMCV:
    TRANSFORMER_LAYER.register_module(module=DetrTransformerDecoderLayer)


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetectionTransformerDecoder(nn.Module):
    """decoder.py:51-128."""

    def __init__(self, transformerlayers=None, num_layers=None, return_intermediate=False, init_cfg=None, **kwargs):
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
        self.fp16_enabled = False

    def forward(self, query, *args, reference_points=None, reg_branches=None, key_padding_mask=None, **kwargs):
        """query (num_query, bs, C); reference_points (bs, num_query, 3) in [0, 1].  -> (stacked layer outputs
        (L, num_query, bs, C), stacked reference points (L, bs, num_query, 3)) with ``return_intermediate``, else the
        last output and reference points."""
        output = query
        intermediate, intermediate_reference_points = [], []
        for lid, layer in enumerate(self.layers):
            reference_points_input = reference_points[..., :2].unsqueeze(2)          # (bs, num_query, 1 level, 2)
            output = layer(output, *args, reference_points=reference_points_input, key_padding_mask=key_padding_mask,
                           **kwargs)
            output = output.permute(1, 0, 2)
            if reg_branches is not None:
                tmp = reg_branches[lid](output)
                assert reference_points.shape[-1] == 3
                new_reference_points = torch.zeros_like(reference_points)
                new_reference_points[..., :2] = tmp[..., :2] + inverse_sigmoid(reference_points[..., :2])
                new_reference_points[..., 2:3] = tmp[..., 4:5] + inverse_sigmoid(reference_points[..., 2:3])
                reference_points = new_reference_points.sigmoid().detach()
            output = output.permute(1, 0, 2)
            if self.return_intermediate:
                intermediate.append(output)
                intermediate_reference_points.append(reference_points)
        if self.return_intermediate:
            return torch.stack(intermediate), torch.stack(intermediate_reference_points)
        return output, reference_points
