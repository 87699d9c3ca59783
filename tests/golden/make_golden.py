#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own
hot-path modules, unmodified, from /root/reference.

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

The reference imports mmcv / mmdet / fontTools / matplotlib, none of which exist
in this image (and mmcv-full 1.3.17 cannot be built against torch 2.11), so this
script first installs small import stubs for exactly the symbols those files
use.  The stubs restate published mmcv 1.3.17 behaviour (registries,
``BaseTransformerLayer`` construction order, ``FFN``, ``MultiScaleDeformableAttention``,
``multi_scale_deformable_attn_pytorch``); everything else that runs -- reference
points, camera projection, masks, rebatch / scatter / count, offset
normalisation, Z-anchor interleave, CNW, fusion, flags -- is the reference's own
code.  The MSDA-core known-answer vectors are produced by the independent
implementation inside ``transformers`` (modeling_deformable_detr.py), not by the
stub.  This script only runs in the build container (it reads /root/reference);
the committed .npz files are what travels to the GPU box.
"""
import copy
import importlib
import math
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF_MODULES = '/root/reference/projects/UniBEV/unibev_plugin/models/modules'


# --------------------------------------------------------------------------- #
# import stubs                                                                #
# --------------------------------------------------------------------------- #
class Registry:
    def __init__(self, name):
        self.name, self.table = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.table[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.table.get(key)

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop('type')
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f'{typ} is not in the {registry.name} registry')
    return cls(**args)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()
        self._is_init = True


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if hasattr(module, 'weight') and module.weight is not None:
        (nn.init.xavier_uniform_ if distribution == 'uniform' else nn.init.xavier_normal_)(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _passthrough_decorator(*dargs, **dkw):
    def deco(fn):
        return fn
    return deco


def digit_version(s):
    out = []
    for tok in str(s).split('+')[0].split('.'):
        num = ''.join(ch for ch in tok if ch.isdigit())
        out.append(int(num) if num else 0)
    return tuple(out)


ATTENTION = Registry('attention')
FEEDFORWARD_NETWORK = Registry('feed-forward network')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
TRANSFORMER = Registry('Transformer')


def msda_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """mmcv multi_scale_deformable_attn_pytorch (published algorithm)."""
    bs, _, nh, dh = value.shape
    _, nq, _, nl, npnt, _ = sampling_locations.shape
    vl = value.split([int(h) * int(w) for h, w in value_spatial_shapes], dim=1)
    grids = 2 * sampling_locations - 1
    svl = []
    for lvl, (h, w) in enumerate(value_spatial_shapes):
        v = vl[lvl].flatten(2).transpose(1, 2).reshape(bs * nh, dh, int(h), int(w))
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        svl.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros', align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(bs * nh, 1, nq, nl * npnt)
    out = (torch.stack(svl, dim=-2).flatten(-2) * aw).sum(-1).view(bs, nh * dh, nq)
    return out.transpose(1, 2).contiguous()


class MSDAFunction:
    @staticmethod
    def apply(value, shapes, lsi, loc, w, step):
        return msda_pytorch(value, shapes, loc, w)


class MultiScaleDeformableAttention(BaseModule):
    """mmcv 1.3.17 module (same forward as the reference's verbatim copy,
    decoder.py:230-338)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                 dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__(init_cfg)
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.embed_dims, self.num_levels, self.num_heads, self.num_points = embed_dims, num_levels, num_heads, num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        constant_init(self.sampling_offsets, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(
            1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid.view(-1)
        constant_init(self.attention_weights, val=0., bias=0.)
        xavier_init(self.value_proj, distribution='uniform', bias=0.)
        xavier_init(self.output_proj, distribution='uniform', bias=0.)
        self._is_init = True

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
        bs, nq, _ = query.shape
        _, nv, _ = value.shape
        assert (spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum() == nv
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, nv, self.num_heads, -1)
        off = self.sampling_offsets(query).view(bs, nq, self.num_heads, self.num_levels, self.num_points, 2)
        aw = self.attention_weights(query).view(bs, nq, self.num_heads, self.num_levels * self.num_points)
        aw = aw.softmax(-1).view(bs, nq, self.num_heads, self.num_levels, self.num_points)
        assert reference_points.shape[-1] == 2
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        out = self.output_proj(msda_pytorch(value, spatial_shapes, loc, aw))
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return self.dropout(out) + identity


ATTENTION.register_module()(MultiScaleDeformableAttention)


@FEEDFORWARD_NETWORK.register_module()
class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2 and act_cfg['type'] == 'ReLU'
        layers, cin = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(cin, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)))
            cin = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = nn.Identity()
        self.add_identity = add_identity
        self.embed_dims = embed_dims

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


def build_norm_layer(cfg, num_features):
    assert cfg['type'] == 'LN'
    return 'ln', nn.LayerNorm(num_features)


class BaseTransformerLayer(BaseModule):
    """mmcv 1.3.17 construction semantics (attentions -> ffns -> norms)."""

    def __init__(self, attn_cfgs=None,
                 ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0.,
                               act_cfg=dict(type='ReLU', inplace=True)),
                 operation_order=None, norm_cfg=dict(type='LN'), init_cfg=None, batch_first=False, **kwargs):
        for ori, new in dict(feedforward_channels='feedforward_channels', ffn_dropout='ffn_drop',
                             ffn_num_fcs='num_fcs').items():
            if ori in kwargs:
                ffn_cfgs[new] = kwargs[ori]
        super().__init__(init_cfg)
        self.batch_first = batch_first
        num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        assert num_attn == len(attn_cfgs)
        self.num_attn, self.operation_order, self.norm_cfg = num_attn, operation_order, norm_cfg
        self.pre_norm = operation_order[0] == 'norm'
        self.attentions = ModuleList()
        idx = 0
        for op in operation_order:
            if op in ('self_attn', 'cross_attn'):
                if 'batch_first' in attn_cfgs[idx]:
                    assert self.batch_first == attn_cfgs[idx]['batch_first']
                else:
                    attn_cfgs[idx]['batch_first'] = self.batch_first
                att = ATTENTION.build(attn_cfgs[idx])
                att.operation_name = op
                self.attentions.append(att)
                idx += 1
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = ModuleList()
        n_ffn = operation_order.count('ffn')
        if isinstance(ffn_cfgs, dict):
            ffn_cfgs = [copy.deepcopy(ffn_cfgs) for _ in range(n_ffn)]
        for i in range(n_ffn):
            if 'embed_dims' not in ffn_cfgs[i]:
                ffn_cfgs[i]['embed_dims'] = self.embed_dims
            else:
                assert ffn_cfgs[i]['embed_dims'] == self.embed_dims
            self.ffns.append(build_from_cfg(ffn_cfgs[i], FEEDFORWARD_NETWORK, dict(type='FFN')))
        self.norms = ModuleList()
        for _ in range(operation_order.count('norm')):
            self.norms.append(build_norm_layer(norm_cfg, self.embed_dims)[1])


def _base_layer_forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                        query_key_padding_mask=None, key_padding_mask=None, **kwargs):
    """mmcv 1.3.17 BaseTransformerLayer.forward (operation dispatch)."""
    norm_index = attn_index = ffn_index = 0
    identity = query
    if attn_masks is None:
        attn_masks = [None for _ in range(self.num_attn)]
    for layer in self.operation_order:
        if layer == 'self_attn':
            temp_key = temp_value = query
            query = self.attentions[attn_index](query, temp_key, temp_value, identity if self.pre_norm else None,
                                                query_pos=query_pos, key_pos=query_pos, attn_mask=attn_masks[attn_index],
                                                key_padding_mask=query_key_padding_mask, **kwargs)
            attn_index += 1
            identity = query
        elif layer == 'norm':
            query = self.norms[norm_index](query)
            norm_index += 1
        elif layer == 'cross_attn':
            query = self.attentions[attn_index](query, key, value, identity if self.pre_norm else None,
                                                query_pos=query_pos, key_pos=key_pos, attn_mask=attn_masks[attn_index],
                                                key_padding_mask=key_padding_mask, **kwargs)
            attn_index += 1
            identity = query
        elif layer == 'ffn':
            query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
            ffn_index += 1
    return query


BaseTransformerLayer.forward = _base_layer_forward


@ATTENTION.register_module()
class MultiheadAttention(BaseModule):
    """mmcv 1.3.17 MultiheadAttention wrapper (published behaviour)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=dict(type='Dropout', drop_prob=0.),
                 init_cfg=None, batch_first=False, **kwargs):
        super().__init__(init_cfg)
        dropout_layer = dict(dropout_layer)
        if 'dropout' in kwargs:
            attn_drop = kwargs['dropout']
            dropout_layer['drop_prob'] = kwargs.pop('dropout')
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Dropout(dropout_layer['drop_prob']) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None:
            if query_pos is not None:
                if query_pos.shape == key.shape:
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


@TRANSFORMER_LAYER.register_module()
class DetrTransformerDecoderLayer(BaseTransformerLayer):
    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), ffn_num_fcs=2, **kwargs):
        super().__init__(attn_cfgs=attn_cfgs, feedforward_channels=feedforward_channels, ffn_dropout=ffn_dropout,
                         operation_order=operation_order, act_cfg=act_cfg, norm_cfg=norm_cfg, ffn_num_fcs=ffn_num_fcs,
                         **kwargs)
        assert len(operation_order) == 6
        assert set(operation_order) == set(['self_attn', 'norm', 'cross_attn', 'ffn'])


class TransformerLayerSequence(BaseModule):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(TRANSFORMER_LAYER.build(transformerlayers[i]))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    ext = types.SimpleNamespace(load_ext=lambda *a, **k: types.SimpleNamespace())
    path = mod('mmcv.utils.path', mkdir_or_exist=lambda p: os.makedirs(p, exist_ok=True))
    utils = mod('mmcv.utils', ext_loader=ext, TORCH_VERSION=torch.__version__, digit_version=digit_version,
                ConfigDict=ConfigDict, build_from_cfg=build_from_cfg,
                deprecated_api_warning=_passthrough_decorator, to_2tuple=lambda x: (x, x), path=path)
    ops_msda = mod('mmcv.ops.multi_scale_deform_attn', multi_scale_deformable_attn_pytorch=msda_pytorch,
                   MultiScaleDeformableAttnFunction=MSDAFunction,
                   MultiScaleDeformableAttention=MultiScaleDeformableAttention)
    ops = mod('mmcv.ops', multi_scale_deform_attn=ops_msda)
    registry = mod('mmcv.cnn.bricks.registry', ATTENTION=ATTENTION, TRANSFORMER_LAYER=TRANSFORMER_LAYER,
                   TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE, FEEDFORWARD_NETWORK=FEEDFORWARD_NETWORK)
    transformer = mod('mmcv.cnn.bricks.transformer', build_attention=ATTENTION.build,
                      build_transformer_layer_sequence=TRANSFORMER_LAYER_SEQUENCE.build,
                      BaseTransformerLayer=BaseTransformerLayer, TransformerLayerSequence=TransformerLayerSequence,
                      FFN=FFN)
    bricks = mod('mmcv.cnn.bricks', registry=registry, transformer=transformer)
    cnn = mod('mmcv.cnn', xavier_init=xavier_init, constant_init=constant_init, bricks=bricks)
    base_module = mod('mmcv.runner.base_module', BaseModule=BaseModule, ModuleList=ModuleList, Sequential=Sequential)
    runner = mod('mmcv.runner', force_fp32=_passthrough_decorator, auto_fp16=_passthrough_decorator,
                 base_module=base_module)
    mod('mmcv', utils=utils, ops=ops, cnn=cnn, runner=runner)
    b = mod('mmdet.models.utils.builder', TRANSFORMER=TRANSFORMER)
    u = mod('mmdet.models.utils', builder=b)
    mm = mod('mmdet.models', utils=u)
    mod('mmdet', models=mm)
    mod('fontTools.ttLib')
    mod('fontTools', ttLib=sys.modules['fontTools.ttLib'])
    mod('matplotlib.pyplot')
    mod('matplotlib', pyplot=sys.modules['matplotlib.pyplot'])
    mod('cv2')


def import_reference():
    install_stubs()
    pkg = types.ModuleType('refmods')
    pkg.__path__ = [REF_MODULES]
    sys.modules['refmods'] = pkg
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        fusion = importlib.import_module('refmods.transformer_fusion')
        enc_img = importlib.import_module('refmods.encoder_unibev_detr_img')
        sca_img = importlib.import_module('refmods.spatial_cross_attention_img')
        sca_pts = importlib.import_module('refmods.spatial_cross_attention_pts')
        importlib.import_module('refmods.encoder_unibev_detr_pts')

    @TRANSFORMER_LAYER_SEQUENCE.register_module()
    class NullDecoder(nn.Module):
        """Stands in for DetectionTransformerDecoder (out of scope): the golden
        vectors stop at fused_bev_embed."""
        def __init__(self, **kw):
            super().__init__()

        def forward(self, **kw):
            return None, None
    return fusion, enc_img, sca_img, sca_pts


# --------------------------------------------------------------------------- #
# synthetic inputs                                                            #
# --------------------------------------------------------------------------- #
def camera_rig(num_cams, img_hw, seed=0, jitter=0.0):
    """Seeded pinhole rig: cameras around the ego origin looking outwards.
    Returns list of num_cams 4x4 float64 lidar2img matrices."""
    rng = np.random.RandomState(seed)
    H, W = img_hw
    mats = []
    for i in range(num_cams):
        yaw = 2 * math.pi * i / num_cams + jitter * rng.randn()
        f = 0.8 * W * (1.0 + 0.05 * rng.randn() * (jitter > 0))
        K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]])
        fwd = np.array([math.cos(yaw), math.sin(yaw), 0.0])
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        R = np.stack([right, -up, fwd], 0)            # camera x right, y down, z forward
        t = np.array([0.3 * math.cos(yaw), 0.3 * math.sin(yaw), -0.3]) + jitter * rng.randn(3)
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = -R @ t
        P = np.eye(4)
        P[:3, :3] = K
        mats.append(P @ E)
    return mats


def transformer_cfg(C=32, heads=4, layers=2, cams=3, fusion='linear', feature_norm='ChannelNormWeights',
                    with_img=True, with_pts=True, d_img=4, d_pts=4, points=8, drop_modality=None,
                    pc_range=(-54, -54, -5, 54, 54, 3), spatial_norm=None, dual_queries=False, bev_h=200, bev_w=200,
                    use_modal_embeds=None):
    def layer(kind):
        return dict(
            type=f'{kind}Layer',
            attn_cfgs=[
                dict(type='MultiScaleDeformableAttention', embed_dims=C, num_levels=1, num_heads=heads, dropout=0.0),
                dict(type=f'SpatialCrossAttention{kind}', pc_range=list(pc_range), dropout=0.0, num_cams=cams,
                     deformable_attention=dict(type=f'MSDeformableAttention3D{kind}', embed_dims=C,
                                               num_heads=heads, num_points=points, num_levels=1),
                     embed_dims=C)],
            ffn_cfgs=dict(type='FFN', embed_dims=C),
            feedforward_channels=2 * C, ffn_dropout=0.0,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))
    cfg = dict(type='UniBEVTransformer', embed_dims=C, num_cams=cams, fusion_method=fusion,
               drop_modality=drop_modality, feature_norm=feature_norm, spatial_norm=spatial_norm,
               dual_queries=dual_queries, bev_h=bev_h, bev_w=bev_w, decoder=dict(type='NullDecoder'),
               use_modal_embeds=use_modal_embeds)
    if with_img:
        cfg['img_encoder'] = dict(type='ImgEncoder', num_layers=layers, pc_range=list(pc_range),
                                  num_points_in_pillar=d_img, return_intermediate=False,
                                  transformerlayers=layer('Img'))
    if with_pts:
        cfg['pts_encoder'] = dict(type='PtsEncoder', num_layers=layers, pc_range=list(pc_range),
                                  num_points_in_pillar_lidar=d_pts, return_intermediate=False,
                                  transformerlayers=layer('Pts'))
    return cfg


def randomize_(module, seed):
    """Replace every parameter by seeded noise scaled so sampling is
    non-degenerate (default init zeroes the offset / attention linears)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith('sampling_offsets.weight'):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.05)
            elif name.endswith('sampling_offsets.bias'):
                prm.add_(torch.randn(prm.shape, generator=g) * 0.5)
            elif name.endswith('attention_weights.weight'):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.3)
            elif 'norms' in name and name.endswith('weight'):
                prm.copy_(1.0 + 0.1 * torch.randn(prm.shape, generator=g))
            elif prm.dim() > 1:
                prm.copy_(torch.randn(prm.shape, generator=g) / math.sqrt(prm.shape[-1]))
            else:
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.2)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB')


# --------------------------------------------------------------------------- #
# vectors                                                                     #
# --------------------------------------------------------------------------- #
def golden_msda_core():
    """Known-answer vectors for the MSDA core, produced by transformers'
    independent implementation; includes adversarial coordinates."""
    from transformers.models.deformable_detr.modeling_deformable_detr import \
        MultiScaleDeformableAttention as HFCore
    g = torch.Generator().manual_seed(1234)
    B, H, D, Nq, P = 2, 4, 8, 37, 4
    shapes = [(5, 7), (3, 4)]
    Nv = sum(h * w for h, w in shapes)
    value = torch.randn(B, Nv, H, D, generator=g)
    loc = torch.rand(B, Nq, H, len(shapes), P, 2, generator=g) * 1.4 - 0.2
    # adversarial rows: exact pixel centres / borders / far outside / huge
    loc[0, 0] = 0.0
    loc[0, 1] = 1.0
    loc[0, 2] = 0.5
    loc[0, 3, :, 0, :, 0] = (torch.arange(P).float() + 0.5) / 7
    loc[0, 3, :, 0, :, 1] = (torch.arange(P).float() + 0.5) / 5
    loc[0, 4] = -1e-3
    loc[0, 5] = 1 + 1e-3
    loc[0, 6] = 3.0e6
    loc[0, 7] = -3.0e6
    loc[0, 8, :, :, :, 0] = 1.0 / 14          # x_pix = 0 exactly on level 0
    loc[0, 9, :, :, :, 1] = -0.1 / 5 + 1e-7   # h_im just above -0.6
    w = torch.rand(B, Nq, H, len(shapes), P, generator=g)
    w = w / w.sum((-1, -2), keepdim=True)
    out = HFCore()(value, torch.tensor(shapes), shapes, None, loc, w, 64)
    save('msda_core', value=value, shapes=np.array(shapes), loc=loc, w=w, out=out)


def golden_point_sampling(enc_img):
    H, W, D = 9, 11, 4
    pc = [-54, -54, -5, 54, 54, 3]
    img_hw = (96, 160)
    metas = [dict(lidar2img=camera_rig(3, img_hw, seed=s, jitter=0.05 * s), img_shape=[(img_hw[0], img_hw[1], 3)] * 3)
             for s in range(2)]
    enc = enc_img.ImgEncoder.__new__(enc_img.ImgEncoder)
    ref_3d = enc_img.ImgEncoder.get_reference_points(H, W, pc[5] - pc[2], D, dim='3d', bs=2, device='cpu')
    ref_2d = enc_img.ImgEncoder.get_reference_points(H, W, dim='2d', bs=2, device='cpu')
    ref_cam, mask = enc_img.ImgEncoder.point_sampling(enc, ref_3d, pc, metas)
    save('point_sampling', lidar2img=np.asarray([m['lidar2img'] for m in metas]), img_hw=np.array(img_hw),
         pc_range=np.array(pc, dtype=np.float64), bev_hw=np.array([H, W]), D=D,
         ref_3d=ref_3d, ref_2d=ref_2d, ref_cam=ref_cam, mask=mask)


def golden_attention_modules(sca_img, sca_pts):
    torch.manual_seed(7)
    C, heads, P, D = 32, 4, 8, 4
    # inner attention, image flavour, 2 levels
    att = sca_img.MSDeformableAttention3DImg(embed_dims=C, num_heads=heads, num_levels=2, num_points=P)
    randomize_(att, 11)
    att.eval()
    shapes = torch.tensor([[6, 8], [3, 4]])
    Nv = int((shapes[:, 0] * shapes[:, 1]).sum())
    q = torch.randn(3, 19, C)
    v = torch.randn(3, Nv, C)
    ref = torch.rand(3, 19, D, 2) * 1.2 - 0.1
    out = att(q, value=v, reference_points=ref, spatial_shapes=shapes, level_start_index=torch.tensor([0, 48]))
    sd = {'p.' + k: t for k, t in att.state_dict().items()}
    save('msda3d_img', query=q, value=v, ref=ref, shapes=shapes, out=out, heads=heads, points=P, **sd)

    # full SpatialCrossAttentionImg with per-item calibration (batch-0 index quirk visible)
    Hb, Wb, cams = 8, 10, 3
    pc = [-54, -54, -5, 54, 54, 3]
    img_hw = (96, 160)
    metas = [dict(lidar2img=camera_rig(cams, img_hw, seed=s, jitter=0.2 * s), img_shape=[(img_hw[0], img_hw[1], 3)] * cams)
             for s in range(2)]
    enc_img_mod = sys.modules['refmods.encoder_unibev_detr_img']
    enc = enc_img_mod.ImgEncoder.__new__(enc_img_mod.ImgEncoder)
    ref_3d = enc_img_mod.ImgEncoder.get_reference_points(Hb, Wb, 8, D, dim='3d', bs=2, device='cpu')
    ref_cam, mask = enc_img_mod.ImgEncoder.point_sampling(enc, ref_3d, pc, metas)
    sca = sca_img.SpatialCrossAttentionImg(embed_dims=C, num_cams=cams, pc_range=pc, dropout=0.0, batch_first=True,
                                           deformable_attention=dict(type='MSDeformableAttention3DImg', embed_dims=C,
                                                                     num_heads=heads, num_points=P, num_levels=1))
    randomize_(sca, 13)
    sca.eval()
    fh, fw = 6, 10
    feats = torch.randn(cams, fh * fw, 2, C)
    query = torch.randn(2, Hb * Wb, C)
    out = sca(query, feats, feats, None, query_pos=None, reference_points_cam=ref_cam, bev_mask=mask,
              spatial_shapes=torch.tensor([[fh, fw]]), level_start_index=torch.tensor([0]))
    sd = {'p.' + k: t for k, t in sca.state_dict().items()}
    save('sca_img', query=query, feats=feats, ref_cam=ref_cam, mask=mask, shapes=np.array([[fh, fw]]), out=out,
         heads=heads, points=P, cams=cams, **sd)

    # SpatialCrossAttentionPts
    scp = sca_pts.SpatialCrossAttentionPts(embed_dims=C, pc_range=pc, dropout=0.0, batch_first=True,
                                           deformable_attention=dict(type='MSDeformableAttention3DPts', embed_dims=C,
                                                                     num_heads=heads, num_points=P, num_levels=1))
    randomize_(scp, 17)
    scp.eval()
    ph, pw = 7, 9
    pfeat = torch.randn(ph * pw, 2, C)
    ref_l = ref_3d.permute(1, 0, 2, 3)[..., :2]
    out = scp(query, pfeat, pfeat, None, query_pos=None, reference_points_lidar=ref_l,
              spatial_shapes=torch.tensor([[ph, pw]]), level_start_index=torch.tensor([0]))
    sd = {'p.' + k: t for k, t in scp.state_dict().items()}
    save('sca_pts', query=query, feats=pfeat, ref_lidar=ref_l, shapes=np.array([[ph, pw]]), out=out,
         heads=heads, points=P, **sd)


def golden_encoder_half(fusion_mod, tag, seed, bs=2, bev_hw=(10, 12), img_fhw=(6, 10), pts_fhw=(9, 9),
                        train_flags=False, **cfg_kw):
    cfg = transformer_cfg(bev_h=bev_hw[0], bev_w=bev_hw[1], **cfg_kw)
    C, cams = cfg['embed_dims'], cfg['num_cams']
    build_cfg = copy.deepcopy(cfg)
    build_cfg.pop('type')
    torch.manual_seed(seed)
    model = fusion_mod.UniBEVTransformer(**build_cfg)
    model.init_weights()
    randomize_(model, seed + 1)
    model.eval()
    g = torch.Generator().manual_seed(seed + 2)
    Hb, Wb = bev_hw
    scale = 2 if cfg['fusion_method'] == 'cat' else 1
    with_img, with_pts = 'img_encoder' in cfg, 'pts_encoder' in cfg
    img_hw = (96, 160)
    img_feats = [torch.randn(bs, cams, C, *img_fhw, generator=g)] if with_img else None
    pts_feats = [torch.randn(bs, C, *pts_fhw, generator=g)] if with_pts else None
    if cfg['dual_queries']:
        bev_q = [torch.randn(Hb * Wb, C, generator=g), torch.randn(Hb * Wb, C, generator=g)]
    else:
        bev_q = torch.randn(Hb * Wb, C, generator=g)
    bev_pos = torch.randn(bs, C, Hb, Wb, generator=g)
    obj_q = torch.randn(5, 2 * C * scale, generator=g)
    metas = [dict(lidar2img=camera_rig(cams, img_hw, seed=seed + s, jitter=0.1 * s),
                  img_shape=[(img_hw[0], img_hw[1], 3)] * cams) for s in range(bs)]
    captured = {}
    if with_img:
        model.img_bev_encoder.register_forward_hook(lambda m, i, o: captured.__setitem__('img_bev_embed', o))
    if with_pts:
        model.pts_bev_encoder.register_forward_hook(lambda m, i, o: captured.__setitem__('pts_bev_embed', o))
    if train_flags:
        model.train()                       # every dropout p is 0.0 in this cfg; only the flags are random
        np.random.seed(seed)
    # the ModalityProjection / MLP-modal-embedding branches move their flag tensors with `.cuda()`
    # (transformer_fusion.py:297-298,305); there is no GPU in the build container, so that call is made a no-op while
    # the reference runs (everything else stays on the CPU anyway)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            fused = model(img_feats, pts_feats, bev_q, obj_q, Hb, Wb, bev_pos=bev_pos, img_metas=metas)[0]
    finally:
        torch.Tensor.cuda = real_cuda
    fused = fused.permute(1, 0, 2)          # fusion.py:549 hands (Nq, B, C) to the decoder; store batch-first
    arrays = dict(fused=fused, bev_pos=bev_pos, bev_hw=np.array(bev_hw), img_hw=np.array(img_hw),
                  lidar2img=np.asarray([m['lidar2img'] for m in metas]), flags=np.array([model.c_flag, model.l_flag]),
                  cfg_json=np.array(__import__('json').dumps(cfg)))
    if cfg['dual_queries']:
        arrays.update(bev_queries_img=bev_q[0], bev_queries_pts=bev_q[1])
    else:
        arrays['bev_queries'] = bev_q
    if with_img:
        arrays['img_feats'] = img_feats[0]
    if with_pts:
        arrays['pts_feats'] = pts_feats[0]
    arrays.update(captured)
    arrays.update({'p.' + k: t for k, t in model.state_dict().items() if not k.startswith('decoder')})
    save('encoder_half_' + tag, **arrays)


def golden_decoder(seed=700, C=32, heads=4, layers=3, nq=37, bs=2, bev_hw=(10, 12)):
    """The reference's own DetectionTransformerDecoder + CustomMSDeformableAttention (decoder.py, unmodified) over the mmcv
    layer / MultiheadAttention stubs: 3 layers, iterative reference-point refinement through seeded reg branches."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        dec_mod = importlib.import_module('refmods.decoder')
    cfg = dict(type='DetectionTransformerDecoder', num_layers=layers, return_intermediate=True,
               transformerlayers=dict(
                   type='DetrTransformerDecoderLayer',
                   attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=heads, dropout=0.0),
                              dict(type='CustomMSDeformableAttention', embed_dims=C, num_heads=heads, num_levels=1,
                                   dropout=0.0)],
                   ffn_cfgs=dict(type='FFN', embed_dims=C), feedforward_channels=2 * C, ffn_dropout=0.0,
                   operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))
    dec = TRANSFORMER_LAYER_SEQUENCE.build(copy.deepcopy(cfg))
    assert type(dec) is dec_mod.DetectionTransformerDecoder
    assert type(dec.layers[0].attentions[1]) is dec_mod.CustomMSDeformableAttention
    randomize_(dec, seed)
    reg = nn.ModuleList([nn.Sequential(nn.Linear(C, C), nn.ReLU(), nn.Linear(C, 10)) for _ in range(layers)])
    randomize_(reg, seed + 1)
    dec.eval()
    g = torch.Generator().manual_seed(seed + 2)
    Hb, Wb = bev_hw
    query, query_pos = torch.randn(nq, bs, C, generator=g), torch.randn(nq, bs, C, generator=g)
    value = torch.randn(Hb * Wb, bs, C, generator=g)
    ref = torch.rand(bs, nq, 3, generator=g)
    ref[0, 0, :2] = torch.tensor([0.0, 1.0])          # border reference points
    ref[0, 1, :2] = torch.tensor([0.999, 0.001])
    args = dict(key=None, value=value, query_pos=query_pos, spatial_shapes=torch.tensor([[Hb, Wb]]),
                level_start_index=torch.tensor([0]))
    with torch.no_grad():
        inter, inter_ref = dec(query, reference_points=ref, reg_branches=reg, cls_branches=None, **args)
        inter_noreg, ref_noreg = dec(query, reference_points=ref, reg_branches=None, **args)
    arrays = dict(query=query, query_pos=query_pos, value=value, reference_points=ref, bev_hw=np.array(bev_hw),
                  inter_states=inter, inter_references=inter_ref, inter_states_noreg=inter_noreg,
                  inter_references_noreg=ref_noreg, cfg_json=np.array(__import__('json').dumps(cfg)))
    arrays.update({'p.' + k: t for k, t in dec.state_dict().items()})
    arrays.update({'p.reg.' + k: t for k, t in reg.state_dict().items()})
    save('decoder', **arrays)


def golden_head(fusion_mod, tag, seed, fusion='linear', dual_queries=False, bs=2, C=32, heads=4, cams=3, bev_hw=(6, 8),
                num_query=12, dec_layers=2, img_fhw=(5, 8), pts_fhw=(7, 7), max_num=9, score_threshold=None):
    """The reference's own UniBEV_Head.forward / get_bboxes (unibev_head.py, unmodified) and NMSFreeCoder.decode
    (nms_free_coder.py, unmodified) around the reference transformer (encoders + DetectionTransformerDecoder).  Third-party
    pieces are stubbed by their published behaviour: mmdet's DETRHead.__init__ (stores the arguments, builds the transformer
    and the positional encoding, calls _init_layers), LearnedPositionalEncoding, inverse_sigmoid, BaseBBoxCoder."""
    import json
    import types

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class LearnedPositionalEncoding(BaseModule):          # mmdet 2.19
        def __init__(self, num_feats, row_num_embed=50, col_num_embed=50, init_cfg=None):
            super().__init__(init_cfg)
            self.row_embed = nn.Embedding(row_num_embed, num_feats)
            self.col_embed = nn.Embedding(col_num_embed, num_feats)

        def forward(self, mask):
            h, w = mask.shape[-2:]
            x, y = torch.arange(w), torch.arange(h)
            x_embed, y_embed = self.col_embed(x), self.row_embed(y)
            pos = torch.cat((x_embed.unsqueeze(0).repeat(h, 1, 1), y_embed.unsqueeze(1).repeat(1, w, 1)), dim=-1)
            return pos.permute(2, 0, 1).unsqueeze(0).repeat(mask.shape[0], 1, 1, 1)

    def inverse_sigmoid(x, eps=1e-5):                      # mmdet.models.utils.transformer
        x = x.clamp(min=0, max=1)
        return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))

    BBOX_CODERS, HEADS = Registry('bbox_coder'), Registry('head')

    class DETRHead(BaseModule):                            # the part of mmdet's DETRHead.__init__ the subclass relies on
        def __init__(self, num_classes, in_channels, num_query=100, num_reg_fcs=2, transformer=None,
                     sync_cls_avg_factor=False, positional_encoding=None, loss_cls=None, loss_bbox=None, loss_iou=None,
                     train_cfg=None, test_cfg=None, init_cfg=None, **kwargs):
            super().__init__(init_cfg)
            self.num_classes, self.in_channels, self.num_query, self.num_reg_fcs = num_classes, in_channels, num_query, num_reg_fcs
            self.loss_cls = types.SimpleNamespace(use_sigmoid=loss_cls.get('use_sigmoid', False))
            self.cls_out_channels = num_classes if self.loss_cls.use_sigmoid else num_classes + 1
            pe = dict(positional_encoding)
            pe.pop('type')
            self.positional_encoding = LearnedPositionalEncoding(**pe)
            self.transformer = TRANSFORMER.build(transformer)
            self.embed_dims = self.transformer.embed_dims
            self._init_layers()

    core = types.ModuleType('refcore')
    core.__path__ = [os.path.join(REF_MODULES, '..', '..', 'core')]
    sys.modules['refcore'] = core
    for sub in ('bbox', 'bbox.coders'):
        m = types.ModuleType('refcore.' + sub)
        m.__path__ = [os.path.join(REF_MODULES, '..', '..', 'core', *sub.split('.'))]
        sys.modules['refcore.' + sub] = m
    mod('mmdet.core.bbox.builder', BBOX_CODERS=BBOX_CODERS)
    mod('mmdet.core.bbox', BaseBBoxCoder=object, builder=sys.modules['mmdet.core.bbox.builder'])
    mod('mmdet.core', multi_apply=None, reduce_mean=None, bbox=sys.modules['mmdet.core.bbox'])
    sys.modules['mmdet'].core = sys.modules['mmdet.core']
    util = importlib.import_module('refcore.bbox.util')
    coder_mod = importlib.import_module('refcore.bbox.coders.nms_free_coder')
    mod('mmdet.models.utils.transformer', inverse_sigmoid=inverse_sigmoid)
    sys.modules['mmdet.models'].HEADS = HEADS
    mod('mmdet.models.dense_heads', DETRHead=DETRHead)
    mod('mmdet3d.core.bbox.coders', build_bbox_coder=BBOX_CODERS.build)
    mod('mmdet3d.core.bbox', coders=sys.modules['mmdet3d.core.bbox.coders'])
    mod('mmdet3d.core', bbox=sys.modules['mmdet3d.core.bbox'])
    mod('mmdet3d.unibev_plugin.core.bbox.util', normalize_bbox=util.normalize_bbox)
    mod('mmdet3d.unibev_plugin.core.bbox')
    mod('mmdet3d.unibev_plugin.core')
    mod('mmdet3d.unibev_plugin')
    mod('mmdet3d', core=sys.modules['mmdet3d.core'])
    sys.modules['mmcv.cnn'].Linear = nn.Linear
    sys.modules['mmcv.cnn'].bias_init_with_prob = lambda p: float(-np.log((1 - p) / p))
    heads_pkg = types.ModuleType('refheads')
    heads_pkg.__path__ = [os.path.join(REF_MODULES, '..', 'dense_heads')]
    sys.modules['refheads'] = heads_pkg
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        importlib.import_module('refmods.decoder')
        head_mod = importlib.import_module('refheads.unibev_head')

    tcfg = transformer_cfg(C=C, heads=heads, layers=1, cams=cams, fusion=fusion, feature_norm='ChannelNormWeights' if fusion != 'cat' else None,
                           dual_queries=dual_queries, bev_h=bev_hw[0], bev_w=bev_hw[1])
    Cd = C * (2 if fusion == 'cat' else 1)
    tcfg['decoder'] = dict(type='DetectionTransformerDecoder', num_layers=dec_layers, return_intermediate=True,
                           transformerlayers=dict(
                               type='DetrTransformerDecoderLayer',
                               attn_cfgs=[dict(type='MultiheadAttention', embed_dims=Cd, num_heads=heads, dropout=0.0),
                                          dict(type='CustomMSDeformableAttention', embed_dims=Cd, num_heads=heads,
                                               num_levels=1, dropout=0.0)],
                               ffn_cfgs=dict(type='FFN', embed_dims=Cd), feedforward_channels=2 * Cd, ffn_dropout=0.0,
                               operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))
    pc_range = list(tcfg['img_encoder']['pc_range'])
    hcfg = dict(type='UniBEV_Head', bev_h=bev_hw[0], bev_w=bev_hw[1], num_query=num_query, num_classes=10, in_channels=C,
                sync_cls_avg_factor=True, with_box_refine=True, as_two_stage=False, transformer=tcfg,
                bbox_coder=dict(type='NMSFreeCoder', post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
                                pc_range=pc_range, max_num=max_num, num_classes=10, score_threshold=score_threshold),
                positional_encoding=dict(type='LearnedPositionalEncoding', num_feats=C // 2, row_num_embed=bev_hw[0],
                                         col_num_embed=bev_hw[1]),
                loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
                loss_bbox=dict(type='L1Loss', loss_weight=0.25), loss_iou=dict(type='GIoULoss', loss_weight=0.0))
    build = copy.deepcopy(hcfg)
    build.pop('type')
    torch.manual_seed(seed)
    head = head_mod.UniBEV_Head(**build)
    assert type(head.bbox_coder) is coder_mod.NMSFreeCoder
    head.init_weights()
    randomize_(head, seed + 1)
    with torch.no_grad():                                  # spread the boxes over the range so the centre mask bites
        for br in head.reg_branches:
            br[-1].bias.copy_(torch.tensor([0.0, 0.0, 0.3, 0.5, 0.0, 0.2, 0.1, 0.9, 0.0, 0.0]))
    head.eval()
    g = torch.Generator().manual_seed(seed + 2)
    img_hw = (96, 160)
    img_feats = [torch.randn(bs, cams, C, *img_fhw, generator=g)]
    pts_feats = [torch.randn(bs, C, *pts_fhw, generator=g)]
    metas = [dict(lidar2img=camera_rig(cams, img_hw, seed=seed + s, jitter=0.1 * s),
                  img_shape=[(img_hw[0], img_hw[1], 3)] * cams, box_type_3d=lambda b, code_size: b) for s in range(bs)]
    with torch.no_grad():
        outs = head(img_feats, pts_feats, metas)
        boxes = head.get_bboxes({k: (v.clone() if torch.is_tensor(v) else v) for k, v in outs.items()}, metas)
    arrays = dict(img_feats=img_feats[0], pts_feats=pts_feats[0], img_hw=np.array(img_hw),
                  lidar2img=np.asarray([m['lidar2img'] for m in metas]), bev_embed=outs['bev_embed'],
                  all_cls_scores=outs['all_cls_scores'], all_bbox_preds=outs['all_bbox_preds'],
                  cfg_json=np.array(json.dumps(hcfg)))
    for i, (b, sc, lb) in enumerate(boxes):
        arrays.update({f'det{i}_bboxes': b, f'det{i}_scores': sc, f'det{i}_labels': lb})
    arrays.update({'p.' + k: t for k, t in head.state_dict().items()})
    save('head_' + tag, **arrays)


def main():
    torch.set_num_threads(1)
    fusion_mod, enc_img, sca_img, sca_pts = import_reference()
    golden_msda_core()
    golden_point_sampling(enc_img)
    golden_attention_modules(sca_img, sca_pts)
    golden_encoder_half(fusion_mod, 'lc_cnw_linear', 100)
    golden_encoder_half(fusion_mod, 'lc_cat', 200, fusion='cat', feature_norm=None)
    golden_encoder_half(fusion_mod, 'lc_avg_spatial', 300, fusion='avg', spatial_norm='SpatialNormWeights',
                        dual_queries=True)
    golden_encoder_half(fusion_mod, 'c_only', 400, with_pts=False, feature_norm=None, bs=1)
    golden_encoder_half(fusion_mod, 'l_only_cnw', 500, with_img=False, bs=1)
    golden_encoder_half(fusion_mod, 'lc_cnw_dropflags', 600, train_flags=True, drop_modality=1.0)
    golden_encoder_half(fusion_mod, 'lc_cnw_dropdict', 601, train_flags=True,
                        drop_modality=dict(dropout_prob=1.0, lidar_prob=0.0))
    golden_encoder_half(fusion_mod, 'lc_mlp_cnw', 610, feature_norm='MLP_ChannelNormWeights')
    golden_encoder_half(fusion_mod, 'lc_sigmoid_mlp_cnw_dropflags', 611, feature_norm='Sigmoid_MLP_ChannelNormWeights',
                        train_flags=True, drop_modality=1.0)
    golden_encoder_half(fusion_mod, 'lc_modproj_cat', 612, feature_norm='ModalityProjection', fusion='cat')
    golden_encoder_half(fusion_mod, 'lc_modproj_cat_dropflags', 613, feature_norm='ModalityProjection', fusion='cat',
                        train_flags=True, drop_modality=dict(dropout_prob=1.0, lidar_prob=1.0))
    golden_encoder_half(fusion_mod, 'lc_cnw_modal_mlp', 614, use_modal_embeds='MLP')
    golden_decoder()
    golden_head(fusion_mod, 'linear', 800)
    golden_head(fusion_mod, 'cat_dual_thresh', 810, fusion='cat', dual_queries=True, score_threshold=0.6095)


if __name__ == '__main__':
    main()
