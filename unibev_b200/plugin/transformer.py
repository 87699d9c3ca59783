"""``UniBEVTransformer`` -- plugin surface of the uniform BEV encoder + fusion.

Registered name, constructor keywords, parameter names and the ``forward`` signature /
return tuple follow transformer_fusion.py:49-118,416-426,586.  The object-query decoder
(transformer_fusion.py:540-586; ``plugin/decoder.py``) is built from the ``decoder`` config like the
encoders; only when its top-level ``type`` is not registered does ``forward`` return
``(bev_embed, None, init_reference, None)``.

Eval-mode forwards run the fused B200 pipeline (``fused.FusedEncoder``); training-mode
forwards run the autograd-capable module path.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..registry import TRANSFORMER, build_transformer_layer_sequence
from .attention import MSDeformableAttention3DImg, MSDeformableAttention3DPts, MultiScaleDeformableAttention
from .fused import FusedEncoder, fused_supported

_MLP_NORMS = {'MLP_ChannelNormWeights': lambda: nn.ReLU(inplace=True),
              'Leaky_ReLU_MLP_ChannelNormWeights': lambda: nn.LeakyReLU(inplace=True),
              'ELU_MLP_ChannelNormWeights': lambda: nn.ELU(inplace=True),
              'Sigmoid_MLP_ChannelNormWeights': nn.Sigmoid}
_FEATURE_NORMS = (None, 'ChannelNormWeights', 'ModalityProjection') + tuple(_MLP_NORMS)


class ModalityProjectionModule(nn.Module):
    """transformer_fusion.py:26-47: Linear - ReLU - LayerNorm with an identity shortcut."""

    def __init__(self, embed_dims, with_norm=True, with_residual=True):
        super().__init__()
        layers = [nn.Linear(embed_dims, embed_dims), nn.ReLU(inplace=True)]
        if with_norm:
            layers.append(nn.LayerNorm(embed_dims))
        self.net = nn.Sequential(*layers)
        self.with_residual = with_residual

    def forward(self, x):
        out = self.net(x)
        return x + out if self.with_residual else out


@TRANSFORMER.register_module()
class UniBEVTransformer(nn.Module):
    def __init__(self, num_feature_levels=4, num_cams=6, two_stage_num_proposals=300, img_encoder=None,
                 pts_encoder=None, decoder=None, embed_dims=256, use_cams_embeds=True, fusion_method='linear',
                 drop_modality=None, feature_norm=None, spatial_norm=None, use_modal_embeds=None, bev_h=200, bev_w=200,
                 dual_queries=False, vis_output=None, cna_constant_init=None, init_cfg=None, **kwargs):
        super().__init__()
        if fusion_method in ('linear', 'avg'):
            self.scale_factor = 1
        elif fusion_method == 'cat':
            self.scale_factor = 2
        else:
            raise ValueError('Unrecognizable fusion method:{}'.format(fusion_method))
        if feature_norm not in _FEATURE_NORMS:
            raise ValueError(f'unknown feature_norm {feature_norm!r}')
        if feature_norm == 'ModalityProjection' and fusion_method != 'cat':
            raise ValueError("feature_norm='ModalityProjection' needs fusion_method='cat' (transformer_fusion.py:151)")
        if spatial_norm not in (None, 'SpatialNormWeights'):
            raise ValueError(f'unknown spatial_norm {spatial_norm!r}')
        if use_modal_embeds not in (None, 'Fixed', 'MLP'):
            raise ValueError(f'unknown use_modal_embeds {use_modal_embeds!r}')
        if vis_output is not None:
            raise NotImplementedError('vis_output dumping is not supported')
        if img_encoder is not None:
            self.img_bev_encoder = build_transformer_layer_sequence(img_encoder)
        if pts_encoder is not None:
            self.pts_bev_encoder = build_transformer_layer_sequence(pts_encoder)
        self.decoder = None
        if decoder is not None:
            try:
                self.decoder = build_transformer_layer_sequence(decoder)
            except KeyError as e:
                # only a missing TOP-LEVEL decoder type (registry without the decoder classes) means "encoder half only";
                # a typo in a nested `type=` must not silently turn the model into one without a decoder
                if str(decoder.get('type')) not in str(e):
                    raise
                self.decoder = None
        self.dual_queries, self.embed_dims = dual_queries, embed_dims
        self.num_feature_levels, self.num_cams = num_feature_levels, num_cams
        self.fp16_enabled = False
        self.cna_constant_norm = cna_constant_init
        self.bev_h, self.bev_w = bev_h, bev_w
        self.use_cams_embeds = use_cams_embeds
        self.fusion_method, self.drop_modality = fusion_method, drop_modality
        self.feature_norm, self.spatial_norm, self.use_modal_embeds = feature_norm, spatial_norm, use_modal_embeds
        self.two_stage_num_proposals = two_stage_num_proposals
        self.vis_output = vis_output
        self.l_flag = self.c_flag = 1
        self._fused = None
        # arithmetic class of the fused eval pipeline: 'fp32' (the reference's class: 3xTF32 tensor-core projections, fp32
        # sampling) or the opt-in 'fp16' (fp16 operands / value maps, ~3e-3 absolute error); see plugin/fused.py
        self.fused_precision = 'fp32'
        self.init_layers()

    @property
    def with_img_bev_encoder(self):
        return getattr(self, 'img_bev_encoder', None) is not None

    @property
    def with_pts_bev_encoder(self):
        return getattr(self, 'pts_bev_encoder', None) is not None

    def init_layers(self):
        C = self.embed_dims
        if self.feature_norm == 'ChannelNormWeights':
            self.pts_channel_weights = nn.Parameter(torch.zeros(C))
            self.img_channel_weights = nn.Parameter(torch.zeros(C))
        elif self.feature_norm in _MLP_NORMS:     # transformer_fusion.py:136-150
            self.channel_weights_proj = nn.Sequential(nn.Linear(self.bev_h * self.bev_w * 2, 2),
                                                      _MLP_NORMS[self.feature_norm]())
        elif self.feature_norm == 'ModalityProjection':
            self.c_modal_proj = ModalityProjectionModule(C)
            self.l_modal_proj = ModalityProjectionModule(C)
        if self.spatial_norm == 'SpatialNormWeights':
            self.pts_spatial_weights = nn.Parameter(torch.zeros(self.bev_h * self.bev_w))
            self.img_spatial_weights = nn.Parameter(torch.zeros(self.bev_h * self.bev_w))
        if self.with_img_bev_encoder:
            self.img_level_embeds = nn.Parameter(torch.zeros(self.num_feature_levels, C))
            self.cams_embeds = nn.Parameter(torch.zeros(self.num_cams, C))
        if self.with_pts_bev_encoder:
            self.pts_level_embeds = nn.Parameter(torch.zeros(self.num_feature_levels, C))
        if self.use_modal_embeds == 'MLP':        # transformer_fusion.py:171-176
            self.modal_embbeding_mlp = nn.Sequential(nn.Linear(2, C // 2), nn.ReLU(inplace=True), nn.Linear(C // 2, C),
                                                     nn.ReLU(inplace=True))
        elif self.use_modal_embeds == 'Fixed':
            self.modal_embbeding_C = nn.Parameter(torch.zeros(C))
            self.modal_embbeding_L = nn.Parameter(torch.zeros(C))
        self.reference_points = nn.Linear(C * self.scale_factor, 3)

    def init_weights(self):
        """transformer_fusion.py:184-225: xavier on every matrix, then the deformable attentions
        re-run their own init, then the embeddings / CNW weights are drawn."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, (MSDeformableAttention3DPts, MSDeformableAttention3DImg, MultiScaleDeformableAttention)):
                m.init_weights()
        if self.with_pts_bev_encoder:
            nn.init.normal_(self.pts_level_embeds)
        if self.with_img_bev_encoder:
            nn.init.normal_(self.img_level_embeds)
            nn.init.normal_(self.cams_embeds)
        if self.feature_norm == 'ChannelNormWeights':
            if self.cna_constant_norm is True:
                nn.init.constant_(self.pts_channel_weights, 0.5)
                nn.init.constant_(self.img_channel_weights, 0.5)
            else:
                nn.init.normal_(self.pts_channel_weights)
                nn.init.normal_(self.img_channel_weights)
        if self.spatial_norm == 'SpatialNormWeights':
            nn.init.normal_(self.pts_spatial_weights)
            nn.init.normal_(self.img_spatial_weights)
        nn.init.xavier_uniform_(self.reference_points.weight)
        nn.init.constant_(self.reference_points.bias, 0.)
        if self.use_modal_embeds == 'Fixed':
            nn.init.normal_(self.modal_embbeding_C)
            nn.init.normal_(self.modal_embbeding_L)
        self._fused = None

    def get_probability(self, prob):
        return True if np.random.random() < prob else False

    # ------------------------------------------------------------------ flags
    def _draw_flags(self, img_mlvl_feats, pts_mlvl_feats):
        """transformer_fusion.py:463-489."""
        self.l_flag = self.c_flag = 1
        if self.drop_modality is not None and self.training is True:
            if isinstance(self.drop_modality, dict):
                dropout_prob, lidar_prob = self.drop_modality['dropout_prob'], self.drop_modality['lidar_prob']
            elif isinstance(self.drop_modality, float):
                dropout_prob = lidar_prob = self.drop_modality
            else:
                raise ValueError('Unrecognized type: {}'.format(type(self.drop_modality)))
            if self.get_probability(dropout_prob):
                self.l_flag = self.get_probability(lidar_prob) * 1
                self.c_flag = 1 - self.l_flag
        if img_mlvl_feats is None:
            self.c_flag = 0
        elif pts_mlvl_feats is None:
            self.l_flag = 0

    # ------------------------------------------------------------------ module (autograd) path
    def _pre_process_img_feats(self, mlvl_img_feats, bev_queries):
        flat, shapes = [], []
        for lvl, feat in enumerate(mlvl_img_feats):
            bs, n, c, h, w = feat.shape
            if torch.is_grad_enabled():      # differentiable w.r.t. the embeddings
                tok = feat.flatten(3).permute(0, 1, 3, 2).reshape(bs * n, h * w, c)
                if self.use_cams_embeds:
                    tok = tok + self.cams_embeds.repeat(bs, 1)[:, None]
                tok = ops.add_row_vector(tok.contiguous(), self.img_level_embeds[lvl])   # token-major once, not per layer
            else:
                tok = ops.flatten_feats(feat, self.cams_embeds if self.use_cams_embeds else None,
                                        self.img_level_embeds[lvl])
            flat.append(tok.view(bs, n, h * w, c).permute(1, 2, 0, 3))       # (n, hw, bs, c)
            shapes.append((h, w))
        flat = torch.cat(flat, 1)
        starts = [0]
        for h, w in shapes[:-1]:
            starts.append(starts[-1] + h * w)
        return (flat, ops.const_tensor([list(hw) for hw in shapes], torch.long, flat.device),
                ops.const_tensor(starts, torch.long, flat.device))

    def _pre_process_pts_feats(self, mlvl_pts_feats, bev_queries):
        if len(mlvl_pts_feats) != 1:
            # the reference concatenates LiDAR levels on the channel axis (transformer_fusion.py:272): one level only
            raise NotImplementedError('LiDAR features: exactly one level is supported')
        feat = mlvl_pts_feats[0]
        bs, c, h, w = feat.shape
        if torch.is_grad_enabled():
            # (the sum inherits the channel-major strides of `feat`: make it token-major once, not in every layer's value_proj)
            tok = ops.add_row_vector(feat.flatten(2).permute(0, 2, 1).contiguous(), self.pts_level_embeds[0])
        else:
            tok = ops.flatten_feats(feat, None, self.pts_level_embeds[0])
        return (tok.permute(1, 0, 2), ops.const_tensor([[h, w]], torch.long, feat.device),
                ops.const_tensor([0], torch.long, feat.device))

    def _encode_modules(self, img_mlvl_feats, pts_mlvl_feats, bev_queries, bev_h, bev_w, bev_pos, **kwargs):
        bs = (img_mlvl_feats or pts_mlvl_feats)[0].size(0)
        if bev_pos is not None:
            # (Nq, bs, C) as the reference passes it, over token-major memory: the encoders' permute back to (bs, Nq, C) is
            # then contiguous and `query + query_pos` reads rows instead of a transposed map in every layer
            bev_pos = bev_pos.flatten(2).permute(0, 2, 1).contiguous().permute(1, 0, 2)
        def rep(q):     # == q.unsqueeze(1).repeat(1, bs, 1) (transformer_fusion.py:493-498), over batch-major memory: the
            return q.unsqueeze(0).repeat(bs, 1, 1).permute(1, 0, 2)     # encoders' permute back to (bs, Nq, C) is contiguous
        if self.dual_queries:
            q_img, q_pts = rep(bev_queries[0]), rep(bev_queries[1])
        else:
            q_img = q_pts = rep(bev_queries)
        img = pts = None
        if img_mlvl_feats is not None:
            flat, shapes, start = self._pre_process_img_feats(img_mlvl_feats, q_img)
            hw = tuple(img_mlvl_feats[0].shape[-2:]) if len(img_mlvl_feats) == 1 else None   # (one level: fused sampling)
            img = self.img_bev_encoder(q_img, flat, flat, bev_h=bev_h, bev_w=bev_w, bev_pos=bev_pos,
                                       spatial_shapes=shapes, level_start_index=start, ub_value_hw=hw, **kwargs)
        if pts_mlvl_feats is not None:
            flat, shapes, start = self._pre_process_pts_feats(pts_mlvl_feats, q_pts)
            pts = self.pts_bev_encoder(q_pts, flat, flat, bev_h=bev_h, bev_w=bev_w, bev_pos=bev_pos,
                                       spatial_shapes=shapes, level_start_index=start,
                                       ub_value_hw=tuple(pts_mlvl_feats[0].shape[-2:]), **kwargs)
        return img, pts

    def _fuse_modules(self, img, pts):
        """CNW + spatial norm + fusion with torch ops (differentiable), transformer_fusion.py:280-337,386-413."""
        c, l = self.c_flag, self.l_flag
        if img is None:
            img = torch.zeros_like(pts)
        elif pts is None:
            pts = torch.zeros_like(img)
        if self.feature_norm == 'ChannelNormWeights':
            w = torch.stack((self.img_channel_weights, self.pts_channel_weights), 0)
            wi, wp = (w.softmax(0)) if (c == 1 and l == 1) else (w[0:1].softmax(0)[0], w[1:2].softmax(0)[0])
            img, pts = img * wi, pts * wp
        elif self.feature_norm in _MLP_NORMS:       # transformer_fusion.py:345-366: weights from the BEV maps themselves
            w = self.channel_weights_proj(torch.cat([img, pts], dim=1).permute(0, 2, 1))       # (bs, C, 2)
            if c == 1 and l == 1:
                n = torch.softmax(w, dim=-1)
                wi, wp = n[:, :, 0], n[:, :, 1]
            else:
                wi, wp = torch.softmax(w[:, :, :1], dim=-1).squeeze(-1), torch.softmax(w[:, :, 1:], dim=-1).squeeze(-1)
            img, pts = img * wi[:, None, :], pts * wp[:, None, :]
        elif self.feature_norm == 'ModalityProjection':   # :375-381: each modality also predicts the other one
            img, pts = torch.cat([img, self.l_modal_proj(img)], dim=-1), torch.cat([self.c_modal_proj(pts), pts], dim=-1)
        if self.spatial_norm == 'SpatialNormWeights':
            w = torch.stack((self.img_spatial_weights, self.pts_spatial_weights), 0)
            wi, wp = (w.softmax(0)) if (c == 1 and l == 1) else (w[:1].softmax(0)[0], w[1:].softmax(0)[0])
            img, pts = img * wi[None, :, None], pts * wp[None, :, None]
        if self.fusion_method == 'linear':
            fused = c * img + l * pts
        elif self.fusion_method == 'avg':
            fused = img * c / (c + l) + pts * l / (c + l)
        elif self.feature_norm == 'ModalityProjection':   # :287-300: true halves by flag, pseudo halves by 1 - flag
            C = self.embed_dims
            img_flags = torch.cat((img.new_full((C,), float(c)), img.new_full((C,), float(1 - l))))
            pts_flags = torch.cat((img.new_full((C,), float(1 - c)), img.new_full((C,), float(l))))
            fused = img * img_flags + pts * pts_flags
        else:
            fused = torch.cat((img * c, pts * l), -1)
        if self.use_modal_embeds == 'MLP':
            fused = fused + self.modal_embbeding_mlp(fused.new_tensor([float(c), float(l)]))
        elif self.use_modal_embeds == 'Fixed':
            fused = fused + c * self.modal_embbeding_C + l * self.modal_embbeding_L
        return fused

    def _kernel_fusable(self):
        """ub_cnw_fuse covers feature_norm None / ChannelNormWeights, SpatialNormWeights, linear / avg / cat and the fixed
        modal embeddings; the ablation variants (MLP norms, ModalityProjection, MLP modal embeds) are torch glue."""
        return self.feature_norm in (None, 'ChannelNormWeights') and self.use_modal_embeds in (None, 'Fixed')

    def fuse(self, img, pts):
        """CNW / spatial norm / fusion of the two BEV maps (either may be None) in inference."""
        if not self._kernel_fusable():
            with torch.no_grad():
                return self._fuse_modules(img, pts)
        return ops.cnw_fuse(img, pts, getattr(self, 'img_channel_weights', None),
                            getattr(self, 'pts_channel_weights', None), self.fusion_method, self.c_flag, self.l_flag,
                            getattr(self, 'img_spatial_weights', None), getattr(self, 'pts_spatial_weights', None),
                            self._modal_embed())

    # ------------------------------------------------------------------ public
    def encode(self, img_mlvl_feats, pts_mlvl_feats, bev_queries, bev_h, bev_w, bev_pos=None, flags=None, **kwargs):
        """The hot path: backbone features -> fused_bev_embed (B, bev_h*bev_w, C*scale_factor).
        ``flags=(c_flag, l_flag)`` replaces the modality-dropout draw (deterministic replay / tests)."""
        if img_mlvl_feats is None and pts_mlvl_feats is None:
            raise ValueError('at least one of img_mlvl_feats / pts_mlvl_feats is required')
        if img_mlvl_feats is not None and not self.with_img_bev_encoder:
            raise ValueError('image features given but no img_encoder was configured')
        if pts_mlvl_feats is not None and not self.with_pts_bev_encoder:
            raise ValueError('LiDAR features given but no pts_encoder was configured')
        self._draw_flags(img_mlvl_feats, pts_mlvl_feats)
        if flags is not None:
            self.c_flag = int(flags[0]) if img_mlvl_feats is not None else 0
            self.l_flag = int(flags[1]) if pts_mlvl_feats is not None else 0
        grad = torch.is_grad_enabled() and (self.training or any(p.requires_grad for p in self.parameters()))
        if not self.training and not grad and fused_supported(self, img_mlvl_feats, pts_mlvl_feats):
            over = getattr(self, 'fused_overrides', None) or {}     # ablation: {'gemm': ..., 'sampling': ...}
            key = (self.fused_precision, over.get('gemm'), over.get('sampling'))
            if self._fused is None or self._fused_key != key:
                # (derived weight copies are re-built by FusedEncoder.refresh() whenever a parameter changes)
                self._fused, self._fused_key = FusedEncoder(self, self.fused_precision, **over), key
            return self._fused(img_mlvl_feats, pts_mlvl_feats, bev_queries, bev_h, bev_w, bev_pos,
                               kwargs.get('img_metas'), kwargs.get('lidar2img'), kwargs.get('img_shape'))
        if self.training:       # a new step of the in-kernel dropout generator (fused dropout + add + LayerNorm)
            ref = (img_mlvl_feats or pts_mlvl_feats)[0]
            if ref.is_cuda:
                ops.DropoutRNG.get(ref.device).advance()
        img, pts = self._encode_modules(img_mlvl_feats, pts_mlvl_feats, bev_queries, bev_h, bev_w, bev_pos, **kwargs)
        if grad:
            return self._fuse_modules(img, pts)
        return self.fuse(img, pts)

    def _modal_embed(self):
        if self.use_modal_embeds != 'Fixed':
            return None
        if self.fusion_method == 'cat':
            # the reference adds the C-wide embedding to the 2C-wide concatenation and fails (transformer_fusion.py:310-312);
            # the module path fails the same way -- keep the two paths in agreement
            raise ValueError("use_modal_embeds='Fixed' cannot be combined with fusion_method='cat' (C-wide embedding, "
                             "2C-wide features; transformer_fusion.py:310-312)")
        return self.c_flag * self.modal_embbeding_C + self.l_flag * self.modal_embbeding_L

    def forward(self, img_mlvl_feats, pts_mlvl_feats, bev_queries, object_query_embed, bev_h, bev_w, bev_pos=None,
                reg_branches=None, cls_branches=None, **kwargs):
        fused = self.encode(img_mlvl_feats, pts_mlvl_feats, bev_queries, bev_h, bev_w, bev_pos=bev_pos, **kwargs)
        bs = fused.size(0)
        query_pos, query = torch.split(object_query_embed, self.embed_dims * self.scale_factor, dim=1)
        query_pos = query_pos.unsqueeze(0).expand(bs, -1, -1)
        query = query.unsqueeze(0).expand(bs, -1, -1)
        init_reference_out = reference_points = self.reference_points(query_pos).sigmoid()
        fused = fused.permute(1, 0, 2)
        if self.decoder is None:
            return fused, None, init_reference_out, None
        inter_states, inter_references = self.decoder(
            query=query.permute(1, 0, 2), key=None, value=fused, query_pos=query_pos.permute(1, 0, 2),
            reference_points=reference_points, reg_branches=reg_branches, cls_branches=cls_branches,
            spatial_shapes=ops.const_tensor([[bev_h, bev_w]], torch.long, query.device),
            level_start_index=ops.const_tensor([0], torch.long, query.device), **kwargs)
        return fused, inter_states, init_reference_out, inter_references
