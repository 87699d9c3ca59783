"""``UniBEV_Head`` -- the producers and consumers either side of the hot path.

Front (unibev_head.py:120-135, 164-182): the learnable BEV query table(s) ``bev_embedding`` (or ``bev_embedding_img`` /
``bev_embedding_pts`` with ``dual_queries``), the object queries ``query_embedding`` (num_query, 2 * C * scale_factor) and
``bev_pos = positional_encoding(zeros(bs, bev_h, bev_w))``.  Back (unibev_head.py:91-112, 200-242): per-decoder-layer
classification / regression branches, iterative box refinement against the inverse-sigmoid reference points, metric
de-normalisation of x / y / z with ``pc_range``; ``get_bboxes`` -> ``NMSFreeCoder.decode`` (+ the gravity-centre shift,
:526-529).  Same registered name, constructor keywords and parameter names (``cls_branches.{i}.{0,1,3,4,6}``,
``reg_branches.{i}.{0,2,4}``, ``bev_embedding.weight``, ``query_embedding.weight``, ``positional_encoding.*``,
``code_weights``, ``transformer.*``), so reference checkpoints load under ``pts_bbox_head.``.

Out of scope here: losses / target assignment (``loss``, ``_get_target_single``: training of the head, not the hot path)
and ``as_two_stage`` (never enabled in a UniBEV config).  The arithmetic of the branches is a handful of (bs * 900)-row
linears on the decoder output -- torch glue; the hot path is ``self.transformer``.
"""
import copy
import math

import torch
import torch.nn as nn

from ..registry import HAVE_MMDET, HEADS, build_bbox_coder, build_positional_encoding, build_transformer
from .decoder import inverse_sigmoid


def bias_init_with_prob(prior_prob):
    return float(-math.log((1 - prior_prob) / prior_prob))


@HEADS.register_module(name='UniBEV_Head', force=HAVE_MMDET)
class UniBEVHead(nn.Module):
    def __init__(self, num_classes, in_channels, num_query=100, num_reg_fcs=2, transformer=None, sync_cls_avg_factor=False,
                 positional_encoding=None, loss_cls=None, loss_bbox=None, loss_iou=None, train_cfg=None, test_cfg=None,
                 with_box_refine=False, as_two_stage=False, bbox_coder=None, num_cls_fcs=2, code_weights=None, bev_h=30,
                 bev_w=30, code_size=10, init_cfg=None, **kwargs):
        super().__init__()
        if as_two_stage:
            raise NotImplementedError('as_two_stage is not used by any UniBEV config and is not supported')
        self.bev_h, self.bev_w = bev_h, bev_w
        self.fp16_enabled = False
        self.with_box_refine, self.as_two_stage = with_box_refine, as_two_stage
        self.scale_factor = 2 if transformer['fusion_method'] == 'cat' else 1     # unibev_head.py:59-62
        self.dual_queries = transformer.get('dual_queries', False)
        self.code_size = code_size
        self.num_classes, self.in_channels, self.num_query, self.num_reg_fcs = num_classes, in_channels, num_query, num_reg_fcs
        self.sync_cls_avg_factor = sync_cls_avg_factor
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.pc_range = self.bbox_coder.pc_range
        self.real_w = self.pc_range[3] - self.pc_range[0]
        self.real_h = self.pc_range[4] - self.pc_range[1]
        self.num_cls_fcs = num_cls_fcs - 1
        # mmdet DETRHead.__init__: sigmoid focal loss -> one logit per class (no background channel)
        self.use_sigmoid_cls = bool((loss_cls or {}).get('use_sigmoid', False))
        self.cls_out_channels = num_classes if self.use_sigmoid_cls else num_classes + 1
        self.positional_encoding = build_positional_encoding(positional_encoding)
        self.transformer = build_transformer(transformer)
        self.embed_dims = self.transformer.embed_dims
        self._init_layers()
        weights = code_weights if code_weights is not None else [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2]
        self.code_weights = nn.Parameter(torch.tensor(weights), requires_grad=False)

    def _init_layers(self):
        """unibev_head.py:91-135."""
        C = self.embed_dims * self.scale_factor
        cls_branch = []
        for _ in range(self.num_reg_fcs):
            cls_branch += [nn.Linear(C, C), nn.LayerNorm(C), nn.ReLU(inplace=True)]
        cls_branch.append(nn.Linear(C, self.cls_out_channels))
        fc_cls = nn.Sequential(*cls_branch)
        reg_branch = []
        for _ in range(self.num_reg_fcs):
            reg_branch += [nn.Linear(C, C), nn.ReLU()]
        reg_branch.append(nn.Linear(C, self.code_size))
        reg_branch = nn.Sequential(*reg_branch)
        num_pred = self.transformer.decoder.num_layers if self.transformer.decoder is not None else 0
        if self.with_box_refine:
            self.cls_branches = nn.ModuleList([copy.deepcopy(fc_cls) for _ in range(num_pred)])
            self.reg_branches = nn.ModuleList([copy.deepcopy(reg_branch) for _ in range(num_pred)])
        else:
            self.cls_branches = nn.ModuleList([fc_cls for _ in range(num_pred)])
            self.reg_branches = nn.ModuleList([reg_branch for _ in range(num_pred)])
        if self.dual_queries:
            self.bev_embedding_img = nn.Embedding(self.bev_h * self.bev_w, self.embed_dims)
            self.bev_embedding_pts = nn.Embedding(self.bev_h * self.bev_w, self.embed_dims)
        else:
            self.bev_embedding = nn.Embedding(self.bev_h * self.bev_w, self.embed_dims)
        self.query_embedding = nn.Embedding(self.num_query, self.embed_dims * 2 * self.scale_factor)

    def init_weights(self):
        """unibev_head.py:137-143."""
        self.transformer.init_weights()
        if self.use_sigmoid_cls:
            bias_init = bias_init_with_prob(0.01)
            for m in self.cls_branches:
                nn.init.constant_(m[-1].bias, bias_init)

    def bev_inputs(self, bs, dtype, device=None):
        """-> (bev_queries, object_query_embeds, bev_pos): unibev_head.py:171-182."""
        object_query_embeds = self.query_embedding.weight.to(dtype)
        if self.dual_queries:
            bev_queries = [self.bev_embedding_img.weight.to(dtype), self.bev_embedding_pts.weight.to(dtype)]
            device = bev_queries[0].device
        else:
            bev_queries = self.bev_embedding.weight.to(dtype)
            device = bev_queries.device
        bev_mask = torch.zeros((bs, self.bev_h, self.bev_w), device=device).to(dtype)
        return bev_queries, object_query_embeds, self.positional_encoding(bev_mask).to(dtype)

    def forward(self, mlvl_img_feats, pts_feats, img_metas, **kwargs):
        """-> dict(bev_embed, all_cls_scores (nb_dec, bs, num_query, classes), all_bbox_preds (nb_dec, bs, num_query, code),
        enc_cls_scores=None, enc_bbox_preds=None): unibev_head.py:145-255."""
        ref = mlvl_img_feats[0] if mlvl_img_feats is not None else pts_feats[0]
        bs, dtype = ref.shape[0], ref.dtype
        bev_queries, object_query_embeds, bev_pos = self.bev_inputs(bs, dtype)
        bev_embed, hs, init_reference, inter_references = self.transformer(
            mlvl_img_feats, pts_feats, bev_queries, object_query_embeds, self.bev_h, self.bev_w,
            grid_length=(self.real_h / self.bev_h, self.real_w / self.bev_w), bev_pos=bev_pos,
            reg_branches=self.reg_branches if self.with_box_refine else None,
            cls_branches=self.cls_branches if self.as_two_stage else None, img_metas=img_metas, **kwargs)
        if hs is None:
            raise RuntimeError('UniBEV_Head needs a transformer with a decoder')
        hs = hs.permute(0, 2, 1, 3)
        classes, coords = [], []
        pc = self.pc_range
        for lvl in range(hs.shape[0]):
            reference = inverse_sigmoid(init_reference if lvl == 0 else inter_references[lvl - 1])
            assert reference.shape[-1] == 3
            tmp = self.reg_branches[lvl](hs[lvl])
            xy = (tmp[..., 0:2] + reference[..., 0:2]).sigmoid()
            z = (tmp[..., 4:5] + reference[..., 2:3]).sigmoid()
            tmp = torch.cat((xy[..., 0:1] * (pc[3] - pc[0]) + pc[0], xy[..., 1:2] * (pc[4] - pc[1]) + pc[1], tmp[..., 2:4],
                             z * (pc[5] - pc[2]) + pc[2], tmp[..., 5:]), -1)
            classes.append(self.cls_branches[lvl](hs[lvl]))
            coords.append(tmp)
        return {'bev_embed': bev_embed, 'all_cls_scores': torch.stack(classes), 'all_bbox_preds': torch.stack(coords),
                'enc_cls_scores': None, 'enc_bbox_preds': None}

    def get_bboxes(self, preds_dicts, img_metas, rescale=False):
        """-> per sample [bboxes, scores, labels]; boxes go from gravity centre to bottom centre (z -= h / 2) and into
        ``img_metas[i]['box_type_3d']`` when the meta carries one (unibev_head.py:511-538)."""
        ret_list = []
        for i, preds in enumerate(self.bbox_coder.decode(preds_dicts)):
            bboxes = preds['bboxes'].clone()
            bboxes[:, 2] = bboxes[:, 2] - bboxes[:, 5] * 0.5
            box_type = img_metas[i].get('box_type_3d') if img_metas is not None else None
            if box_type is not None:
                bboxes = box_type(bboxes, bboxes.shape[-1])
            ret_list.append([bboxes, preds['scores'], preds['labels']])
        return ret_list
