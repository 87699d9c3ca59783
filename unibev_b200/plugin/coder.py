"""``NMSFreeCoder`` -- box decoding behind the head (output side of the hot path).

Same registered name, constructor keywords and ``decode`` / ``decode_single`` results as
projects/UniBEV/unibev_plugin/core/bbox/coders/nms_free_coder.py:9-121 (+ ``denormalize_bbox``, core/bbox/util.py:27-55):
sigmoid scores, top-``max_num`` over (query, class), label = index % num_classes, box = index // num_classes, boxes
de-normalised (exp of the log sizes, atan2 of the sine / cosine pair), score threshold with the reference's 0.9-decay
fallback, centre-range mask.  Batch results are a list of dicts with ``bboxes`` / ``scores`` / ``labels``.
"""
import torch

from ..registry import BBOX_CODERS, HAVE_MMDET


def denormalize_bbox(normalized_bboxes, pc_range=None):
    """(cx, cy, log w, log l, cz, log h, sin, cos[, vx, vy]) -> (cx, cy, cz, w, l, h, rot[, vx, vy]) (util.py:27-55)."""
    rot = torch.atan2(normalized_bboxes[..., 6:7], normalized_bboxes[..., 7:8])
    cx, cy, cz = normalized_bboxes[..., 0:1], normalized_bboxes[..., 1:2], normalized_bboxes[..., 4:5]
    w, l, h = normalized_bboxes[..., 2:3].exp(), normalized_bboxes[..., 3:4].exp(), normalized_bboxes[..., 5:6].exp()
    if normalized_bboxes.size(-1) > 8:
        return torch.cat([cx, cy, cz, w, l, h, rot, normalized_bboxes[:, 8:9], normalized_bboxes[:, 9:10]], dim=-1)
    return torch.cat([cx, cy, cz, w, l, h, rot], dim=-1)


@BBOX_CODERS.register_module(force=HAVE_MMDET)
class NMSFreeCoder:
    def __init__(self, pc_range, voxel_size=None, post_center_range=None, max_num=100, score_threshold=None,
                 num_classes=10):
        self.pc_range, self.voxel_size = pc_range, voxel_size
        self.post_center_range = post_center_range
        self.max_num, self.score_threshold, self.num_classes = max_num, score_threshold, num_classes

    def encode(self):
        pass

    def decode_single(self, cls_scores, bbox_preds):
        """cls_scores (num_query, num_classes) logits, bbox_preds (num_query, code_size) -> dict of the kept boxes
        (nms_free_coder.py:40-100)."""
        scores, indexs = cls_scores.sigmoid().view(-1).topk(self.max_num)
        labels = indexs % self.num_classes
        bbox_index = torch.div(indexs, self.num_classes, rounding_mode='floor')
        final_box_preds = denormalize_bbox(bbox_preds[bbox_index], self.pc_range)
        thresh_mask = None
        if self.score_threshold is not None:
            thresh_mask = scores > self.score_threshold
            tmp_score = self.score_threshold
            while thresh_mask.sum() == 0:             # nothing above the threshold: relax it by 0.9 per round (:68-75)
                tmp_score *= 0.9
                if tmp_score < 0.01:
                    thresh_mask = scores > -1
                    break
                thresh_mask = scores >= tmp_score
        if self.post_center_range is None:
            raise NotImplementedError('Need to reorganize output as a batch, only support post_center_range is not None '
                                      'for now!')
        rng = torch.as_tensor(self.post_center_range, device=scores.device, dtype=final_box_preds.dtype)
        mask = (final_box_preds[..., :3] >= rng[:3]).all(1) & (final_box_preds[..., :3] <= rng[3:]).all(1)
        if self.score_threshold:
            mask &= thresh_mask
        return {'bboxes': final_box_preds[mask], 'scores': scores[mask], 'labels': labels[mask]}

    def decode(self, preds_dicts):
        """preds_dicts['all_cls_scores'] (nb_dec, bs, num_query, classes), ['all_bbox_preds'] (nb_dec, bs, num_query, code):
        the LAST decoder layer is decoded, one dict per sample (:102-121)."""
        all_cls_scores = preds_dicts['all_cls_scores'][-1]
        all_bbox_preds = preds_dicts['all_bbox_preds'][-1]
        return [self.decode_single(all_cls_scores[i], all_bbox_preds[i]) for i in range(all_cls_scores.size(0))]
