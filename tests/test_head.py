"""UniBEV_Head (query / positional-encoding producers, cls / reg branches with box refinement) and
NMSFreeCoder against golden vectors frozen from the reference's OWN unibev_head.py / nms_free_coder.py
(tests/golden/make_golden.py::golden_head: reference classes unmodified, mmdet's DETRHead / LearnedPositionalEncoding /
BaseBBoxCoder stubbed by their published behaviour)."""
import json

import pytest
import torch

from tests.helpers import load_golden

TAGS = ['linear', 'cat_dual_thresh']


def _head(tag, device='cpu'):
    import unibev_b200.plugin  # noqa: F401
    from unibev_b200.registry import build_head
    a, p = load_golden('head_' + tag)
    cfg = json.loads(str(a['cfg_json']))
    head = build_head(cfg)
    head.load_state_dict(p, strict=True)          # same parameter names as the reference head (checkpoint compatible)
    return a, cfg, head.to(device).eval()


def _metas(a):
    l2i = a['lidar2img'].numpy()
    h, w = (int(v) for v in a['img_hw'])
    return [dict(lidar2img=[l2i[b, n] for n in range(l2i.shape[1])], img_shape=[(h, w, 3)] * l2i.shape[1])
            for b in range(l2i.shape[0])]


@pytest.mark.parametrize('tag', TAGS)
def test_state_dict_and_producers_match_reference(tag):
    a, cfg, head = _head(tag)
    C, scale = cfg['in_channels'], 2 if cfg['transformer']['fusion_method'] == 'cat' else 1
    Nq = cfg['bev_h'] * cfg['bev_w']
    q, obj, pos = head.bev_inputs(2, torch.float32)
    if cfg['transformer'].get('dual_queries'):
        assert [tuple(t.shape) for t in q] == [(Nq, C), (Nq, C)]
    else:
        assert tuple(q.shape) == (Nq, C)
    assert tuple(obj.shape) == (cfg['num_query'], 2 * C * scale)
    assert tuple(pos.shape) == (2, C, cfg['bev_h'], cfg['bev_w'])
    # bev_pos[b, :C/2, y, x] = col_embed[x], bev_pos[b, C/2:, y, x] = row_embed[y]
    assert torch.equal(pos[1, :C // 2, 3, 5], head.positional_encoding.col_embed.weight[5])
    assert torch.equal(pos[0, C // 2:, 3, 5], head.positional_encoding.row_embed.weight[3])
    assert len(head.cls_branches) == len(head.reg_branches) == cfg['transformer']['decoder']['num_layers']
    assert head.cls_branches[0] is not head.cls_branches[1]          # with_box_refine: independent clones


@pytest.mark.parametrize('tag', TAGS)
def test_coder_and_get_bboxes_match_reference(tag):
    """NMSFreeCoder.decode + the gravity-centre shift of get_bboxes on the reference's own head outputs: exact."""
    a, cfg, head = _head(tag)
    preds = {'all_cls_scores': a['all_cls_scores'], 'all_bbox_preds': a['all_bbox_preds']}
    got = head.get_bboxes(preds, _metas(a))
    assert len(got) == a['all_cls_scores'].shape[1]
    kept = 0
    for i, (boxes, scores, labels) in enumerate(got):
        torch.testing.assert_close(boxes, a[f'det{i}_bboxes'], rtol=0, atol=0)
        torch.testing.assert_close(scores, a[f'det{i}_scores'], rtol=0, atol=0)
        assert torch.equal(labels, a[f'det{i}_labels'])
        assert boxes.shape[1] == 9
        kept += boxes.shape[0]
    assert 0 < kept
    if cfg['bbox_coder'].get('score_threshold'):
        assert kept < cfg['bbox_coder']['max_num'] * len(got)        # the threshold / centre-range masks did drop boxes


def test_coder_threshold_decay_and_errors():
    from unibev_b200.plugin import NMSFreeCoder
    rng = [-10, -10, -10, 10, 10, 10]
    scores = torch.full((5, 10), -3.0)                               # sigmoid ~ 0.047 everywhere: nothing above 0.5
    boxes = torch.zeros(5, 10)
    boxes[:, 7] = 1.0
    out = NMSFreeCoder(rng, post_center_range=rng, max_num=4, score_threshold=0.5).decode_single(scores, boxes)
    assert out['bboxes'].shape == (4, 9)                             # threshold relaxed by 0.9 per round until boxes remain
    with pytest.raises(NotImplementedError):
        NMSFreeCoder(rng, post_center_range=None, max_num=4).decode_single(scores, boxes)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', TAGS)
def test_gpu_head_forward_matches_reference(tag):
    """Backbone features -> encoders -> fusion -> decoder -> branches on the GPU (fused eval pipeline + ub_msda_fwd in the
    decoder) against the reference head's outputs."""
    a, cfg, head = _head(tag, 'cuda')
    with torch.no_grad():
        outs = head([a['img_feats'].cuda()], [a['pts_feats'].cuda()], _metas(a))
    assert head.transformer._fused is not None
    torch.testing.assert_close(outs['bev_embed'].cpu(), a['bev_embed'], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(outs['all_cls_scores'].cpu(), a['all_cls_scores'], rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(outs['all_bbox_preds'].cpu(), a['all_bbox_preds'], rtol=1e-3, atol=2e-4)
    dets = head.get_bboxes({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in outs.items()}, _metas(a))
    for i, (boxes, scores, labels) in enumerate(dets):
        assert torch.equal(labels, a[f'det{i}_labels'])              # same boxes kept, in the same order
        torch.testing.assert_close(boxes, a[f'det{i}_bboxes'], rtol=1e-3, atol=2e-4)
