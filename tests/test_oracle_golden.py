"""The CPU oracle against the golden vectors frozen from the reference's own code
(tests/golden/make_golden.py) -- this is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import mmcv_semantics as ms
from oracle import unibev_encoder as oe
from tests.helpers import ENCODER_HALF_TAGS, encoder_half_inputs, load_golden, metas_from

TOL = dict(rtol=1e-5, atol=2e-6)      # same fp32 ops, possibly different association


def test_msda_core_matches_transformers_kat():
    a, _ = load_golden('msda_core')
    shapes = [tuple(int(v) for v in r) for r in a['shapes']]
    out = ms.msda_core(a['value'], shapes, a['loc'], a['w'])
    torch.testing.assert_close(out, a['out'], **TOL)


def test_msda_core_scalar_matches_grid_sample_semantics():
    """mmcv's CUDA kernel semantics (per-corner bounds checks) == grid_sample zeros padding."""
    a, _ = load_golden('msda_core')
    shapes = [tuple(int(v) for v in r) for r in a['shapes']]
    sl = slice(0, 12)
    out = ms.msda_core_scalar(a['value'][:1], shapes, a['loc'][:1, sl], a['w'][:1, sl])
    torch.testing.assert_close(out, a['out'][:1, sl], **TOL)


def test_reference_points_and_camera_projection():
    a, _ = load_golden('point_sampling')
    H, W = (int(v) for v in a['bev_hw'])
    D = int(a['D'])
    pc = [float(v) for v in a['pc_range']]
    ref_3d = oe.pillar_points_3d(H, W, pc[5] - pc[2], D, 2)
    ref_2d = oe.grid_points_2d(H, W, 2)
    assert torch.equal(ref_3d, a['ref_3d'])
    assert torch.equal(ref_2d, a['ref_2d'])
    ref_cam, mask = oe.project_to_cameras(ref_3d, pc, metas_from(a))
    assert torch.equal(mask, a['mask'])                    # index path: exact
    torch.testing.assert_close(ref_cam, a['ref_cam'], rtol=1e-6, atol=0)
    assert 0 < int(mask.sum()) < mask.numel()


def test_msda3d_img_module():
    a, p = load_golden('msda3d_img')
    acfg = dict(num_heads=int(a['heads']), num_levels=2, num_points=int(a['points']))
    shapes = [tuple(int(v) for v in r) for r in a['shapes']]
    out = oe.msda3d_forward({'m.' + k: v for k, v in p.items()}, 'm', acfg, a['query'], a['value'], a['ref'], shapes)
    torch.testing.assert_close(out, a['out'], **TOL)


def _prefixed(p, prefix):
    return {prefix + '.' + k: v for k, v in p.items()}


def test_sca_img_module_batch0_quirk():
    a, p = load_golden('sca_img')
    acfg = dict(deformable_attention=dict(num_heads=int(a['heads']), num_levels=1, num_points=int(a['points'])))
    shapes = [tuple(int(v) for v in r) for r in a['shapes']]
    out = oe.sca_img_forward(_prefixed(p, 'm'), 'm', acfg, a['query'], a['feats'], a['feats'], a['ref_cam'],
                             a['mask'], shapes)
    torch.testing.assert_close(out, a['out'], **TOL)
    # the fixture really exercises the quirk: item 1's mask differs from item 0's
    assert not torch.equal(a['mask'][:, 0], a['mask'][:, 1])


def test_sca_pts_module():
    a, p = load_golden('sca_pts')
    acfg = dict(deformable_attention=dict(num_heads=int(a['heads']), num_levels=1, num_points=int(a['points'])))
    shapes = [tuple(int(v) for v in r) for r in a['shapes']]
    out = oe.sca_pts_forward(_prefixed(p, 'm'), 'm', acfg, a['query'], a['feats'], a['feats'], a['ref_lidar'], shapes)
    torch.testing.assert_close(out, a['out'], **TOL)


@pytest.mark.parametrize('tag', ENCODER_HALF_TAGS)
def test_encoder_half(tag):
    a, p = load_golden('encoder_half_' + tag)
    cfg, img, pts, q, bev_h, bev_w = encoder_half_inputs(a)
    flags = tuple(int(v) for v in a['flags'])
    fused, (img_e, pts_e) = oe.encoder_half(p, cfg, img, pts, q, bev_h, bev_w, bev_pos=a['bev_pos'],
                                            img_metas=metas_from(a), flags=flags, return_parts=True)
    tol = dict(rtol=1e-4, atol=2e-5)      # 2 layers x (5 linears + 3 LN) deep
    if img_e is not None:
        torch.testing.assert_close(img_e, a['img_bev_embed'], **tol)
    if pts_e is not None:
        torch.testing.assert_close(pts_e, a['pts_bev_embed'], **tol)
    torch.testing.assert_close(fused, a['fused'], **tol)


def test_dropflags_fixtures_cover_both_branches():
    f1 = load_golden('encoder_half_lc_cnw_dropflags')[0]['flags'].tolist()
    f2 = load_golden('encoder_half_lc_cnw_dropdict')[0]['flags'].tolist()
    assert sorted([tuple(f1), tuple(f2)]) == [(0, 1), (1, 0)] or {tuple(f1), tuple(f2)} <= {(0, 1), (1, 0)}


def test_c_oracle_matches_kat_and_torch_oracle():
    """The plain-C restatement of mmcv's CUDA-kernel semantics (oracle/msda_core.c) against the transformers-made
    known-answer vectors, and against the grid_sample oracle on adversarial coordinates (outside, on borders,
    exactly on pixel centres, huge magnitudes from z < eps projections, NaN-free)."""
    from oracle import c_msda
    a, _ = load_golden('msda_core')
    shapes = [tuple(int(v) for v in r) for r in a['shapes']]
    got = c_msda.msda_forward(a['value'].numpy(), shapes, a['loc'].numpy(), a['w'].numpy())
    np.testing.assert_allclose(got, a['out'].numpy(), rtol=1e-5, atol=2e-6)
    g = torch.Generator().manual_seed(11)
    shapes = [(5, 7), (3, 2)]
    B, H, D, Nq, P = 2, 3, 4, 40, 4
    value = torch.randn(B, sum(h * w for h, w in shapes), H, D, generator=g)
    loc = torch.rand(B, Nq, H, len(shapes), P, 2, generator=g) * 1.6 - 0.3
    loc[0, :6] = torch.tensor([0.0, 1.0, 0.5, -1e7, 1e7, 0.1])[:, None, None, None, None]      # borders / blow-ups
    loc[1, :5, :, 0] = (torch.arange(5)[:, None, None, None] + 0.5) / 7.0                        # exact pixel centres
    w = torch.rand(B, Nq, H, len(shapes), P, generator=g)
    want = ms.msda_core(value, shapes, loc, w)
    got = c_msda.msda_forward(value.numpy(), shapes, loc.numpy(), w.numpy())
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-5, atol=2e-6)


def test_config1_camera_only_cpu_plumbing():
    """BASELINE configs[0]: unibev_nus_C (camera only, 6 x (3 x 256 x 704) -> 8 x 22 stride-32 maps, 1 sample) end to end on
    the CPU through the oracle, with the parameters of the plugin model built from the same config subtree."""
    from unibev_b200 import synth
    model, cfg = synth.build_model('unibev_nus_C')
    inp = synth.make_inputs('unibev_nus_C', batch=1)
    params = {k: v.detach() for k, v in model.state_dict().items()}
    with torch.no_grad():
        out = oe.encoder_half(params, cfg, inp['img_feats'], None, inp['bev_queries'], inp['bev_h'], inp['bev_w'],
                              bev_pos=inp['bev_pos'], img_metas=inp['img_metas'])
    assert out.shape == (1, 200 * 200, 256) and bool(torch.isfinite(out).all())
    assert float(out.std()) > 0.1                       # LayerNorm'd BEV features, not a degenerate map
