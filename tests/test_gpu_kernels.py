"""GPU parity of every C-ABI entry point against the CPU oracle / golden vectors.
All calls go through libunibev_b200.so (ctypes)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mmcv_semantics as ms
from oracle import unibev_encoder as oe
from tests.helpers import load_golden, metas_from

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from unibev_b200 import ops as _ops
    return _ops


def _lsi(shapes):
    s = torch.tensor(shapes, dtype=torch.long)
    return s, torch.cat((s.new_zeros(1), s.prod(1).cumsum(0)[:-1]))


def test_msda_fwd_golden_kat(ops):
    a, _ = load_golden('msda_core')
    shapes, lsi = _lsi(a['shapes'].tolist())
    out = ops.msda_forward(a['value'].cuda(), shapes.cuda(), lsi.cuda(), a['loc'].cuda(), a['w'].cuda())
    torch.testing.assert_close(out.cpu(), a['out'], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize('B,H,D,Nq,P,shapes', [
    (2, 8, 32, 300, 8, [(29, 50)]),
    (1, 8, 32, 257, 4, [(20, 20), (10, 10), (5, 5), (3, 3)]),
    (3, 8, 16, 100, 8, [(13, 7)]),
    (2, 4, 8, 65, 4, [(6, 9), (3, 5)]),
    (1, 2, 64, 33, 2, [(8, 8)]),
    (1, 1, 128, 9, 1, [(4, 4)]),
    (2, 3, 4, 50, 3, [(5, 6)]),
    (2, 3, 12, 41, 3, [(5, 6), (2, 2)]),      # head dim 12: scalar path
    (1, 5, 1, 17, 2, [(3, 3)]),               # head dim 1: scalar path
])
def test_msda_fwd_bwd_vs_oracle(ops, B, H, D, Nq, P, shapes):
    g = torch.Generator().manual_seed(B * 1000 + D)
    L = len(shapes)
    Nv = sum(h * w for h, w in shapes)
    value = torch.randn(B, Nv, H, D, generator=g)
    loc = torch.rand(B, Nq, H, L, P, 2, generator=g) * 1.3 - 0.15
    w = torch.rand(B, Nq, H, L, P, generator=g) + 0.1
    w = w / w.sum((-1, -2), keepdim=True)
    go = torch.randn(B, Nq, H * D, generator=g)
    # oracle fwd + autograd bwd on CPU
    v_c, l_c, w_c = value.clone().requires_grad_(), loc.clone().requires_grad_(), w.clone().requires_grad_()
    want = ms.msda_core(v_c, shapes, l_c, w_c)
    want.backward(go)
    s, lsi = _lsi(shapes)
    v_g, l_g, w_g = (t.cuda().requires_grad_() for t in (value, loc, w))
    got = ops.MultiScaleDeformableAttnFunction.apply(v_g, s.cuda(), lsi.cuda(), l_g, w_g, 64)
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-5)
    got.backward(go.cuda())
    torch.testing.assert_close(v_g.grad.cpu(), v_c.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(w_g.grad.cpu(), w_c.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(l_g.grad.cpu(), l_c.grad, rtol=1e-4, atol=2e-4)


def test_msda_rejects_bad_input(ops):
    v = torch.randn(1, 4, 2, 8)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.msda_forward(v, torch.tensor([[2, 2]]), torch.tensor([0]), torch.rand(1, 3, 2, 1, 2, 2),
                         torch.rand(1, 3, 2, 1, 2))
    with pytest.raises(ValueError):
        ops.msda_forward(v.cuda(), torch.tensor([[2, 2]]).cuda(), torch.tensor([0]).cuda(),
                         torch.rand(1, 3, 2, 1, 2, 2).cuda(), torch.rand(1, 3, 2, 1, 3).cuda())


def test_project_points_golden(ops):
    """Index path: visibility bits exact; coordinates to 1 ulp-ish (reference: ImgEncoder.point_sampling)."""
    a, _ = load_golden('point_sampling')
    H, W = (int(v) for v in a['bev_hw'])
    D = int(a['D'])
    pc = [float(v) for v in a['pc_range']]
    ih, iw = (int(v) for v in a['img_hw'])
    zs = (torch.linspace(0.5, (pc[5] - pc[2]) - 0.5, D) / (pc[5] - pc[2])).tolist()
    ref, mask = ops.project_points(a['lidar2img'].float().cuda(), zs, pc, ih, iw, H, W)
    bits = ((mask.cpu()[..., None] >> torch.arange(D, dtype=torch.uint8)) & 1).bool()     # (B, Nq, N, D)
    assert torch.equal(bits.permute(2, 0, 1, 3), a['mask'])
    got, want = ref.cpu().permute(2, 0, 1, 3, 4), a['ref_cam']
    # points near the camera plane (depth clamped at 1e-5) land millions of pixels away and are ill-conditioned
    # (catastrophic cancellation in the depth row); they can never be sampled, so only their magnitude matters
    near = want.abs().amax(-1, keepdim=True).expand_as(want) < 8.0
    torch.testing.assert_close(got[near], want[near], rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(got[~near], want[~near], rtol=2e-3, atol=0)


def _qproj(p, prefix, query):
    off = F.linear(query, p[prefix + '.sampling_offsets.weight'], p[prefix + '.sampling_offsets.bias'])
    lg = F.linear(query, p[prefix + '.attention_weights.weight'], p[prefix + '.attention_weights.bias'])
    return torch.cat((off, lg), -1)


def test_bev_sample_vs_reference_sca_pts(ops):
    """ub_bev_sample_fwd fed with CPU-computed projections must reproduce the reference's
    SpatialCrossAttentionPts inner attention (before output_proj)."""
    a, p = load_golden('sca_pts')
    H, P = int(a['heads']), int(a['points'])
    fh, fw = (int(v) for v in a['shapes'][0])
    query = a['query']                                   # (B, Nq, C), Nq = 8*10 grid
    B, Nq, C = query.shape
    value = F.linear(a['feats'].permute(1, 0, 2), p['deformable_attention.value_proj.weight'],
                     p['deformable_attention.value_proj.bias'])
    qp = _qproj(p, 'deformable_attention', query)
    got = ops.bev_sample(value.cuda(), qp.cuda(), 8, 10, fh, fw, H, P, 0, H * P * 2).cpu()
    want_inner = oe.msda3d_forward({'m.' + k: v for k, v in p.items()}, 'm.deformable_attention',
                                   dict(num_heads=H, num_levels=1, num_points=P), query, a['feats'].permute(1, 0, 2),
                                   a['ref_lidar'].permute(1, 2, 0, 3), [(fh, fw)])
    torch.testing.assert_close(got, want_inner, rtol=1e-5, atol=1e-5)
    full = F.linear(got, p['output_proj.weight'], p['output_proj.bias']) + query
    torch.testing.assert_close(full, a['out'], rtol=1e-4, atol=1e-5)


def test_img_sample_vs_reference_sca_img(ops):
    """ub_project_points + ub_img_sample_fwd reproduce the reference's rebatch/sample/scatter/count,
    including hit lists taken from batch item 0 (the two items have different calibrations)."""
    a, p = load_golden('sca_img')
    H, P, N = int(a['heads']), int(a['points']), int(a['cams'])
    fh, fw = (int(v) for v in a['shapes'][0])
    query = a['query']
    B, Nq, C = query.shape
    D = a['ref_cam'].shape[3]
    value = F.linear(a['feats'].permute(2, 0, 1, 3), p['deformable_attention.value_proj.weight'],
                     p['deformable_attention.value_proj.bias'])                       # (B, N, hw, C)
    qp = _qproj(p, 'deformable_attention', query)
    ref = a['ref_cam'].permute(1, 2, 0, 3, 4).contiguous()                               # (B, Nq, N, D, 2)
    bits = (a['mask'].permute(1, 2, 0, 3).to(torch.uint8) << torch.arange(D, dtype=torch.uint8)).sum(-1).to(torch.uint8)
    got = ops.img_sample(value.cuda(), qp.cuda(), ref.cuda(), bits.cuda(), 8, 10, fh, fw, H, P, 0, H * P * 2).cpu()
    full = F.linear(got, p['output_proj.weight'], p['output_proj.bias']) + query
    torch.testing.assert_close(full, a['out'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('rows,C', [(1, 32), (77, 128), (1000, 256), (333, 512), (64, 1024), (5, 36)])
def test_add_layernorm(ops, rows, C):
    g = torch.Generator().manual_seed(C)
    x, r = torch.randn(rows, C, generator=g) * 3, torch.randn(rows, C, generator=g)
    b, gm, bt = torch.randn(C, generator=g), torch.randn(C, generator=g), torch.randn(C, generator=g)
    want = F.layer_norm(x + b + r, (C,), gm, bt, 1e-5)
    got = ops.add_layernorm(x.cuda(), gm.cuda(), bt.cuda(), bias=b.cuda(), residual=r.cuda()).cpu()
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    want2 = F.layer_norm(x, (C,), gm, bt, 1e-5)
    xg = x.cuda()
    got2 = ops.add_layernorm(xg, gm.cuda(), bt.cuda(), out=xg).cpu()          # in place, no bias / residual
    torch.testing.assert_close(got2, want2, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('mode', ['linear', 'avg', 'cat'])
@pytest.mark.parametrize('flags', [(1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize('cnw,spatial', [(True, False), (False, False), (True, True)])
def test_cnw_fuse(ops, mode, flags, cnw, spatial):
    g = torch.Generator().manual_seed(5)
    B, Nq, C = 2, 50, 32
    img, pts = torch.randn(B, Nq, C, generator=g), torch.randn(B, Nq, C, generator=g)
    p = dict(img_channel_weights=torch.randn(C, generator=g), pts_channel_weights=torch.randn(C, generator=g),
             img_spatial_weights=torch.randn(Nq, generator=g), pts_spatial_weights=torch.randn(Nq, generator=g))
    cfg = dict(fusion_method=mode, feature_norm='ChannelNormWeights' if cnw else None,
               spatial_norm='SpatialNormWeights' if spatial else None)
    c, l = flags
    i2, p2 = oe.channel_norm_weights(p, cfg, img, pts, c, l)
    i2, p2 = oe.spatial_norm_weights(p, cfg, i2, p2, c, l)
    want = oe.fuse(p, cfg, i2, p2, c, l)
    cu = lambda t: t.cuda()
    got = ops.cnw_fuse(cu(img), cu(pts), cu(p['img_channel_weights']) if cnw else None,
                       cu(p['pts_channel_weights']) if cnw else None, mode, c, l,
                       cu(p['img_spatial_weights']) if spatial else None,
                       cu(p['pts_spatial_weights']) if spatial else None).cpu()
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)


def test_cnw_fuse_missing_modality(ops):
    g = torch.Generator().manual_seed(6)
    pts = torch.randn(1, 20, 16, generator=g)
    wi, wp = torch.randn(16, generator=g), torch.randn(16, generator=g)
    got = ops.cnw_fuse(None, pts.cuda(), wi.cuda(), wp.cuda(), 'linear', 0, 1).cpu()
    torch.testing.assert_close(got, pts, rtol=0, atol=0)        # softmax over a single row == 1 (fusion:332-334)


@pytest.mark.parametrize('G,C,h,w', [(6, 256, 29, 50), (2, 32, 9, 9), (3, 33, 5, 7)])
def test_flatten_feats(ops, G, C, h, w):
    g = torch.Generator().manual_seed(G)
    feat = torch.randn(G, C, h, w, generator=g)
    ea, eb = torch.randn(3, C, generator=g), torch.randn(C, generator=g)
    want = (feat.flatten(2).permute(0, 2, 1) + ea[torch.arange(G) % 3][:, None]) + eb
    got = ops.flatten_feats(feat.cuda(), ea.cuda(), eb.cuda()).cpu()
    assert torch.equal(got, want)
    assert torch.equal(ops.flatten_feats(feat.cuda()).cpu(), feat.flatten(2).permute(0, 2, 1))


@pytest.mark.parametrize('G,C,h,w', [(6, 256, 29, 50), (2, 32, 9, 9), (3, 33, 5, 7), (1, 130, 4, 40)])
def test_flatten_feats_fp16_outputs(ops, G, C, h, w):
    """fp16 rows next to / instead of the fp32 rows: the fp32 result rounded once (bit-exact), on both kernels (the
    64-channel fp16-only kernel needs an even C; odd C takes the generic one)."""
    g = torch.Generator().manual_seed(G + C)
    feat = torch.randn(G, C, h, w, generator=g)
    ea, eb = torch.randn(3, C, generator=g), torch.randn(C, generator=g)
    want = ((feat.flatten(2).permute(0, 2, 1) + ea[torch.arange(G) % 3][:, None]) + eb).half()
    both32, both16 = ops.flatten_feats(feat.cuda(), ea.cuda(), eb.cuda(), fp32=True, fp16=True)
    none32, only16 = ops.flatten_feats(feat.cuda(), ea.cuda(), eb.cuda(), fp32=False, fp16=True)
    assert none32 is None and both32.dtype == torch.float32 and only16.dtype == torch.float16
    assert torch.equal(both16.cpu(), want) and torch.equal(only16.cpu(), want)
    assert torch.equal(ops.flatten_feats(feat.cuda(), fp32=False, fp16=True)[1].cpu(), feat.flatten(2).permute(0, 2, 1).half())


def test_broadcast_rows(ops):
    g = torch.Generator().manual_seed(0)
    src = torch.randn(1000, 256, generator=g)
    o32, o16 = ops.broadcast_rows(src.cuda(), 3)
    assert torch.equal(o32.cpu(), src.unsqueeze(0).expand(3, -1, -1)) and torch.equal(o16.cpu(), src.half().unsqueeze(0).expand(3, -1, -1))
    o32, o16 = ops.broadcast_rows(src.cuda(), 2, fp32=False)
    assert o32 is None and torch.equal(o16.cpu()[1], src.half())
