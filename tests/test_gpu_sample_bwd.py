"""Backward twins of the fused sampling kernels (ub_bev_sample_bwd / ub_img_sample_bwd, through the C ABI) against
autograd through the CPU oracle: the same raw offset | logit rows are turned into sampling locations / softmax weights with
torch ops (spatial_cross_attention_img.py:390-419, spatial_cross_attention_pts.py:396-426; camera mode also the hit /
count logic of :141-212) and pushed through the oracle's MSDA core."""
import pytest
import torch

from oracle import mmcv_semantics as ms

pytestmark = pytest.mark.gpu


def _split_rows(qp, H, P, off_col, logit_col):
    B, Nq, _ = qp.shape
    off = qp[..., off_col:off_col + 2 * H * P].reshape(B, Nq, H, 1, P, 2)
    w = qp[..., logit_col:logit_col + H * P].reshape(B, Nq, H, P).softmax(-1).reshape(B, Nq, H, 1, P)
    return off, w


def _grid(bev_h, bev_w):
    xs = (torch.arange(bev_w, dtype=torch.float32) + 0.5) / bev_w
    ys = (torch.arange(bev_h, dtype=torch.float32) + 0.5) / bev_h
    return torch.stack((xs.repeat(bev_h), ys.repeat_interleave(bev_w)), -1)          # (Nq, 2)


def _close(got, want, what):
    scale = float(want.abs().max()) + 1e-6
    err = float((got - want).abs().max())
    assert err <= 2e-4 * scale + 2e-6, (what, err, scale)


@pytest.mark.parametrize('B,H,Dh,P,bev,fhw,pad', [
    (2, 8, 32, 8, (13, 11), (9, 10), 0),       # LiDAR cross-attention geometry (P = 8, map != grid)
    (1, 8, 32, 4, (12, 12), (12, 12), 0),      # BEV self-attention (value map = the grid)
    (2, 4, 16, 4, (7, 9), (5, 6), 8),          # rows with foreign columns around the offset / logit blocks
    (1, 2, 8, 3, (5, 5), (4, 7), 0),           # odd point count
    (1, 1, 64, 16, (4, 6), (6, 4), 0),
    (2, 3, 8, 2, (6, 5), (3, 3), 4),
])
def test_bev_sample_bwd_vs_oracle_autograd(B, H, Dh, P, bev, fhw, pad):
    from unibev_b200 import ops
    g = torch.Generator().manual_seed(H * 100 + P)
    (bev_h, bev_w), (fH, fW) = bev, fhw
    Nq, Nv, C = bev_h * bev_w, fH * fW, H * Dh
    ld = 3 * H * P + 2 * pad
    off_col, logit_col = pad, pad + 2 * H * P + pad
    value = torch.randn(B, Nv, C, generator=g)
    qp = torch.randn(B, Nq, ld, generator=g) * 2.5             # offsets of a few pixels: some samples leave the map
    go = torch.randn(B, Nq, C, generator=g)

    v_c, q_c = value.clone().requires_grad_(), qp.clone().requires_grad_()
    off, w = _split_rows(q_c, H, P, off_col, logit_col)
    loc = _grid(bev_h, bev_w)[None, :, None, None, None, :] + off / torch.tensor([fW, fH], dtype=torch.float32)
    want = ms.msda_core(v_c.view(B, Nv, H, Dh), [(fH, fW)], loc, w)
    want.backward(go)

    v_g, q_g = value.cuda().requires_grad_(), qp.cuda().requires_grad_()
    got = ops.BevSampleFunction.apply(v_g, q_g, bev_h, bev_w, fH, fW, H, P, off_col, logit_col)
    _close(got.detach().cpu(), want.detach(), 'out')
    got.backward(go.cuda())
    _close(v_g.grad.cpu(), v_c.grad, 'grad_value')
    _close(q_g.grad.cpu(), q_c.grad, 'grad_qproj')
    if pad:        # columns that are neither offsets nor logits carry no gradient
        cols = torch.ones(ld, dtype=torch.bool)
        cols[off_col:off_col + 2 * H * P] = False
        cols[logit_col:logit_col + H * P] = False
        assert float(q_g.grad[..., cols.cuda()].abs().max()) == 0.0


@pytest.mark.parametrize('B,N,H,Dh,P,D,bev,fhw', [
    (2, 6, 8, 32, 8, 4, (10, 9), (5, 8)),      # the nuScenes layout: six cameras, 4 anchors, 8 points
    (3, 3, 4, 16, 4, 2, (6, 7), (4, 5)),
    (1, 2, 2, 8, 6, 3, (5, 4), (3, 6)),
])
def test_img_sample_bwd_vs_oracle_autograd(B, N, H, Dh, P, D, bev, fhw):
    from unibev_b200 import ops
    g = torch.Generator().manual_seed(N * 10 + P)
    (bev_h, bev_w), (fH, fW) = bev, fhw
    Nq, Nv, C = bev_h * bev_w, fH * fW, H * Dh
    value = torch.randn(B, N, Nv, C, generator=g)
    qp = torch.randn(B, Nq, 3 * H * P, generator=g) * 1.5
    ref = torch.rand(B, Nq, N, D, 2, generator=g) * 1.2 - 0.1
    # visibility differs between batch items: item 0 decides WHICH cameras contribute, each item's own mask the divisor;
    # some queries are seen by no camera, some by several
    mask = (torch.rand(B, Nq, N, generator=g) < 0.35).to(torch.uint8) * 5
    go = torch.randn(B, Nq, C, generator=g)

    v_c, q_c = value.clone().requires_grad_(), qp.clone().requires_grad_()
    off, w = _split_rows(q_c, H, P, 0, 2 * H * P)
    off = off / torch.tensor([fW, fH], dtype=torch.float32)
    hit0 = (mask[0] != 0).float()                                            # (Nq, N)
    count = (mask != 0).sum(-1).clamp(min=1).float()                         # (B, Nq)
    want = torch.zeros(B, Nq, C)
    for n in range(N):
        anchors = ref[:, :, n][:, :, None, None, torch.arange(P) % D, :]     # point p uses anchor p % D
        out_n = ms.msda_core(v_c[:, n].reshape(B, Nv, H, Dh), [(fH, fW)], anchors + off, w)
        want = want + out_n * hit0[None, :, n, None]
    want = want / count[..., None]
    want.backward(go)

    v_g, q_g = value.cuda().requires_grad_(), qp.cuda().requires_grad_()
    got = ops.ImgSampleFunction.apply(v_g, q_g, ref.cuda(), mask.cuda(), bev_h, bev_w, fH, fW, H, P, 0, 2 * H * P)
    _close(got.detach().cpu(), want.detach(), 'out')
    got.backward(go.cuda())
    _close(v_g.grad.cpu(), v_c.grad, 'grad_value')
    _close(q_g.grad.cpu(), q_c.grad, 'grad_qproj')


def test_fused_sample_rejects_uncovered_shapes():
    from unibev_b200 import ops
    v = torch.randn(1, 9, 2 * 12, device='cuda').requires_grad_()           # head size 12: no fused kernel
    q = torch.randn(1, 4, 3 * 2 * 2, device='cuda').requires_grad_()
    assert not ops.fused_sample_supported(12, 2) and ops.fused_sample_supported(32, 8)
    with pytest.raises(ValueError):
        ops.BevSampleFunction.apply(v, q, 2, 2, 3, 3, 2, 2, 0, 8)
