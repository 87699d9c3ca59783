"""fp32 window-staged sampling kernels (win_sample32.cu) against the oracle's MSDA core: fp32 tolerance (the same
rtol 1e-4 / atol 5e-5 as the fp32 tile kernels), slow path, ragged tiles, full size, caller-owned work counters."""
import pytest
import torch

from oracle import mmcv_semantics as ms
from tests.helpers import bev_loc_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from unibev_b200 import ops as _ops
    return _ops


def _case(ops, B, bev_h, bev_w, fH, fW, H, P, off_scale, seed, halo=0):
    from unibev_b200 import _cabi
    g = torch.Generator().manual_seed(seed)
    Nq, C = bev_h * bev_w, H * 32
    value = torch.randn(B, fH * fW, C, generator=g)
    qproj = torch.cat((torch.randn(B, Nq, H * P * 2, generator=g) * off_scale,
                       torch.randn(B, Nq, H * P, generator=g)), -1)
    loc, aw = bev_loc_weights(qproj, bev_h, bev_w, fH, fW, H, P)
    want = ms.msda_core(value.view(B, fH * fW, H, 32), [(fH, fW)], loc.unsqueeze(3), aw.unsqueeze(3))
    planes = ops.value_to_planes32(value.cuda().view(B * fH * fW, C), B, fH * fW, H)
    assert planes.shape == (B, 2 * H, fH * fW, 16)
    torch.testing.assert_close(ops.planes32_to_rows(planes).cpu(), value.view(-1, C), rtol=0, atol=0)
    _cabi.check(_cabi.lib().ub_set_window_halo(halo), 'ub_set_window_halo')
    try:
        ws = torch.zeros(2, dtype=torch.int32).cuda()
        got = ops.bev_sample_win32(planes, qproj.cuda(), bev_h, bev_w, fH, fW, H, P, 0, H * P * 2, workspace=ws).cpu()
        got2 = ops.bev_sample_win32(planes, qproj.cuda(), bev_h, bev_w, fH, fW, H, P, 0, H * P * 2, workspace=ws).cpu()
        assert ws.cpu().tolist() == [0, 0]
    finally:
        _cabi.lib().ub_set_window_halo(0)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=5e-5)
    torch.testing.assert_close(got2, got, rtol=0, atol=1e-6)       # (far samples accumulate with red.add)
    return got


@pytest.mark.parametrize('B,bev,f,H,P,off_scale', [
    (1, (32, 32), (32, 32), 8, 4, 1.5),        # self-attention geometry, offsets inside the halo
    (2, (40, 24), (36, 20), 8, 8, 3.0),        # LiDAR-cross geometry (scale != 1), ragged tiles
    (1, (16, 16), (12, 12), 2, 8, 40.0),       # huge offsets: most samples far or outside the map
    (1, (50, 50), (45, 45), 4, 4, 6.0),        # offsets beyond the halo: slow path mixed with the window path
    (3, (17, 33), (9, 29), 1, 8, 2.0),         # one head, odd sizes
])
def test_bev_sample_win32_vs_oracle(ops, B, bev, f, H, P, off_scale):
    _case(ops, B, bev[0], bev[1], f[0], f[1], H, P, off_scale, seed=B * 100 + P)


def test_bev_sample_win32_small_halo_forces_slow_path(ops):
    _case(ops, 1, 32, 32, 32, 32, 8, 8, 4.0, seed=5, halo=1)


def test_bev_sample_win32_matches_tile_kernel_at_full_size(ops):
    """BASELINE sizes (200 x 200 queries, 180 x 180 LiDAR map, 8 heads x 8 points, and the P = 4 self-attention): equal to
    the fp32 tile kernel (itself pinned to the oracle) to fp32 rounding."""
    g = torch.Generator().manual_seed(0)
    for fH, P, B in ((180, 8, 2), (200, 4, 1)):
        H, C = 8, 256
        value = torch.randn(B, fH * fH, C, generator=g).cuda()
        qproj = torch.cat((torch.randn(B, 40000, H * P * 2, generator=g) * 2.5, torch.randn(B, 40000, H * P, generator=g)), -1).cuda()
        planes = ops.value_to_planes32(value.view(B * fH * fH, C), B, fH * fH, H)
        got = ops.bev_sample_win32(planes, qproj, 200, 200, fH, fH, H, P, 0, H * P * 2)
        ref = ops.bev_sample(value, qproj, 200, 200, fH, fH, H, P, 0, H * P * 2)
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=2e-6)


def test_bev_sample_win32_rejects_uncovered_shapes(ops):
    from unibev_b200 import _cabi
    planes = torch.zeros(1, 8, 81, 16).cuda()
    qp = torch.zeros(1, 100, 4 * 2 * 3).cuda()
    with pytest.raises(_cabi.UnsupportedShape):
        ops.bev_sample_win32(planes, qp, 10, 10, 9, 9, 4, 2, 0, 16)       # 2 points: not covered
    with pytest.raises(RuntimeError):
        ops.bev_sample_win32(planes.cpu(), qp, 10, 10, 9, 9, 4, 8, 0, 64)


# ---------------------------------------------------------------------------------------------------------------------
# camera mode: hit-list-ordered inputs (ub_hit_order), the row-scattering offset|logit projection and the fp32 window kernel
def _camera(ops, B, bev, fhw, P, seed):
    from tests.test_gpu_window import _camera_inputs
    H = 8
    l2i, value, qproj, zs = _camera_inputs(B, bev[0], bev[1], fhw[0], fhw[1], H, P, seed=seed)
    ref_cam, mask = ops.project_points(l2i.cuda(), zs, [-54, -54, -5, 54, 54, 3], 928, 1600, bev[0], bev[1])
    return H, value, qproj, ref_cam, mask


def test_hit_order_exact(ops):
    H, value, qproj, ref_cam, mask = _camera(ops, 2, (40, 36), (29, 50), 8, 3)
    hits = ops.build_hits(mask)
    q_dst, hit_ref, hit_meta = (t.cpu() for t in ops.hit_order(mask, ref_cam, hits))
    hit_idx, hit_cnt, inv_cnt = hits[0].cpu(), hits[1].cpu(), hits[2].cpu()
    m, rc = mask.cpu(), ref_cam.cpu()
    B, Nq, N = m.shape
    n_pairs = 0
    for q in range(Nq):
        cams = (m[0, q] != 0).nonzero().squeeze(1).tolist()          # batch item 0's visibility (reference quirk)
        row = q_dst[q].tolist()
        assert row[len(cams):] == [-1] * (N - len(cams))
        for j, n in enumerate(cams):
            d = row[j]
            assert d // Nq == n
            pos = d % Nq
            assert int(hit_idx[n, pos]) == q
            assert pos < int(hit_cnt[n]) or pos >= Nq - int(hit_cnt[N + n])
            for b in range(B):
                assert torch.equal(hit_ref[b, n, pos], rc[b, q, n].reshape(-1))
                assert int(hit_meta[b, n, pos, :1].view(torch.int32)) == q
                assert float(hit_meta[b, n, pos, 1]) == float(inv_cnt[b, q])
            n_pairs += 1
    assert n_pairs == int(hit_cnt[:2 * N].sum())


def test_linear_x3_scatter_rows(ops):
    g = torch.Generator().manual_seed(1)
    B, Nq, N, K, Nout = 2, 1000, 6, 256, 192
    x = torch.randn(B * Nq, K, generator=g)
    w, b = torch.randn(Nout, K, generator=g) / 16, torch.randn(Nout, generator=g)
    q_dst = torch.full((Nq, N), -1, dtype=torch.int32)
    perm = torch.randperm(N * Nq, generator=g)
    k = 0
    for q in range(Nq):                                              # 0 .. 3 distinct destination rows per query
        c = q % 4
        q_dst[q, :c] = perm[k:k + c].int()
        k += c
    out = torch.full((B, N * Nq, Nout), float('nan')).cuda()
    ops.linear_tf32x3_scatter(x.cuda(), ops.split_tf32(w.cuda()), b.cuda(), q_dst.cuda(), Nq, out)
    want = torch.nn.functional.linear(x.double(), w.double(), b.double()).float().view(B, Nq, Nout)
    got = out.cpu()
    written = torch.zeros(N * Nq, dtype=torch.bool)
    for q in range(Nq):
        for d in q_dst[q].tolist():
            if d < 0:
                break
            written[d] = True
            torch.testing.assert_close(got[:, d], want[:, q], rtol=1e-5, atol=5e-5)
    assert bool(torch.isnan(got[:, ~written]).all())                 # no other row is touched


@pytest.mark.parametrize('B,bev,fhw,P', [(1, (40, 36), (29, 50), 8), (2, (50, 50), (29, 50), 8), (2, (24, 24), (8, 22), 4)])
def test_img_sample_win32_vs_fp32_kernel(ops, B, bev, fhw, P):
    """Same result as the fp32 tile kernel (pinned to the oracle / the golden vectors in test_gpu_kernels.py) to fp32
    rounding; every output row is written by the call (poisoned buffer); differing calibrations per batch item."""
    H, value, qproj, ref_cam, mask = _camera(ops, B, bev, fhw, P, seed=B + P)
    bev_h, bev_w = bev
    fh, fw = fhw
    Nq, N = bev_h * bev_w, 6
    vg, qg = value.cuda(), qproj.cuda()
    ref = ops.img_sample(vg, qg, ref_cam, mask, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2).cpu()
    hits = ops.build_hits(mask)
    order = ops.hit_order(mask, ref_cam, hits)
    q_dst = order[0]
    planes = ops.value_to_planes32(vg.view(-1, H * 32), B * N, fh * fw, H)
    # hit-ordered offset|logit rows by torch glue here (the product path: linear_tf32x3_scatter); other rows poisoned
    qp_hit = torch.full((B, N * Nq, qproj.shape[2]), float('nan')).cuda()
    for q_rows, d in ((q_dst[:, j] >= 0, q_dst[:, j]) for j in range(N)):
        qp_hit[:, d[q_rows].long()] = qg[:, q_rows]
    out = torch.full((B, Nq, H * 32), float('nan')).cuda()
    got = ops.img_sample_win32(planes, qp_hit, order, hits, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2, out=out).cpu()
    assert not bool(torch.isnan(got).any())
    if bev == (40, 36):
        assert int((got.abs().sum(-1) == 0).sum()) > 0              # this case has rows no camera sees
    assert float(ref.abs().max()) > 0.1                            # the rig does see the grid
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=5e-6)
