"""GPU parity of the window-staged (TMA + fp16 planes) sampling kernels.

Three references: (1) the CPU oracle's fp32 MSDA core (oracle/mmcv_semantics.py) -- stated tolerance: the
storage rounding of these kernels (fp16 values, fp16 combined weights, fp32 accumulation) bounds the error by
2 * 2^-11 * sum_k w_k |v_k| <= 1e-3 * max|v|; (2) a quantisation-aware emulation of exactly that storage
(tests/helpers.msda_quantised) -- tight on average; (3) the fp32 kernels of the same library, which are
themselves pinned to the reference's golden vectors in test_gpu_kernels.py."""
import numpy as np
import pytest
import torch

from oracle import mmcv_semantics as ms
from tests.helpers import bev_loc_weights, msda_quantised

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from unibev_b200 import ops as _ops
    return _ops


def test_value_to_half_layout(ops):
    g = torch.Generator().manual_seed(0)
    G, Nv, H, Dh = 3, 37, 8, 32
    v = torch.randn(G * Nv, H * Dh, generator=g) * 5
    v[0, 0], v[1, 1] = 1e6, -1e6                                   # saturates instead of overflowing to inf
    got = ops.value_to_half(v.cuda(), G, Nv, H).cpu()
    want = v.clamp(-65504, 65504).view(G, Nv, H, Dh).permute(0, 2, 1, 3).half()
    assert got.shape == (G, H, Nv, Dh)
    assert torch.equal(got, want)


def _bev_case(ops, B, bev_h, bev_w, fH, fW, H, P, off_scale, seed, halo=0):
    from unibev_b200 import _cabi
    g = torch.Generator().manual_seed(seed)
    Nq, C = bev_h * bev_w, H * 32
    value = torch.randn(B, fH * fW, C, generator=g)
    qproj = torch.cat((torch.randn(B, Nq, H * P * 2, generator=g) * off_scale,
                       torch.randn(B, Nq, H * P, generator=g)), -1)
    loc, aw = bev_loc_weights(qproj, bev_h, bev_w, fH, fW, H, P)
    want = ms.msda_core(value.view(B, fH * fW, H, 32), [(fH, fW)], loc.unsqueeze(3), aw.unsqueeze(3))
    emul = msda_quantised(value.view(B, fH * fW, H, 32), (fH, fW), loc, aw)
    vg, qg = value.cuda(), qproj.cuda()
    _cabi.check(_cabi.lib().ub_set_window_halo(halo), 'ub_set_window_halo')
    try:
        v16 = ops.value_to_half(vg.view(B * fH * fW, C), B, fH * fW, H)
        ws = torch.zeros(2, dtype=torch.int32).cuda()      # caller-owned work counters, left zero by every call
        got = ops.bev_sample_win(v16, qg, bev_h, bev_w, fH, fW, H, P, 0, H * P * 2, workspace=ws).cpu()
        got2 = ops.bev_sample_win(v16, qg, bev_h, bev_w, fH, fW, H, P, 0, H * P * 2, workspace=ws).cpu()
        assert ws.cpu().tolist() == [0, 0]
        got16 = ops.bev_sample_win(v16, qg, bev_h, bev_w, fH, fW, H, P, 0, H * P * 2, out_dtype=torch.float16).cpu()
    finally:
        _cabi.lib().ub_set_window_halo(0)
    fp32 = ops.bev_sample(vg, qg, bev_h, bev_w, fH, fW, H, P, 0, H * P * 2).cpu()
    torch.testing.assert_close(fp32, want, rtol=1e-4, atol=5e-5)
    vmax = float(value.abs().max())
    err, err_q = (got - want).abs(), (got - emul).abs()
    assert float(err.max()) <= 1e-3 * vmax, (float(err.max()), vmax)
    assert float(err.mean()) <= 1.5e-4, float(err.mean())
    assert float(err_q.max()) <= 1e-3 * vmax and float(err_q.mean()) <= 2e-5, (float(err_q.max()), float(err_q.mean()))
    torch.testing.assert_close(got2, got, rtol=0, atol=1e-6)
    # fp16 output rows: the fp32 result rounded once (rows merged from the slow path by fp16 red.add: a few more ulps)
    assert got16.dtype == torch.float16
    torch.testing.assert_close(got16.float(), got, rtol=2e-3 if off_scale > 4 or halo else 1e-3, atol=2e-3 * vmax / 4)


@pytest.mark.parametrize('B,bev,f,H,P,off_scale', [
    (1, (32, 32), (29, 29), 8, 8, 2.0),      # LiDAR-like down-scaled map, offsets inside the halo
    (2, (10, 12), (9, 9), 8, 8, 1.5),        # tiny map, windows hang over every border
    (1, (33, 17), (40, 25), 8, 8, 2.0),      # partial tiles, up-scaled value map
    (1, (48, 48), (48, 48), 8, 4, 1.5),      # BEV self-attention shape (P = 4)
    (2, (20, 20), (20, 20), 4, 4, 1.0),
    (1, (32, 32), (29, 29), 8, 8, 12.0),     # many samples beyond the halo: slow path + red.add merge
    (1, (16, 16), (15, 15), 2, 8, 40.0),     # most samples off the map or far
])
def test_bev_sample_win_vs_oracle(ops, B, bev, f, H, P, off_scale):
    _bev_case(ops, B, bev[0], bev[1], f[0], f[1], H, P, off_scale, seed=B * 100 + P)


def test_bev_sample_win_small_halo_forces_slow_path(ops):
    _bev_case(ops, 1, 32, 32, 29, 29, 8, 8, 4.0, seed=7, halo=1)


def test_bev_sample_win_full_size(ops):
    """BASELINE.json sizes (200 x 200 BEV, 180 x 180 LiDAR map / 200 x 200 self-attention), against the fp32
    kernel of the same library (pinned to the oracle above and in test_gpu_kernels.py)."""
    for fH, P, off in ((180, 8, 3.0), (200, 4, 1.5)):
        g = torch.Generator().manual_seed(P)
        B, H, Nq = 1, 8, 200 * 200
        value = torch.randn(B, fH * fH, 256, generator=g).cuda()
        qproj = torch.cat((torch.randn(B, Nq, H * P * 2, generator=g) * off, torch.randn(B, Nq, H * P, generator=g)),
                          -1).cuda()
        fp32 = ops.bev_sample(value, qproj, 200, 200, fH, fH, H, P, 0, H * P * 2)
        v16 = ops.value_to_half(value.view(-1, 256), B, fH * fH, H)
        got = ops.bev_sample_win(v16, qproj, 200, 200, fH, fH, H, P, 0, H * P * 2)
        err = (got - fp32).abs()
        assert float(err.max()) <= 1e-3 * float(value.abs().max())
        assert float(err.mean()) <= 1.5e-4


def test_bev_sample_win_rejects_uncovered_shapes(ops):
    from unibev_b200 import _cabi
    v16 = torch.zeros(1, 4, 81, 16, dtype=torch.float16, device='cuda')
    qp = torch.zeros(1, 100, 4 * 8 * 3, device='cuda')
    with pytest.raises(_cabi.UnsupportedShape):
        ops.bev_sample_win(v16, qp, 10, 10, 9, 9, 4, 8, 0, 64)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.bev_sample_win(v16.cpu(), qp, 10, 10, 9, 9, 4, 8, 0, 64)


def _camera_inputs(B, bev_h, bev_w, fh, fw, H, P, seed, perturb=True):
    from unibev_b200 import synth
    g = torch.Generator().manual_seed(seed)
    N, C, D = 6, H * 32, 4
    rig = synth.nominal_rig((928, 1600), N).astype(np.float32)
    l2i = torch.from_numpy(np.stack([rig] * B))
    if perturb and B > 1:
        l2i[1, :, 0, 3] += 40.0                                      # item 1 sees a shifted scene (hit-list quirk)
    value = torch.randn(B, N, fh * fw, C, generator=g)
    qproj = torch.cat((torch.randn(B, bev_h * bev_w, H * P * 2, generator=g) * 2.0,
                       torch.randn(B, bev_h * bev_w, H * P, generator=g)), -1)
    zs = [(i + 0.5) * 2 / 8 for i in range(D)]
    return l2i, value, qproj, zs


def want_inv_of(m):
    return 1.0 / (m != 0).sum(-1).clamp(min=1).float()


def test_build_hits_exact(ops):
    l2i, _, _, zs = _camera_inputs(2, 40, 36, 29, 50, 8, 8, 3)
    _, mask = ops.project_points(l2i.cuda(), zs, [-54, -54, -5, 54, 54, 3], 928, 1600, 40, 36)
    hit_idx, hit_cnt, inv_cnt, hit_ic = (t.cpu() for t in ops.build_hits(mask))
    m = mask.cpu()
    Nq = m.shape[1]
    seen_before = torch.zeros(Nq, dtype=torch.bool)
    for n in range(6):
        hit = m[0, :, n] != 0
        n_first, n_later = int(hit_cnt[n]), int(hit_cnt[6 + n])
        # first hits (no lower camera sees the query) ascending from the front, later hits from the back
        assert torch.equal(hit_idx[n, :n_first], (hit & ~seen_before).nonzero().squeeze(1).int())
        assert torch.equal(hit_idx[n, Nq - n_later:].flip(0), (hit & seen_before).nonzero().squeeze(1).int())
        # together: exactly the reference's per-camera list (spatial_cross_attention_img.py:141-152)
        both = torch.cat((hit_idx[n, :n_first], hit_idx[n, Nq - n_later:])).sort().values
        assert torch.equal(both, hit.nonzero().squeeze(1).int())
        # 1 / count in list order, for both batch items
        for b in range(m.shape[0]):
            assert torch.equal(hit_ic[b, n, :n_first], want_inv_of(m)[b, hit_idx[n, :n_first].long()])
            assert torch.equal(hit_ic[b, n, Nq - n_later:], want_inv_of(m)[b, hit_idx[n, Nq - n_later:].long()])
        seen_before |= hit
    assert torch.equal(hit_idx[6, :int(hit_cnt[12])], (~seen_before).nonzero().squeeze(1).int())
    assert int(hit_cnt[6:12].sum()) > 0 and int(hit_cnt[12]) > 0      # the case has multi-camera and unseen queries
    want_inv = 1.0 / (m != 0).sum(-1).clamp(min=1).float()
    assert torch.equal(inv_cnt, want_inv)
    assert int(hit_cnt[:6].sum()) > 0


@pytest.mark.parametrize('B,bev,fhw,P', [(1, (40, 36), (29, 50), 8), (2, (50, 50), (29, 50), 8), (2, (24, 24), (8, 22), 4)])
def test_img_sample_win_vs_fp32_kernel(ops, B, bev, fhw, P):
    H = 8
    bev_h, bev_w = bev
    fh, fw = fhw
    l2i, value, qproj, zs = _camera_inputs(B, bev_h, bev_w, fh, fw, H, P, seed=B + P)
    ref_cam, mask = ops.project_points(l2i.cuda(), zs, [-54, -54, -5, 54, 54, 3], 928, 1600, bev_h, bev_w)
    vg, qg = value.cuda(), qproj.cuda()
    fp32 = ops.img_sample(vg, qg, ref_cam, mask, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2).cpu()
    hits = ops.build_hits(mask)
    v16 = ops.value_to_half(vg.view(-1, H * 32), B * 6, fh * fw, H).view(B, 6, H, fh * fw, 32)
    got = ops.img_sample_win(v16, qg, ref_cam, hits, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2).cpu()
    got16 = ops.img_sample_win(v16, qg, ref_cam, hits, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2,
                               out_dtype=torch.float16).cpu()
    assert got16.dtype == torch.float16
    torch.testing.assert_close(got16.float(), got, rtol=2e-3, atol=5e-4 * float(value.abs().max()))
    # every row is written by the call itself (no zero-fill pass): poison the output buffer first
    poisoned = torch.full_like(got, float('nan')).cuda()
    ops.img_sample_win(v16, qg, ref_cam, hits, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2, out=poisoned)
    torch.testing.assert_close(poisoned.cpu(), got, rtol=0, atol=1e-6)
    if bev == (40, 36):
        assert int((got.abs().sum(-1) == 0).sum()) > 0              # this case has rows no camera sees
    assert float(fp32.abs().max()) > 0.1                           # the rig does see the grid
    err = (got - fp32).abs()
    assert float(err.max()) <= 1e-3 * float(value.abs().max()), float(err.max())
    assert float(err.mean()) <= 1.5e-4, float(err.mean())
    # quantisation-aware check through the oracle's semantics for camera 0 .. N-1 of item 0 is covered by the
    # encoder-level tests (test_gpu_encoder.py::test_full_size_vs_oracle_tf32)


def test_bev_sample_win_concurrent_graph_replays(ops):
    """no shared work counters -- two CUDA graphs of the same call, each with its own workspace,
    replayed at the same time on two streams give the same result as the serial call."""
    g = torch.Generator().manual_seed(9)
    B, bev, f, H, P = 2, 200, 180, 8, 8
    v16 = ops.value_to_half(torch.randn(B * f * f, H * 32, generator=g).cuda(), B, f * f, H)
    qp = torch.cat((torch.randn(B, bev * bev, H * P * 2, generator=g) * 3, torch.randn(B, bev * bev, H * P, generator=g)),
                   -1).cuda()
    want = ops.bev_sample_win(v16, qp, bev, bev, f, f, H, P, 0, H * P * 2)
    streams = [torch.cuda.Stream() for _ in range(2)]
    graphs, outs = [], []
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ops.bev_sample_win(v16, qp, bev, bev, f, f, H, P, 0, H * P * 2)
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                outs.append(ops.bev_sample_win(v16, qp, bev, bev, f, f, H, P, 0, H * P * 2))
        graphs.append(gr)
    torch.cuda.synchronize()
    for _ in range(5):
        for o in outs:
            o.zero_()
        torch.cuda.synchronize()
        for s, gr in zip(streams, graphs):
            with torch.cuda.stream(s):
                gr.replay()
        torch.cuda.synchronize()
        for o in outs:
            torch.testing.assert_close(o, want, rtol=0, atol=1e-6)
