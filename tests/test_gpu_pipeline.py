"""The streaming / graph-replay front ends the end-to-end figure of bench.py is measured through
(unibev_b200/pipeline.py): results must equal the plain ``UniBEVTransformer.encode`` call on the same frame."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BEV = 48


def _setup(batch=2, n_frames=5):
    from unibev_b200 import synth
    model, _ = synth.build_model('unibev_nus_LC_cnw_256', num_layers=2)
    model = model.cuda().eval()
    frames = [synth.make_inputs('unibev_nus_LC_cnw_256', batch=batch, bev_hw=(BEV, BEV), seed=10 + i, pin=True)
              for i in range(n_frames)]
    for i, f in enumerate(frames):                     # a different rig per frame: hit lists must follow the frame
        for m in f['img_metas']:
            m['lidar2img'] = [np.asarray(a) @ np.diag([1.0, 1.0, 1.0, 1.0 + 0.02 * i]) for a in m['lidar2img']]
    bev_q = frames[0]['bev_queries'].cuda()
    bev_pos = frames[0]['bev_pos'].cuda()
    return model, frames, bev_q, bev_pos


def _eager(model, f, bev_q, bev_pos):
    with torch.no_grad():
        return model.encode([f['img_feats'][0].cuda()], [f['pts_feats'][0].cuda()], bev_q, BEV, BEV, bev_pos=bev_pos,
                            img_metas=f['img_metas'])


@pytest.mark.parametrize('graphs', [False, True])
def test_frame_pipeline_matches_plain_encode(graphs):
    from unibev_b200.pipeline import FramePipeline
    model, frames, bev_q, bev_pos = _setup()
    f0 = frames[0]
    pipe = FramePipeline(model, bev_q, BEV, BEV, bev_pos=bev_pos, img_shape=tuple(f0['img_feats'][0].shape),
                         pts_shape=tuple(f0['pts_feats'][0].shape), img_hw=tuple(f0['img_metas'][0]['img_shape'][0][:2]),
                         depth=3, graphs=graphs)
    got = {}
    for i, f in enumerate(frames + frames[:2]):          # 7 submits over 3 slots: every slot (and graph) is reused
        t = pipe.submit(f['img_feats'][0], f['pts_feats'][0], f['img_metas'])
        assert t == i
        if t >= 2:
            got[t - 2] = pipe.result(t - 2).clone()
    for t in range(pipe.n_submitted - 2, pipe.n_submitted):
        got[t] = pipe.result(t).clone()
    pipe.drain()
    with pytest.raises(ValueError):
        pipe.result(0)                                   # slot long since reused
    assert pipe.h2d_bytes == sum(t.numel() * 4 for t in (f0['img_feats'][0], f0['pts_feats'][0])) + 2 * 6 * 16 * 4
    assert pipe.d2h_bytes == 2 * BEV * BEV * 256 * 4
    for t, out in got.items():
        want = _eager(model, (frames + frames[:2])[t], bev_q, bev_pos).cpu()
        assert float(want.abs().max()) > 0.1
        torch.testing.assert_close(out, want, rtol=0, atol=1e-5)     # same kernels, same inputs


def test_graphed_encoder_follows_rewritten_input_buffers():
    from unibev_b200.pipeline import GraphedEncoder
    model, frames, bev_q, bev_pos = _setup(batch=1, n_frames=2)
    f0, f1 = frames
    img, pts = f0['img_feats'][0].cuda(), f0['pts_feats'][0].cuda()
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in f0['img_metas']], dtype=np.float32)).cuda()
    hw = tuple(f0['img_metas'][0]['img_shape'][0][:2])
    g = GraphedEncoder(model, img, pts, bev_q, BEV, BEV, bev_pos=bev_pos[:1], lidar2img=l2i, img_hw=hw)
    torch.testing.assert_close(g.replay().cpu(), _eager(model, f0, bev_q, bev_pos[:1]).cpu(), rtol=0, atol=1e-5)
    img.copy_(f1['img_feats'][0])
    pts.copy_(f1['pts_feats'][0])
    l2i.copy_(torch.from_numpy(np.asarray([m['lidar2img'] for m in f1['img_metas']], dtype=np.float32)))
    torch.testing.assert_close(g.replay().cpu(), _eager(model, f1, bev_q, bev_pos[:1]).cpu(), rtol=0, atol=1e-5)


def test_frame_pipeline_half_result_option():
    """Opt-in fp16 result (half the device->host bytes): the fp32 result rounded once."""
    from unibev_b200.pipeline import FramePipeline
    model, frames, bev_q, bev_pos = _setup(batch=1, n_frames=2)
    f0 = frames[0]
    pipe = FramePipeline(model, bev_q, BEV, BEV, bev_pos=bev_pos[:1], img_shape=tuple(f0['img_feats'][0].shape),
                         pts_shape=tuple(f0['pts_feats'][0].shape), img_hw=tuple(f0['img_metas'][0]['img_shape'][0][:2]),
                         depth=2, graphs=True, result_dtype=torch.float16)
    for f in frames:
        t = pipe.submit(f['img_feats'][0], f['pts_feats'][0], f['img_metas'])
    out = pipe.result(t).clone()
    pipe.drain()
    assert out.dtype == torch.float16 and pipe.d2h_bytes == BEV * BEV * 256 * 2
    want = _eager(model, frames[1], bev_q, bev_pos[:1]).cpu()
    torch.testing.assert_close(out.float(), want.half().float(), rtol=0, atol=1e-3)
