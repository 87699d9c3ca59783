#!/usr/bin/env python
"""Time the window-staged sampling kernels against the fp32 ones at LC-CNW-256 size (GPU box only).
CUDA events around `iters` launches, inputs rotate over 4 copies and a 256 MB buffer is rewritten between
launches (L2 flush).  Prints avg microseconds and algorithmic GB/s (SURVEY.md section 8d byte counts)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unibev_b200 import _cabi, ops, synth
from unibev_b200.plugin.encoder import anchor_heights

flush_buf = None


def timeit(fn, iters=20, flush=True):
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(iters):
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot * 1e3 / iters


def main():
    dev = torch.device('cuda')
    torch.manual_seed(0)
    B, Nq, C, H = 1, 40000, 256, 8
    model, _ = synth.build_model('unibev_nus_LC_cnw_256', num_layers=1)
    model = model.to(dev)
    x = torch.randn(B, Nq, C, device=dev)
    li, lp = model.img_bev_encoder.layers[0], model.pts_bev_encoder.layers[0]

    def qproj(att):
        return torch.cat((att.sampling_offsets(x), att.attention_weights(x)), -1).contiguous()
    with torch.no_grad():
        qp_self, qp_pts, qp_img = qproj(lp.attentions[0]), qproj(lp.attentions[1].deformable_attention), \
            qproj(li.attentions[1].deformable_attention)
    R = 4
    v_self = [torch.randn(B, Nq, C, device=dev) for _ in range(R)]
    v_pts = [torch.randn(B, 180 * 180, C, device=dev) for _ in range(R)]
    v_img = [torch.randn(B, 6, 1450, C, device=dev) for _ in range(R)]
    h_self = [ops.value_to_half(v.view(-1, C), B, Nq, H) for v in v_self]
    h_pts = [ops.value_to_half(v.view(-1, C), B, 180 * 180, H) for v in v_pts]
    h_img = [ops.value_to_half(v.view(-1, C), B * 6, 1450, H).view(B, 6, H, 1450, 32) for v in v_img]
    metas = synth.img_metas(B)
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas], dtype=np.float32)).to(dev)
    ref_cam, mask = ops.project_points(l2i, anchor_heights(8, 4).tolist(), synth.PC_RANGE, 928, 1600, 200, 200)
    hits = ops.build_hits(mask)
    out = torch.empty(B, Nq, C, device=dev)
    res = {}

    def rec(name, us, nbytes):
        res[name] = dict(us=us, gbs=nbytes / us / 1e3)
        print('%-34s %8.1f us   %7.0f GB/s algorithmic' % (name, us, nbytes / us / 1e3), flush=True)

    rec('bev_self fp32', timeit(lambda i: ops.bev_sample(v_self[i % R], qp_self, 200, 200, 200, 200, H, 4, 0, 64, out=out)), 97.3e6)
    rec('pts_cross fp32', timeit(lambda i: ops.bev_sample(v_pts[i % R], qp_pts, 200, 200, 180, 180, H, 8, 0, 128, out=out)), 104.9e6)
    rec('img_cross fp32', timeit(lambda i: ops.img_sample(v_img[i % R], qp_img, ref_cam, mask, 200, 200, 29, 50, H, 8, 0, 128, out=out)), 82.5e6)
    tmp = torch.empty_like(h_self[0])
    rec('value_to_half self (41->20 MB)', timeit(lambda i: ops.value_to_half(v_self[i % R].view(-1, C), B, Nq, H, out=tmp)), 61.4e6)
    rec('build_hits', timeit(lambda i: ops.build_hits(mask)), 1.0)
    for halo in (0, 6, 12):
        _cabi.check(_cabi.lib().ub_set_window_halo(halo), 'halo')
        tag = 'halo=%s' % (halo or 'default')
        rec('bev_self win ' + tag, timeit(lambda i: ops.bev_sample_win(h_self[i % R], qp_self, 200, 200, 200, 200, H, 4, 0, 64, out=out)), 97.3e6)
        rec('pts_cross win ' + tag, timeit(lambda i: ops.bev_sample_win(h_pts[i % R], qp_pts, 200, 200, 180, 180, H, 8, 0, 128, out=out)), 104.9e6)
    _cabi.lib().ub_set_window_halo(0)
    rec('img_cross win', timeit(lambda i: ops.img_sample_win(h_img[i % R], qp_img, ref_cam, hits, 200, 200, 29, 50, H, 8, 0, 128, out=out)), 82.5e6)
    rec('pts_cross win (L2 warm)', timeit(lambda i: ops.bev_sample_win(h_pts[0], qp_pts, 200, 200, 180, 180, H, 8, 0, 128, out=out), flush=False), 104.9e6)
    # far-sample statistics of the synthetic weights (how many samples leave the default halo)
    off = qp_pts[..., :128].view(B, Nq, H, 8, 2).abs().amax(-1)
    print('pts offsets: max |off| = %.2f px, share beyond 9 px = %.4f' % (float(off.max()), float((off > 9).float().mean())))
    off = qp_self[..., :64].view(B, Nq, H, 4, 2).abs().amax(-1)
    print('self offsets: max |off| = %.2f px, share beyond 5 px = %.4f' % (float(off.max()), float((off > 5).float().mean())))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/bench_window.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
