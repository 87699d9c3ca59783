#!/usr/bin/env python
"""Time the three sampling-kernel flavours at LC-CNW-256 size over tile / CTA-shape settings (GPU box only).
Prints one line per configuration: avg microseconds over `iters` launches, CUDA events, L2-cold-ish inputs
(value maps rotate over 4 copies)."""
import itertools
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unibev_b200 import _cabi, ops, synth
from unibev_b200.plugin.encoder import anchor_heights


def timeit(fn, iters=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


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
    metas = synth.img_metas(B)
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas], dtype=np.float32)).to(dev)
    ref_cam, mask = ops.project_points(l2i, anchor_heights(8, 4).tolist(), synth.PC_RANGE, 928, 1600, 200, 200)
    out = torch.empty(B, Nq, C, device=dev)
    runs = {
        'bev_self': (0, lambda i: ops.bev_sample(v_self[i % R], qp_self, 200, 200, 200, 200, H, 4, 0, 64, out=out), 97.3e6),
        'pts_cross': (0, lambda i: ops.bev_sample(v_pts[i % R], qp_pts, 200, 200, 180, 180, H, 8, 0, 128, out=out), 104.9e6),
        'img_cross': (1, lambda i: ops.img_sample(v_img[i % R], qp_img, ref_cam, mask, 200, 200, 29, 50, H, 8, 0, 128, out=out), 82.5e6),
    }
    results = []
    for name, (which, fn, nbytes) in runs.items():
        for tw, cps in [(8, 0), (8, 2), (8, 3), (8, 5), (8, 6), (4, 0), (16, 0), (32, 0)]:
            _cabi.set_tuning(which, tw, cps)
            us = timeit(fn)
            results.append((name, tw, 64 // tw, cps, us, nbytes / us / 1e3))
            print('%-10s tile %2dx%-2d ctas/SM %d : %7.1f us  (%.0f GB/s algorithmic)' % results[-1], flush=True)
    _cabi.set_tuning(0)
    _cabi.set_tuning(1)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(results, open('gpurun_out/sweep_sampling.json', 'w'))


if __name__ == '__main__':
    main()
