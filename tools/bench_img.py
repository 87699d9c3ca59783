#!/usr/bin/env python
"""Time ub_img_sample_win_fwd (both passes) at LC-CNW-256 size, B frames per launch: CUDA events around batches of
launches over rotating inputs (> L2 in total), so neither host launch gaps nor L2 reuse enter the figure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unibev_b200 import ops, synth
from unibev_b200.plugin.encoder import anchor_heights


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device('cuda')
    torch.manual_seed(0)
    Nq, C, H, R = 40000, 256, 8, 3
    qp = [torch.cat((torch.randn(B, Nq, 128, device=dev) * 2.0, torch.randn(B, Nq, 64, device=dev)), -1) for _ in range(R)]
    h_img = [ops.value_to_half(torch.randn(B * 6 * 1450, C, device=dev), B * 6, 1450, H).view(B, 6, H, 1450, 32) for _ in range(R)]
    metas = synth.img_metas(B)
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas], dtype=np.float32)).to(dev)
    ref_cam, mask = ops.project_points(l2i, anchor_heights(8, 4).tolist(), synth.PC_RANGE, 928, 1600, 200, 200)
    hits = ops.build_hits(mask)
    outs = [torch.empty(B, Nq, C, device=dev, dtype=torch.float16) for _ in range(R)]
    g = torch.cuda.CUDAGraph()
    n = 12
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(3):
            ops.img_sample_win(h_img[i % R], qp[i % R], ref_cam, hits, 200, 200, 29, 50, H, 8, 0, 128, out=outs[i % R])
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for i in range(n):
                ops.img_sample_win(h_img[i % R], qp[i % R], ref_cam, hits, 200, 200, 29, 50, H, 8, 0, 128, out=outs[i % R])
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * n)
    print('img_sample_win B=%d: %.1f us per call (graph of %d calls, %d rotating input sets)' % (B, us, n, R))


if __name__ == '__main__':
    main()
