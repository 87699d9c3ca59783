#!/usr/bin/env python
"""Launch the fused sampling backward kernels at the training shapes of BASELINE configs[4] (for `ncu`): B = 2, 128 channels,
8 heads x 16; BEV self-attention (P = 4), LiDAR cross (P = 8), camera cross (P = 8, 6 cameras).  L2 flushed between launches.
  ncu --set full --clock-control none --import-source on -k regex:sample_bwd_kernel -o gpurun_out/<name> python tools/profile_sample_bwd.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from unibev_b200 import ops, synth
from unibev_b200.plugin.encoder import anchor_heights


def main():
    dev = torch.device('cuda')
    torch.manual_seed(0)
    B, Nq, C, H, N = 2, 40000, 128, 8, 6
    metas = synth.img_metas(B)
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas], dtype=np.float32)).to(dev)
    ref_cam, mask = ops.project_points(l2i, anchor_heights(8, 4).tolist(), synth.PC_RANGE, 928, 1600, 200, 200)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    go = torch.randn(B, Nq, C, device=dev)

    def rows(P):
        return (torch.randn(B, Nq, 3 * H * P, device=dev) * 2.0).requires_grad_()
    for _ in range(2):
        for P, fhw in ((4, (200, 200)), (8, (180, 180))):
            v = torch.randn(B, fhw[0] * fhw[1], C, device=dev).requires_grad_()
            q = rows(P)
            out = ops.BevSampleFunction.apply(v, q, 200, 200, fhw[0], fhw[1], H, P, 0, 2 * H * P)
            flush.zero_()
            out.backward(go)
        v = torch.randn(B, N, 29 * 50, C, device=dev).requires_grad_()
        q = rows(8)
        out = ops.ImgSampleFunction.apply(v, q, ref_cam, mask, 200, 200, 29, 50, H, 8, 0, 2 * H * 8)
        flush.zero_()
        out.backward(go)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
