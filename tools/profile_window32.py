#!/usr/bin/env python
"""Launch each fp32 window-staged sampling kernel at LC-CNW-256 size (for `ncu`): 2 launches per flavour, L2 flushed.
  ncu --set full --clock-control none --import-source on -k regex:win32_kernel -o gpurun_out/<name> python tools/profile_window32.py 4
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
    B, Nq, C, H, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 1), 40000, 256, 8, 6
    model, _ = synth.build_model('unibev_nus_LC_cnw_256', num_layers=1)
    model = model.to(dev)
    x = torch.randn(B, Nq, C, device=dev)
    li, lp = model.img_bev_encoder.layers[0], model.pts_bev_encoder.layers[0]

    def qproj(att):
        return torch.cat((att.sampling_offsets(x), att.attention_weights(x)), -1).contiguous()
    with torch.no_grad():
        qp_self, qp_pts, qp_img = qproj(lp.attentions[0]), qproj(lp.attentions[1].deformable_attention), \
            qproj(li.attentions[1].deformable_attention)
    metas = synth.img_metas(B)
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas], dtype=np.float32)).to(dev)
    ref_cam, mask = ops.project_points(l2i, anchor_heights(8, 4).tolist(), synth.PC_RANGE, 928, 1600, 200, 200)
    hits = ops.build_hits(mask)
    order = ops.hit_order(mask, ref_cam, hits)
    q_dst = order[0]
    qp_hit = torch.zeros(B, N * Nq, qp_img.shape[2], device=dev)
    for j in range(N):
        rows = q_dst[:, j] >= 0
        qp_hit[:, q_dst[rows, j].long()] = qp_img[:, rows]
    out = torch.empty(B, Nq, C, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(2):
        p_self = ops.value_to_planes32(torch.randn(B * Nq, C, device=dev), B, Nq, H)
        p_pts = ops.value_to_planes32(torch.randn(B * 180 * 180, C, device=dev), B, 180 * 180, H)
        p_img = ops.value_to_planes32(torch.randn(B * N * 1450, C, device=dev), B * N, 1450, H)
        flush.zero_()
        ops.bev_sample_win32(p_self, qp_self, 200, 200, 200, 200, H, 4, 0, 64, out=out)
        flush.zero_()
        ops.bev_sample_win32(p_pts, qp_pts, 200, 200, 180, 180, H, 8, 0, 128, out=out)
        flush.zero_()
        ops.img_sample_win32(p_img, qp_hit, order, hits, 200, 200, 29, 50, H, 8, 0, 128, out=out)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
