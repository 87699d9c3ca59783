#!/usr/bin/env python
"""Numbers BASELINE.md section 5 asks for beside the headline: (1) ub_hard_voxelize + ub_voxel_mean on the config-2 cloud
(262 144 points x 5, 0.075 m voxels): points/s and achieved bytes/s against the per-point traffic; (2) ub_msda_fwd / ub_msda_bwd
at the encoder's LiDAR cross-attention shape: achieved GB/s of the SURVEY 8(d) byte count; (3) unibev_nus_L (configs[1]) frames/s
through the fused pipeline.  CUDA events, L2 flushed between launches.  One JSON object on stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unibev_b200 import ops, synth
from tools.bench_gemm import timeit


def main():
    dev = 'cuda'
    out = {}
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
    peak = peaks.get('hbm_gbs', 6650.0)
    # (1) voxelisation
    pts = torch.from_numpy(synth.make_cloud()).to(dev)
    vl = synth.VOXEL_LAYER
    for mode, max_vox in (('train', vl['max_voxels'][0]), ('test', vl['max_voxels'][1])):
        fn = lambda: ops.hard_voxelize(pts, vl['voxel_size'], vl['point_cloud_range'], vl['max_num_points'], max_vox)  # noqa: E731
        us = timeit(fn, iters=10)
        voxels, coors, num, n = fn()
        M = int(n.item())
        us_mean = timeit(lambda: ops.voxel_mean(voxels, num, 5), iters=10)
        N, C, T = pts.shape[0], pts.shape[1], vl['max_num_points']
        # algorithmic bytes: every point read once (20 B) + its key written / sorted / read (8 B key + 4 B index, 2 passes)
        # + the voxel slab written (max_voxels x T x C x 4), coordinates and counts
        alg = N * C * 4 + N * 12 * 2 + max_vox * (T * C * 4 + 12 + 4)
        out[f'hard_voxelize_{mode}'] = {'points': N, 'voxels': M, 'us': us, 'points_per_s': N / us * 1e6, 'alg_bytes': alg,
                                        'achieved_gbs': alg / us / 1e3, 'frac_of_hbm_peak': alg / us / 1e3 / peak,
                                        'voxel_mean_us': us_mean}
    # (2) generic MSDA forward / backward at the LiDAR cross-attention shape of one frame
    B, Nv, H, D, Nq, P = 1, 180 * 180, 8, 32, 40000, 8
    g = torch.Generator().manual_seed(0)
    value = torch.randn(B, Nv, H, D, generator=g).to(dev).requires_grad_()
    loc = (torch.rand(B, Nq, H, 1, P, 2, generator=g)).to(dev).requires_grad_()
    w = torch.softmax(torch.randn(B, Nq, H, 1, P, generator=g), -1).to(dev).requires_grad_()
    shapes = torch.tensor([[180, 180]], device=dev)
    lsi = torch.tensor([0], device=dev)
    go = torch.randn(B, Nq, H * D, generator=g).to(dev)
    fwd = lambda: ops.msda_forward(value.detach(), shapes, lsi, loc.detach(), w.detach())   # noqa: E731
    us_f = timeit(fwd, iters=10)

    def bwd():
        o = ops.MultiScaleDeformableAttnFunction.apply(value, shapes, lsi, loc, w, 64)
        o.backward(go)
        value.grad = loc.grad = w.grad = None
    us_fb = timeit(bwd, iters=10)
    alg_f = 4 * (B * Nv * H * D + B * Nq * H * P * 3 + B * Nq * H * D)
    alg_b = 4 * (2 * B * Nv * H * D + 2 * B * Nq * H * P * 3 + B * Nq * H * D)      # + grad_value, grad_loc / grad_w, grad_out
    out['msda_fwd'] = {'shape': 'B=1 Nv=32400 H=8 D=32 Nq=40000 P=8', 'us': us_f, 'alg_bytes': alg_f,
                       'achieved_gbs': alg_f / us_f / 1e3, 'frac_of_hbm_peak': alg_f / us_f / 1e3 / peak}
    out['msda_bwd'] = {'us': us_fb - us_f, 'alg_bytes': alg_b, 'achieved_gbs': alg_b / (us_fb - us_f) / 1e3,
                       'frac_of_hbm_peak': alg_b / (us_fb - us_f) / 1e3 / peak, 'note': 'fwd+bwd minus fwd'}
    # (3) unibev_nus_L / unibev_nus_C through the fused pipeline, 1 frame per step
    for wl in ('unibev_nus_L', 'unibev_nus_C', 'unibev_nus_LC_cnw_256'):
        model, _ = synth.build_model(wl)
        model = model.to(dev).eval()
        inp = synth.make_inputs(wl, batch=1, device=dev)
        for prec in ('fp32', 'fp16'):
            model.fused_precision = prec
            run = lambda: model.encode(inp['img_feats'], inp['pts_feats'], inp['bev_queries'], 200, 200,  # noqa: E731
                                       bev_pos=inp['bev_pos'], img_metas=inp['img_metas'])
            with torch.no_grad():
                us = timeit(run, iters=10)
            out[f'{wl}_batch1_{prec}'] = {'ms_per_frame': us / 1e3, 'frames_per_s': 1e6 / us,
                                         'note': 'eager launches (no CUDA graph), L2 flushed between frames'}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
