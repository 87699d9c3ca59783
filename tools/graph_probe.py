#!/usr/bin/env python
"""Is the fused pipeline CPU-launch-bound?  Times eager steps vs replays of one captured CUDA graph of the same frame."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unibev_b200 import synth


def main():
    dev = torch.device('cuda')
    prec = sys.argv[1] if len(sys.argv) > 1 else 'tf32'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    model, _ = synth.build_model('unibev_nus_LC_cnw_256')
    model = model.to(dev).eval()
    model.fused_precision = prec
    inp = synth.make_inputs('unibev_nus_LC_cnw_256', batch=B)
    img = [inp['img_feats'][0].to(dev)]
    pts = [inp['pts_feats'][0].to(dev)]
    q, pos = inp['bev_queries'].to(dev), inp['bev_pos'].to(dev)

    import numpy as np
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in inp['img_metas']], dtype=np.float32)).to(dev)

    def step():
        with torch.no_grad():
            return model.encode(img, pts, q, 200, 200, bev_pos=pos, img_metas=inp['img_metas'], lidar2img=l2i,
                                img_shape=(928, 1600))

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    eager = timeit(step)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        out = step()
    graph = timeit(g.replay)
    ref = step()
    g.replay()
    torch.cuda.synchronize()
    print(f'{prec} batch {B}: eager {eager:.3f} ms/step ({B * 1e3 / eager:.0f} fps) | graph replay {graph:.3f} ms/step '
          f'({B * 1e3 / graph:.0f} fps) | max|graph - eager| = {float((out - ref).abs().max()):.2e}')


if __name__ == '__main__':
    main()
