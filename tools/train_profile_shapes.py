#!/usr/bin/env python
"""Which tensors the copy / clone / add / reduce kernels of a training step work on (torch.profiler with shapes + stacks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from unibev_b200 import synth
from unibev_b200.train import GradBuckets, train_step


def main():
    dev = torch.device('cuda')
    torch.backends.cuda.matmul.allow_tf32 = True
    np.random.seed(0)
    model, _ = synth.build_model('unibev_nus_LC_cat_128', drop_modality=0.5)
    model = model.to(dev).train()
    inp = synth.make_inputs('unibev_nus_LC_cat_128', batch=2, seed=1, device=dev)
    emb = torch.nn.Parameter(inp['bev_queries'].clone())
    params = list(model.parameters()) + [emb]
    opt = torch.optim.AdamW(params, lr=2e-4, fused=True)
    buckets = GradBuckets(params)
    for _ in range(3):
        train_step(model, emb, inp, opt, buckets)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        train_step(model, emb, inp, opt, buckets)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
        if e.key in ('aten::copy_', 'aten::clone', 'aten::add', 'aten::sum', 'aten::mul', 'aten::contiguous', 'aten::cat',
                     'aten::add_', 'aten::fill_', 'aten::zero_', 'aten::repeat', 'aten::square', 'aten::mean'):
            stack = [s for s in e.stack if 'unibev_b200' in s or 'bench' in s or 'tools/' in s][:3]
            rows.append((e.device_time_total, e.key, e.count, str(e.input_shapes)[:90], ' <- '.join(s.split('/')[-1][:60] for s in stack)))
    rows.sort(reverse=True)
    for t, k, n, shp, st in rows[:40]:
        print(f'{t:9.1f} us  {k:16s} x{n:<3d} {shp:90s} {st}')


if __name__ == '__main__':
    main()
