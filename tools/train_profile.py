#!/usr/bin/env python
"""Where a training step (tools/train_bench.py, 1 GPU) spends its time: torch.profiler table of the top device kernels
and the wall-clock / device-time split (a large gap = host-bound)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from unibev_b200 import synth
from unibev_b200.train import GradBuckets, train_step


def main():
    dev = torch.device('cuda')
    torch.backends.cuda.matmul.allow_tf32 = os.environ.get('UB_TRAIN_TF32', '1') == '1'     # as bench.py --mode train
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
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            train_step(model, emb, inp, opt, buckets)
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 3
    print('wall per step: %.1f ms' % (wall * 1e3))
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=70))


if __name__ == '__main__':
    main()
