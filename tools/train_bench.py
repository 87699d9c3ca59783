#!/usr/bin/env python
"""Training step of the hot path, BASELINE configs[4]: unibev_nus_LC_cat_128 (C_enc = 128, fusion 'cat', modality
dropout 0.5), 2 samples per GPU, gradient all-reduce over NCCL (unibev_b200.train.GradBuckets), AdamW.

    python tools/train_bench.py                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/train_bench.py --steps K --warmup W                   # N GPUs, one rank each

A step = forward through the plugin's module path (ub_msda_fwd and the fused sampling kernels' autograd twins are the
generic ub_msda_fwd / ub_msda_bwd here) + synthetic scalar loss (mean square of fused_bev_embed) + backward with the
bucketed all-reduce overlapped + optimizer step.  Every rank uses the same numpy seed, as the reference does
(train_UniBEV.py:200-204), so the modality-dropout flags coincide across ranks.  Prints one JSON line on rank 0:
frames/s over all ranks (max-over-ranks CUDA-event time), bytes all-reduced per step."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from unibev_b200 import _cabi, synth
from unibev_b200.train import GradBuckets, train_step

WORKLOAD = 'unibev_nus_LC_cat_128'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=2, help='samples per GPU (configs[4]: 16 over 8 GPUs)')
    ap.add_argument('--bucket-mb', type=float, default=8.0)
    ap.add_argument('--fp32-matmul', action='store_true', help='strict fp32 (SIMT) GEMMs instead of TF32 tensor cores')
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    # TF32 tensor-core matmuls: what torch 1.10, the reference's stack, does by default on Ampere and later
    torch.backends.cuda.matmul.allow_tf32 = not args.fp32_matmul
    np.random.seed(0)                                    # same flags on every rank
    model, cfg = synth.build_model(WORKLOAD, drop_modality=0.5)
    model = model.to(dev).train()
    inp = synth.make_inputs(WORKLOAD, batch=args.batch, seed=1 + rank, device=dev)
    g = torch.Generator().manual_seed(7)                 # the query table is a parameter: same initial value on every rank
    bev_embedding = torch.nn.Parameter(torch.randn(inp['bev_queries'].shape, generator=g).to(dev))
    params = list(model.parameters()) + [bev_embedding]
    opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=0.01)
    buckets = GradBuckets(params, bucket_bytes=int(args.bucket_mb * (1 << 20)), uniform_usage=True)   # ranks seeded alike

    def step():
        return train_step(model, bev_embedding, inp, opt, buckets)

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _cabi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # parameters must agree across ranks after identical averaged updates
    check = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
    if world > 1:
        lo, hi = check.clone(), check.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.allclose(lo, hi, rtol=1e-9, atol=0))
    else:
        in_sync = True
    if rank == 0:
        print(json.dumps({
            'metric': 'nuScenes frames/sec train step (L+C cat-128, modality dropout)', 'unit': 'frames/s',
            'value': world * args.batch * args.steps / (float(ms.item()) / 1e3), 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': float(ms.item()) / args.steps, 'scaling': 'weak',
            'dtype': 'f32' if args.fp32_matmul else 'tf32',
            'data': 'synthetic', 'loss': float(loss), 'params_in_sync': in_sync,
            'allreduce_bytes_per_step': buckets.nbytes() if world > 1 else 0, 'buckets': len(buckets.buckets),
            'gpu_launches': _cabi.launch_count(),
            'config': {'workload': f'{WORKLOAD} training step, {args.batch} samples per GPU, {world} GPU(s), '
                                   'synthetic loss (mean square of fused_bev_embed), AdamW'}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
