#!/usr/bin/env python
"""The encoder's projection shapes at M = B * 40000 rows through ub_linear_f16 / ub_add_layernorm16:
CUDA-event timing (L2 flushed between launches), or -- with `ncu` in front -- one launch per flavour:
  ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -o gpurun_out/<name> \
      python tools/profile_gemm.py 4 once
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unibev_b200 import ops


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    once = len(sys.argv) > 2 and sys.argv[2] == 'once'
    dev = 'cuda'
    M = B * 40000
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    a256 = torch.randn(M, 256, device=dev).half()
    a512 = torch.randn(M, 512, device=dev).half()
    w = {(n, k): (torch.randn(n, k, device=dev) / 16).half() for n, k in ((256, 256), (96, 256), (192, 256), (512, 256), (256, 512))}
    bias = {n: torch.randn(n, device=dev) for n in (96, 192, 256, 512)}
    r = torch.randn(M, 256, device=dev)
    g, bt = torch.randn(256, device=dev), torch.randn(256, device=dev)
    out = torch.empty(M, 256, device=dev)
    o16 = torch.empty(M, 512, device=dev, dtype=torch.float16)
    o16_256 = torch.empty(M, 256, device=dev, dtype=torch.float16)
    qp = torch.empty(M, 192, device=dev)
    planes = torch.empty(B, 8, 40000, 32, device=dev, dtype=torch.float16)
    cases = [
        ('out-proj plain fp32 out   (K256 N256)', lambda: ops.linear_f16(a256, w[256, 256], bias[256], out=out), 2 * M * 256 + 4 * M * 256),
        ('out-proj + LN, both outs  (K256 N256)', lambda: ops.linear_f16(a256, w[256, 256], bias[256], residual=r, ln=(g, bt, 1e-5), out=out, out16=o16_256), 2 * M * 256 + 4 * M * 256 * 2 + 2 * M * 256),
        ('add_layernorm16 alone               ', lambda: ops.add_layernorm(out, g, bt, residual=r, out=out, out16=o16_256), 4 * M * 256 * 3 + 2 * M * 256),
        ('value planes              (K256 N256)', lambda: ops.linear_f16(a256, w[256, 256], bias[256], planes_nv=40000), 2 * M * 256 * 2),
        ('qp cross fp32 out         (K256 N192)', lambda: ops.linear_f16(a256, w[192, 256], bias[192], out=qp), 2 * M * 256 + 4 * M * 192),
        ('ffn1 relu fp16 out        (K256 N512)', lambda: ops.linear_f16(a256, w[512, 256], bias[512], relu=True, fp32_out=False, out16=o16), 2 * M * 256 + 2 * M * 512),
        ('ffn2 plain fp32 out       (K512 N256)', lambda: ops.linear_f16(a512, w[256, 512], bias[256], out=out), 2 * M * 512 + 4 * M * 256),
        ('ffn2 + LN, both outs      (K512 N256)', lambda: ops.linear_f16(a512, w[256, 512], bias[256], residual=r, ln=(g, bt, 1e-5), out=out, out16=o16_256), 2 * M * 512 + 4 * M * 256 * 2 + 2 * M * 256),
    ]
    if os.environ.get('UB_STREAM_W') == '1':
        from unibev_b200 import _cabi
        _cabi.lib().ub_set_gemm_stream_w_with_residual(1)
    for name, fn, nbytes in cases:
        for _ in range(1 if once else 3):
            fn()
        torch.cuda.synchronize()
        tot, iters = 0.0, (1 if once else 10)
        for _ in range(iters):
            flush.zero_()
            sink = flush[:160 << 20].sum()      # leave L2 full of clean lines
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot * 1e3 / iters
        print('%s M=%d  %7.1f us  %6.0f GB/s (compulsory bytes %.0f MB)' % (name, M, us, nbytes / us / 1e3, nbytes / 1e6), flush=True)


if __name__ == '__main__':
    main()
