#!/usr/bin/env python
"""Time ub_linear_tf32x3 (3xTF32) next to the one-pass TF32 / fp16 tensor-core projections at the encoder's shapes
(M = 160000 rows = 4 frames x 40000 BEV queries).  CUDA events, L2 flushed between launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unibev_b200 import _cabi, ops
from tools.bench_gemm import timeit


def main():
    dev = 'cuda'
    M = int(os.environ.get('M', 160000))
    for opt in os.environ.get('OPTS', '1,0,4,0').split(';'):      # "inplace,direct,cluster,stagger_ns" variants of the 3xTF32 mode
        inplace, direct, cs, stag = (int(v) for v in opt.split(','))
        _cabi.check(_cabi.lib().ub_set_gemm_x3(inplace, direct, cs, stag), 'ub_set_gemm_x3')
        _cabi.lib().ub_set_gemm_cluster(cs)
        for name, N, K, ln in (('value 256x256', 256, 256, False), ('out+LN 256x256', 256, 256, True), ('qp 96', 96, 256, False),
                               ('qp 192', 192, 256, False), ('ffn1 512', 512, 256, False), ('ffn2+LN K512', 256, 512, True)):
            a = torch.randn(M, K, device=dev)
            w = torch.randn(N, K, device=dev) / 16
            b = torch.randn(N, device=dev)
            r = torch.randn(M, N, device=dev) if ln else None
            g, bt = torch.randn(N, device=dev), torch.randn(N, device=dev)
            lnp = (g, bt, 1e-5) if ln else None
            out = torch.empty(M, N, device=dev)
            ws = ops.split_tf32(w)
            a16, w16 = a.half(), w.half()
            if os.environ.get('ONCE') == '1':      # under ncu: one launch per shape and split mode
                ops.linear_tf32x3(a, ws, b, residual=r, ln=lnp, out=out)
                ops.linear_f16x3(a, ops.split_f16(w, 64.0), b, residual=r, ln=lnp, out=out)
                torch.cuda.synchronize()
                continue
            t3 = timeit(lambda: ops.linear_tf32x3(a, ws, b, residual=r, ln=lnp, out=out))
            wh = ops.split_f16(w, 64.0)
            t6 = timeit(lambda: ops.linear_f16x3(a, wh, b, residual=r, ln=lnp, out=out))
            t1 = timeit(lambda: ops.linear_tf32(a, w, b, residual=r, ln=lnp, out=out))
            th = timeit(lambda: ops.linear_f16(a16, w16, b, residual=r, ln=lnp, out=out))
            bytes_ = 4 * (M * K + M * N * (2 if ln else 1))
            print('inplace=%d direct=%d cs=%d stagger=%d %-16s x3 %7.1f us (%5.1f TFLOP/s eff, %5.0f GB/s) | f16x3 %7.1f us | tf32 %7.1f us | f16 %7.1f us' %
                  (inplace, direct, cs, stag, name, t3, 2.0 * M * N * K / t3 / 1e6, bytes_ / t3 / 1e3, t6, t1, th), flush=True)


if __name__ == '__main__':
    main()
