#!/usr/bin/env python
"""Time ub_linear_tf32 against cuBLAS (torch, TF32) + the separate epilogue kernels it replaces, at the encoder's
GEMM shapes (M = 40000 BEV queries).  CUDA events, L2 flushed between launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unibev_b200 import _cabi, ops

flush_buf = None


def timeit(fn, iters=20):
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for i in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(iters):
        flush_buf.zero_()
        flush_sink = flush_buf[:160 << 20].sum()      # leave L2 full of CLEAN lines: evictions cost nothing
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot * 1e3 / iters


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = 'cuda'
    M = 40000
    x = torch.randn(M, 256, device=dev)
    x5 = torch.randn(M, 512, device=dev)
    r = torch.randn(M, 256, device=dev)
    g, bt = torch.randn(256, device=dev), torch.randn(256, device=dev)
    for cs in (1, 2, 4):
        _cabi.lib().ub_set_gemm_cluster(cs)
        w = torch.randn(256, 256, device=dev) / 16
        b = torch.randn(256, device=dev)
        out = torch.empty(M, 256, device=dev)
        print('cluster %d: 256x256 plain %6.1f us | LN %6.1f us | N=96 %6.1f us' % (
            cs, timeit(lambda: ops.linear_tf32(x, w, b, out=out)),
            timeit(lambda: ops.linear_tf32(x, w, b, residual=r, ln=(g, bt, 1e-5), out=out)),
            timeit(lambda: ops.linear_tf32(x, w[:96].contiguous(), b[:96].contiguous()))), flush=True)
    _cabi.lib().ub_set_gemm_cluster(4)
    for name, N, K in (('value/out 256x256', 256, 256), ('qp 96', 96, 256), ('qp 192', 192, 256), ('ffn1 512', 512, 256),
                       ('ffn2 K512', 256, 512)):
        a = x5 if K == 512 else x
        w = torch.randn(N, K, device=dev) / 16
        b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)
        t_cublas = timeit(lambda: torch.addmm(b, a, w.t(), out=out))
        t_ours = timeit(lambda: ops.linear_tf32(a, w, b, out=out))
        bytes_ = 4 * (M * K + M * N + N * K)
        print('%-20s cuBLAS %7.1f us | tcgen05 %7.1f us (%5.0f GB/s, %5.1f TFLOP/s)' %
              (name, t_cublas, t_ours, bytes_ / t_ours / 1e3, 2.0 * M * N * K / t_ours / 1e6), flush=True)
        if N == 256:
            o2 = torch.empty(M, N, device=dev)
            t_ln = timeit(lambda: ops.add_layernorm(out, g, bt, bias=b, residual=r, out=o2))
            t_f = timeit(lambda: ops.linear_tf32(a, w, b, residual=r, ln=(g, bt, 1e-5), out=out))
            print('%-20s cuBLAS + add_layernorm %7.1f us | fused LN %7.1f us' % ('', t_cublas + t_ln, t_f), flush=True)
            if K == 256:
                pl = torch.empty(1, 8, M, 32, device=dev, dtype=torch.float16)
                t_c = timeit(lambda: ops.value_to_half(out, 1, M, 8, out=pl))
                t_p = timeit(lambda: ops.linear_tf32(a, w, b, planes_nv=M, out=pl))
                print('%-20s cuBLAS + value_to_half %7.1f us | fused planes %7.1f us' % ('', t_cublas + t_c, t_p), flush=True)


def main_f16():
    dev = 'cuda'
    M = 40000
    r = torch.randn(M, 256, device=dev)
    g, bt = torch.randn(256, device=dev), torch.randn(256, device=dev)
    for name, N, K in (('f16 256x256', 256, 256), ('f16 qp 96', 96, 256), ('f16 qp 192', 192, 256), ('f16 K512', 256, 512)):
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) / 16).half()
        b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)
        o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
        t_plain = timeit(lambda: ops.linear_f16(a, w, b, out=out))
        t_16 = timeit(lambda: ops.linear_f16(a, w, b, fp32_out=False, out16=o16))
        line = '%-14s fp32 out %6.1f us | fp16 out %6.1f us' % (name, t_plain, t_16)
        if N == 256:
            t_ln = timeit(lambda: ops.linear_f16(a, w, b, residual=r, ln=(g, bt, 1e-5), out=out, out16=o16))
            t_pl = timeit(lambda: ops.linear_f16(a, w, b, planes_nv=M))
            line += ' | LN + both outputs %6.1f us | planes %6.1f us' % (t_ln, t_pl)
        print(line, flush=True)


if __name__ == '__main__':
    main_f16()
    main()
