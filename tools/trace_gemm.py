#!/usr/bin/env python
"""Timeline of one ub_linear_tf32 launch: per-role event timestamps (globaltimer) of a few CTAs."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unibev_b200 import _cabi, ops


def main():
    M, N, K = int(os.environ.get('UB_TRACE_M', 40000)), int(sys.argv[1]) if len(sys.argv) > 1 else 256, 256
    cs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    mode = sys.argv[3] if len(sys.argv) > 3 else 'plain'
    dev = 'cuda'
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / 16
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev)
    g = torch.randn(N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _cabi.lib().ub_set_gemm_cluster(cs)

    x16, w16 = x.half(), w.half()
    ws = ops.split_tf32(w)
    wh = ops.split_f16(w, 64.0)

    def run():
        if mode == 'ln':
            return ops.linear_tf32(x, w, b, residual=r, ln=(g, g, 1e-5))
        if mode == 'f16':
            return ops.linear_f16(x16, w16, b)
        if mode == 'x3':
            return ops.linear_tf32x3(x, ws, b)
        if mode == 'h3':
            return ops.linear_f16x3(x, wh, b)
        if mode == 'h3ln':
            return ops.linear_f16x3(x, wh, b, residual=r, ln=(g, g, 1e-5))
        if mode == 'x3ln':
            return ops.linear_tf32x3(x, ws, b, residual=r, ln=(g, g, 1e-5))
        if mode == 'f16ln':
            return ops.linear_f16(x16, w16, b, residual=r, ln=(g, g, 1e-5), f16_out=True)
        return ops.linear_tf32(x, w, b)
    for _ in range(3):
        run()
    trace = torch.zeros(148 * 128, dtype=torch.int64, device=dev)
    flush.zero_()
    sink = flush[:160 << 20].sum()
    torch.cuda.synchronize()
    _cabi.lib().ub_set_gemm_trace(ctypes.c_void_p(trace.data_ptr()))
    run()
    torch.cuda.synchronize()
    _cabi.lib().ub_set_gemm_trace(None)
    t = trace.cpu().view(148, 128)
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    for cta in (0, 1, 73, 147):
        row = t[cta]
        if int(row[0]) == 0:
            continue
        rel = lambda v: (int(v) - t0) / 1e3 if int(v) else None
        prod = [rel(v) for v in row[1:41] if int(v)]
        mma = [rel(v) for v in row[41:81] if int(v)]
        epi = [rel(v) for v in row[81:121] if int(v)]
        print(f'CTA {cta}: start {rel(row[0]):.2f} us')
        print('  producer issued k-blocks at :', ' '.join('%.2f' % v for v in prod), '  (h3: converter warp 0, four stamps per 64-k block = operand slot free, A k-block 0 landed, A k-block 1 landed, slot published)')
        print('  mma saw k-blocks full at    :', ' '.join('%.2f' % v for v in mma),
              '   (x3: four stamps per k-block = A landed, a_lo ready, W_hi landed, W_lo landed; h3: four per 64-k block = start, operand slot ready, W_hi landed, W_lo landed)')
        print('  epilogue (acc ready, done)  :', ' '.join('%.2f' % v for v in epi))
    ends = [max(int(v) for v in t[c] if int(v)) for c in range(148) if int(t[c, 0])]
    print('last event over CTAs: %.2f us after first start' % ((max(ends) - t0) / 1e3))


if __name__ == '__main__':
    main()
