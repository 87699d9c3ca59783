#!/usr/bin/env python
"""Where the error of each precision class comes from: fused_bev_embed of unibev_nus_LC_cnw_256 at full
size (batch 1) against the CPU oracle, one half of the pipeline switched at a time:

    gemm      tf32x3 (3xTF32, default) | tf32 (one TF32 pass, fp32 activations) | f16 (fp16 operands incl. input tokens)
    sampling  fp32 (fp32 value maps / weights) | win16 (fp16-staged value maps, fp16 weights, approximate softmax)

Prints max / mean / p99.9 absolute error per combination (outputs are post-LayerNorm, O(1))."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import unibev_encoder as oe          # checker only
from unibev_b200 import synth


def main():
    wl = 'unibev_nus_LC_cnw_256'
    model, cfg = synth.build_model(wl)
    model = model.cuda().eval()
    inp = synth.make_inputs(wl, batch=1)
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    want = oe.encoder_half(params, cfg, inp['img_feats'], inp['pts_feats'], inp['bev_queries'], 200, 200,
                           bev_pos=inp['bev_pos'], img_metas=inp['img_metas'])
    print('%-8s %-8s %12s %12s %12s   %s' % ('gemm', 'sampling', 'max|err|', 'mean|err|', 'p99.9|err|', 'meets rtol 1e-3 / atol 1e-4'))
    for gemm in ('tf32x3', 'tf32', 'f16'):
        for sampling in ('fp32', 'win16'):
            model.fused_precision = 'fp32'
            model.fused_overrides = {'gemm': gemm, 'sampling': sampling}
            with torch.no_grad():
                out = model.encode([inp['img_feats'][0].cuda()], [inp['pts_feats'][0].cuda()], inp['bev_queries'].cuda(), 200,
                                   200, bev_pos=inp['bev_pos'].cuda(), img_metas=inp['img_metas']).cpu()
            err = (out - want).abs()
            ok = bool((err <= 1e-4 + 1e-3 * want.abs()).all())
            p999 = err.flatten().kthvalue(int(err.numel() * 0.999)).values
            print('%-8s %-8s %12.3e %12.3e %12.3e   %s' % (gemm, sampling, err.max(), err.mean(), p999, 'yes' if ok else 'NO'),
                  flush=True)


if __name__ == '__main__':
    main()
