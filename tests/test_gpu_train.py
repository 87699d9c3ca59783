"""Training step of the hot path (BASELINE configs[4] shapes scaled down): forward through the plugin's module path
(ub_msda_fwd), backward (ub_msda_bwd), gradient buckets, optimizer step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_step_updates_parameters_and_matches_plain_autograd():
    from unibev_b200 import _cabi, synth
    from unibev_b200.train import GradBuckets, train_step
    np.random.seed(0)
    model, _ = synth.build_model('unibev_nus_LC_cat_128', num_layers=1, drop_modality=None, dropout=0.0)
    model = model.cuda().train()
    inp = synth.make_inputs('unibev_nus_LC_cat_128', batch=1, bev_hw=(24, 24), device='cuda')
    emb = torch.nn.Parameter(inp['bev_queries'].clone())
    params = list(model.parameters()) + [emb]

    # reference gradients: the same forward / loss with plain autograd, no buckets
    fused = model.encode(inp['img_feats'], inp['pts_feats'], emb, 24, 24, bev_pos=inp['bev_pos'], img_metas=inp['img_metas'])
    assert fused.shape == (1, 576, 256)                      # 'cat' fusion of two 128-channel BEV maps
    fused.square().mean().backward()
    want = [p.grad.clone() if p.grad is not None else None for p in params]
    for p in params:
        p.grad = None

    opt = torch.optim.SGD(params, lr=0.1)
    buckets = GradBuckets(params, bucket_bytes=1 << 16)
    before = [p.detach().clone() for p in params]
    _cabi.reset_launch_count()
    loss0 = float(train_step(model, emb, inp, opt, buckets))
    assert _cabi.launch_count() > 0                          # native forward + backward kernels ran
    moved = 0
    for p, b, g in zip(params, before, want):
        if g is None or float(g.abs().max()) == 0.0:
            continue
        torch.testing.assert_close(p.detach(), b - 0.1 * g, rtol=1e-4, atol=1e-5)     # SGD step with the same gradient
        moved += 1
    assert moved > 30
    loss1 = float(train_step(model, emb, inp, opt, buckets))
    assert np.isfinite(loss0) and np.isfinite(loss1) and loss1 < loss0
