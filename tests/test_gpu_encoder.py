"""GPU parity of the plugin modules (fused eval pipeline and module/autograd path) against the golden
vectors frozen from the reference and against the CPU oracle at BASELINE.json's sizes."""
import copy
import time

import pytest
import torch

from oracle import unibev_encoder as oe
from tests.helpers import ENCODER_HALF_TAGS, encoder_half_inputs, load_golden, metas_from

pytestmark = pytest.mark.gpu

# Stated tolerances (BASELINE.json north_star: encoder BEV-feature output within rtol 1e-3 of the reference).
# Outputs are post-LayerNorm, O(1); atol covers elements near zero.
TOL_FP32 = dict(rtol=1e-3, atol=1e-4)      # fp32 GEMMs: same arithmetic as the reference up to summation order
TOL_TF32 = dict(rtol=1e-3, atol=5e-3)      # TF32-class path: tensor-core GEMMs (TF32 / fp16 operands, fp32 accumulate; the
                                           # activation operand of kind::tf32 is truncated, weights are pre-rounded) and
                                           # fp16-staged sampling; torch 1.10, the reference's stack, defaulted to TF32 too


def _build(cfg, params):
    from unibev_b200.registry import build_transformer
    import unibev_b200.plugin  # noqa: F401
    cfg = copy.deepcopy(cfg)
    cfg.pop('decoder', None)
    m = build_transformer(cfg)
    m.load_state_dict(params, strict=True)
    return m.cuda()


def _cuda(x):
    if x is None:
        return None
    if isinstance(x, (list, tuple)):
        return [_cuda(t) for t in x]
    return x.cuda()


@pytest.mark.parametrize('tag', ENCODER_HALF_TAGS)
@pytest.mark.parametrize('path', ['fused_fp32', 'fused_tf32', 'modules'])
def test_encoder_half_golden(tag, path):
    a, p = load_golden('encoder_half_' + tag)
    cfg, img, pts, q, bev_h, bev_w = encoder_half_inputs(a)
    m = _build(cfg, p)
    flags = tuple(int(v) for v in a['flags'])
    if path == 'modules':
        m.train()                             # dropout p = 0 in the fixtures: train mode only selects the module path
    else:
        m.eval()
        m.fused_precision = path.split('_')[1]
    with torch.no_grad():
        out = m.encode(_cuda(img), _cuda(pts), _cuda(q), bev_h, bev_w, bev_pos=a['bev_pos'].cuda(), flags=flags,
                       img_metas=metas_from(a))
    assert (m._fused is not None) == (path != 'modules')
    assert (m.c_flag, m.l_flag) == flags
    tol = TOL_TF32 if path == 'fused_tf32' else TOL_FP32
    torch.testing.assert_close(out.cpu(), a['fused'], **tol)


def test_fused_path_is_taken_and_native():
    from unibev_b200 import _cabi
    a, p = load_golden('encoder_half_lc_cnw_linear')
    cfg, img, pts, q, bev_h, bev_w = encoder_half_inputs(a)
    m = _build(cfg, p).eval()
    layers = cfg['img_encoder']['num_layers']
    counts = {}
    for prec in ('fp32', 'tf32'):
        m.fused_precision = prec
        _cabi.reset_launch_count()
        with torch.no_grad():
            m.encode(_cuda(img), _cuda(pts), _cuda(q), bev_h, bev_w, bev_pos=a['bev_pos'].cuda(),
                     img_metas=metas_from(a))
        assert m._fused is not None
        counts[prec] = _cabi.launch_count()
    # fp32: per encoder flatten + query broadcast + layers*(2 samples + 3 LN); + bev_pos flatten + project + fuse
    # (GEMMs: cuBLAS)
    assert counts['fp32'] == 2 * (2 + layers * 5) + 3
    # tf32: the covered projections run on the tcgen05 GEMM as well, and the residual + LayerNorm steps that follow a
    # covered projection happen inside it (no ub_add_layernorm pass): at least the two output projections and the two
    # FFN linears of every layer are ub_linear_* launches, at most 3 LayerNorm passes per layer disappear
    assert counts['tf32'] >= counts['fp32'] + 2 * layers * (4 - 3)
    assert counts['tf32'] != counts['fp32']


def test_module_path_backward_matches_oracle_autograd():
    """Training path: gradients through ub_msda_bwd vs autograd through the CPU oracle."""
    a, p = load_golden('encoder_half_lc_cnw_linear')
    cfg, img, pts, q, bev_h, bev_w = encoder_half_inputs(a)
    m = _build(cfg, p).train()
    m.drop_modality = None
    out = m.encode(_cuda(img), _cuda(pts), _cuda(q), bev_h, bev_w, bev_pos=a['bev_pos'].cuda(), img_metas=metas_from(a))
    g = torch.Generator().manual_seed(0)
    go = torch.randn(out.shape, generator=g)
    out.backward(go.cuda())
    pc = {k: v.clone().requires_grad_() for k, v in p.items()}
    want = oe.encoder_half(pc, cfg, img, pts, q, bev_h, bev_w, bev_pos=a['bev_pos'], img_metas=metas_from(a))
    want.backward(go)
    torch.testing.assert_close(out.detach().cpu(), want.detach(), **TOL_FP32)
    checked = 0
    for name, prm in m.named_parameters():
        if prm.grad is None:
            assert pc[name].grad is None or float(pc[name].grad.abs().max()) == 0.0, name
            continue
        ref = pc[name].grad
        scale = float(ref.abs().max()) + 1e-6
        err = float((prm.grad.cpu() - ref).abs().max())
        assert err <= 2e-3 * scale + 1e-5, (name, err, scale)
        checked += 1
    assert checked > 40


def _full_size(workload, batch, precision):
    from unibev_b200 import synth
    model, cfg = synth.build_model(workload)
    model = model.cuda().eval()
    model.fused_precision = precision
    inp = synth.make_inputs(workload, batch=batch)
    with torch.no_grad():
        out = model.encode(_cuda(inp['img_feats']), _cuda(inp['pts_feats']), inp['bev_queries'].cuda(),
                           inp['bev_h'], inp['bev_w'], bev_pos=inp['bev_pos'].cuda(), img_metas=inp['img_metas'])
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    torch.set_num_threads(max(1, torch.get_num_threads()))
    t0 = time.time()
    want = oe.encoder_half(params, cfg, inp['img_feats'], inp['pts_feats'], inp['bev_queries'], inp['bev_h'],
                           inp['bev_w'], bev_pos=inp['bev_pos'], img_metas=inp['img_metas'])
    dt = time.time() - t0
    err = (out.cpu() - want).abs()
    rel = err / (want.abs() + 1e-12)
    print(f'\n[{workload} B={batch} {precision}] oracle {dt:.1f}s  max|err|={err.max():.3e} mean|err|={err.mean():.3e} '
          f'p99.9 rel={rel.flatten().kthvalue(int(rel.numel() * 0.999)).values:.3e}')
    return out.cpu(), want


@pytest.mark.parametrize('workload,batch', [('unibev_nus_LC_cnw_256', 1), ('unibev_nus_C', 1), ('unibev_nus_L', 1),
                                            ('unibev_nus_LC_cat_128', 2)])
def test_full_size_vs_oracle_fp32(workload, batch):
    """BASELINE.json configs at full size, fp32 GEMMs: rtol 1e-3 / atol 1e-4."""
    out, want = _full_size(workload, batch, 'fp32')
    torch.testing.assert_close(out, want, **TOL_FP32)


@pytest.mark.parametrize('workload,batch', [('unibev_nus_LC_cnw_256', 1), ('unibev_nus_LC_cnw_256', 3), ('unibev_nus_C', 1),
                                            ('unibev_nus_L', 2), ('unibev_nus_LC_cat_128', 1)])
def test_full_size_vs_oracle_tf32(workload, batch):
    """The bench configuration (fp16 / TF32 tensor-core operands, window-staged sampling, fused epilogues) at full size,
    single-modality and 128-channel variants included: rtol 1e-3 / atol 4e-3."""
    out, want = _full_size(workload, batch, 'tf32')
    torch.testing.assert_close(out, want, **TOL_TF32)


def test_batch_items_are_independent_at_full_size():
    """Size-independent property: with identical calibrations, item b of a batch equals the same
    sample run alone (the path shards by sample; no cross-item state but the item-0 hit lists)."""
    from unibev_b200 import synth
    model, _ = synth.build_model('unibev_nus_LC_cnw_256')
    model = model.cuda().eval()
    inp = synth.make_inputs('unibev_nus_LC_cnw_256', batch=2)
    with torch.no_grad():
        both = model.encode(_cuda(inp['img_feats']), _cuda(inp['pts_feats']), inp['bev_queries'].cuda(), 200, 200,
                            bev_pos=inp['bev_pos'].cuda(), img_metas=inp['img_metas'])
        one = model.encode([inp['img_feats'][0][1:2].cuda()], [inp['pts_feats'][0][1:2].cuda()],
                           inp['bev_queries'].cuda(), 200, 200, bev_pos=inp['bev_pos'][1:2].cuda(),
                           img_metas=inp['img_metas'][1:2])
    torch.testing.assert_close(both[1:2], one, rtol=1e-5, atol=1e-5)
