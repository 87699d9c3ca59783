"""Object-query decoder: oracle vs the golden vector frozen from the reference's own decoder.py, and
the plugin classes (deformable cross-attention through ub_msda_fwd) vs both."""
import json

import pytest
import torch

from oracle import unibev_decoder as od
from tests.helpers import load_golden

TOL = dict(rtol=1e-4, atol=2e-5)


def _case():
    a, p = load_golden('decoder')
    cfg = json.loads(str(a['cfg_json']))
    return a, p, cfg


def test_oracle_matches_reference_decoder():
    a, p, cfg = _case()
    L = cfg['num_layers']
    heads = cfg['transformerlayers']['attn_cfgs'][0]['num_heads']
    reg = od.reg_branches_from(p, L)
    inter, ref = od.decoder_forward(p, a['query'], a['value'], a['query_pos'], a['reference_points'], a['bev_hw'], L,
                                    num_heads=heads, reg_branches=reg)
    torch.testing.assert_close(inter, a['inter_states'], **TOL)
    torch.testing.assert_close(ref, a['inter_references'], **TOL)
    inter, ref = od.decoder_forward(p, a['query'], a['value'], a['query_pos'], a['reference_points'], a['bev_hw'], L,
                                    num_heads=heads, reg_branches=None)
    torch.testing.assert_close(inter, a['inter_states_noreg'], **TOL)
    torch.testing.assert_close(ref[-1], a['inter_references_noreg'][-1], rtol=0, atol=0)
    assert float((a['inter_references'][-1] - a['reference_points']).abs().max()) > 1e-3     # refinement moved the points


def test_plugin_decoder_builds_with_reference_keys():
    import unibev_b200.plugin  # noqa: F401
    from unibev_b200.registry import build_transformer_layer_sequence
    a, p, cfg = _case()
    dec = build_transformer_layer_sequence(cfg)
    own = {k for k in p if not k.startswith('reg.')}
    assert set(dec.state_dict().keys()) == own
    dec.load_state_dict({k: p[k] for k in own})
    with pytest.raises(ValueError):
        bad = json.loads(json.dumps(cfg))
        bad['transformerlayers']['operation_order'] = ('self_attn', 'norm', 'ffn', 'norm')
        build_transformer_layer_sequence(bad)


def _plugin_decoder(p, cfg, device):
    import unibev_b200.plugin  # noqa: F401
    from unibev_b200.registry import build_transformer_layer_sequence
    dec = build_transformer_layer_sequence(cfg)
    dec.load_state_dict({k: v for k, v in p.items() if not k.startswith('reg.')})
    L = cfg['num_layers']
    C = p['layers.0.norms.0.weight'].numel()
    reg = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(C, C), torch.nn.ReLU(), torch.nn.Linear(C, 10))
                               for _ in range(L)])
    reg.load_state_dict({k[4:]: v for k, v in p.items() if k.startswith('reg.')})
    return dec.to(device).eval(), reg.to(device).eval()


@pytest.mark.gpu
def test_gpu_plugin_decoder_matches_reference():
    a, p, cfg = _case()
    dec, reg = _plugin_decoder(p, cfg, 'cuda')
    Hb, Wb = (int(v) for v in a['bev_hw'])
    kw = dict(key=None, value=a['value'].cuda(), query_pos=a['query_pos'].cuda(),
              spatial_shapes=torch.tensor([[Hb, Wb]], device='cuda'), level_start_index=torch.tensor([0], device='cuda'))
    from unibev_b200 import _cabi
    _cabi.reset_launch_count()
    with torch.no_grad():
        inter, ref = dec(a['query'].cuda(), reference_points=a['reference_points'].cuda(), reg_branches=reg,
                         cls_branches=None, **kw)
        inter2, ref2 = dec(a['query'].cuda(), reference_points=a['reference_points'].cuda(), reg_branches=None, **kw)
    assert _cabi.launch_count() == 2 * cfg['num_layers']            # one ub_msda_fwd per layer and run
    torch.testing.assert_close(inter.cpu(), a['inter_states'], **TOL)
    torch.testing.assert_close(ref.cpu(), a['inter_references'], **TOL)
    torch.testing.assert_close(inter2.cpu(), a['inter_states_noreg'], **TOL)


@pytest.mark.gpu
def test_gpu_decoder_full_size_vs_oracle():
    """cfgCNW sizes: 900 object queries, 200 x 200 BEV map (40 000 values), C = 256, 6 layers, batch 1."""
    import unibev_b200.plugin  # noqa: F401
    from unibev_b200.registry import build_transformer_layer_sequence
    C, L, nq, Hb = 256, 6, 900, 200
    cfg = dict(type='DetectionTransformerDecoder', num_layers=L, return_intermediate=True,
               transformerlayers=dict(
                   type='DetrTransformerDecoderLayer',
                   attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                              dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1)],
                   ffn_cfgs=dict(type='FFN', embed_dims=C), feedforward_channels=2 * C, ffn_dropout=0.1,
                   operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))
    torch.manual_seed(0)
    dec = build_transformer_layer_sequence(cfg).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, prm in dec.named_parameters():                     # default init zeroes the offset / weight linears
            if n.endswith('sampling_offsets.weight'):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.02)
            elif n.endswith('attention_weights.weight'):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.1)
    query, query_pos = torch.randn(nq, 1, C, generator=g), torch.randn(nq, 1, C, generator=g)
    value = torch.randn(Hb * Hb, 1, C, generator=g)
    ref = torch.rand(1, nq, 3, generator=g)
    p = {k: v.detach() for k, v in dec.state_dict().items()}
    want, want_ref = od.decoder_forward(p, query, value, query_pos, ref, (Hb, Hb), L, num_heads=8)
    dec = dec.cuda()
    with torch.no_grad():
        got, got_ref = dec(query.cuda(), key=None, value=value.cuda(), query_pos=query_pos.cuda(),
                           reference_points=ref.cuda(), reg_branches=None,
                           spatial_shapes=torch.tensor([[Hb, Hb]], device='cuda'),
                           level_start_index=torch.tensor([0], device='cuda'))
    assert got.shape == (L, nq, 1, C)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(got_ref.cpu(), want_ref, rtol=0, atol=0)


@pytest.mark.gpu
def test_gpu_transformer_forward_runs_encoder_and_decoder():
    """UniBEVTransformer.forward (transformer_fusion.py:416-586): fused_bev_embed from the fused sm_100a pipeline feeds the
    decoder; the decoder half is checked against the oracle decoder fed with the same fused map."""
    from unibev_b200 import synth
    C, L, nq = 256, 2, 50
    dec_cfg = dict(type='DetectionTransformerDecoder', num_layers=L, return_intermediate=True,
                   transformerlayers=dict(
                       type='DetrTransformerDecoderLayer',
                       attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                                  dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1)],
                       ffn_cfgs=dict(type='FFN', embed_dims=C), feedforward_channels=2 * C, ffn_dropout=0.1,
                       operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))
    model, cfg = synth.build_model('unibev_nus_LC_cnw_256', num_layers=1, decoder=dec_cfg)
    assert model.decoder is not None
    model = model.cuda().eval()
    model.fused_precision = 'fp32'
    inp = synth.make_inputs('unibev_nus_LC_cnw_256', batch=1, bev_hw=(40, 40))
    g = torch.Generator().manual_seed(3)
    obj = torch.randn(nq, 2 * C, generator=g)
    with torch.no_grad():
        fused, inter, init_ref, inter_ref = model(
            [t.cuda() for t in inp['img_feats']], [t.cuda() for t in inp['pts_feats']], inp['bev_queries'].cuda(),
            obj.cuda(), 40, 40, bev_pos=inp['bev_pos'].cuda(), img_metas=inp['img_metas'])
    assert fused.shape == (1600, 1, C) and inter.shape == (L, nq, 1, C) and inter_ref.shape == (L, 1, nq, 3)
    p = {k[len('decoder.'):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith('decoder.')}
    query_pos, query = torch.split(obj, C, dim=1)
    want, want_ref = od.decoder_forward(p, query.unsqueeze(1), fused.cpu(), query_pos.unsqueeze(1), init_ref.cpu(),
                                        (40, 40), L, num_heads=8)
    torch.testing.assert_close(inter.cpu(), want, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(inter_ref.cpu(), want_ref, rtol=0, atol=0)
