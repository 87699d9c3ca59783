"""CPU-only tests: plugin surface, registry/config handling, ABI symbol table, loud failure without CUDA,
batch sharding over gloo (world_size 2)."""
import copy
import os
import re
import subprocess
import sys

import pytest
import torch

from tests.helpers import ENCODER_HALF_TAGS, encoder_half_inputs, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def _built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()


def test_header_symbols_are_exported_and_bound():
    import ctypes
    from unibev_b200 import _cabi
    header = open(os.path.join(ROOT, 'include', 'unibev_b200.h')).read()
    declared = set(re.findall(r'\b(ub_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_cabi.PROTOTYPES), declared ^ set(_cabi.PROTOTYPES)
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f'{name} not exported by libunibev_b200.so'
    assert _cabi.lib().ub_version() == 1000
    assert _cabi.lib().ub_last_error() is not None


def test_library_targets_sm100a_only():
    out = subprocess.run(['cuobjdump', '-lelf', os.path.join(ROOT, 'unibev_b200', 'libunibev_b200.so')],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_argument_validation_happens_before_any_launch():
    """Bad shapes are rejected on the host side of the ABI (no GPU needed): error code + message."""
    from unibev_b200 import _cabi
    lib = _cabi.lib()
    rc = lib.ub_add_layernorm(None, None, None, None, None, None, 10, 256, 1e-5, None)
    assert rc == -1 and b'null pointer' in lib.ub_last_error()
    rc = lib.ub_bev_sample_fwd(1, 1, 1, 1, 10, 10, 10, 10, 8, 12, 8, 192, 0, 128, None)
    assert rc == -1 and b'head dim 12' in lib.ub_last_error()
    rc = lib.ub_img_sample_fwd(16, 16, 16, 16, 16, 1, 6, 10, 10, 5, 5, 8, 32, 6, 4, 192, 0, 96, None)
    assert rc == -1 and b'multiple of the 4 Z-anchors' in lib.ub_last_error()
    rc = lib.ub_cnw_fuse(None, None, None, None, None, None, None, 16, 1, 1, 4, 0, 1, 1, None)
    assert rc == -1


def test_ops_refuse_cpu_tensors():
    from unibev_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.add_layernorm(torch.randn(4, 32), torch.ones(32), torch.zeros(32))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.bev_sample(torch.randn(1, 16, 32), torch.randn(1, 16, 96), 4, 4, 4, 4, 4, 8, 0, 64)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'unibev_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f


def test_registered_names_match_reference_plugin():
    import unibev_b200.plugin  # noqa: F401
    from unibev_b200 import registry as R
    for name in ('SpatialCrossAttentionImg', 'MSDeformableAttention3DImg', 'SpatialCrossAttentionPts',
                 'MSDeformableAttention3DPts', 'MultiScaleDeformableAttention', 'MSDeformableAttention3DUniQueryImg'):
        assert R.ATTENTION.get(name) is not None, name
    assert R.ATTENTION.get('MSDeformableAttention3DUniQueryImg') is R.ATTENTION.get('MSDeformableAttention3DImg')
    for name in ('ImgLayer', 'PtsLayer'):
        assert R.TRANSFORMER_LAYER.get(name) is not None
    for name in ('ImgEncoder', 'PtsEncoder'):
        assert R.TRANSFORMER_LAYER_SEQUENCE.get(name) is not None
    assert R.TRANSFORMER.get('UniBEVTransformer') is not None


@pytest.mark.parametrize('tag', ENCODER_HALF_TAGS)
def test_state_dict_keys_equal_the_reference_modules(tag):
    """Fixtures hold state_dicts of the reference's own UniBEVTransformer: ours must load them strictly."""
    from unibev_b200.registry import build_transformer
    import unibev_b200.plugin  # noqa: F401
    a, p = load_golden('encoder_half_' + tag)
    cfg = encoder_half_inputs(a)[0]
    cfg.pop('decoder')
    m = build_transformer(cfg)
    assert set(m.state_dict()) == set(p)
    m.load_state_dict(p, strict=True)
    for k, v in m.state_dict().items():
        assert v.shape == p[k].shape, k


def test_full_config_builds_with_reference_shapes():
    from unibev_b200 import synth
    model, cfg = synth.build_model('unibev_nus_LC_cnw_256')
    sd = model.state_dict()
    assert sd['img_level_embeds'].shape == (4, 256)           # num_feature_levels default 4 (transformer_fusion.py:62)
    assert sd['cams_embeds'].shape == (6, 256)
    assert sd['img_bev_encoder.layers.2.attentions.1.deformable_attention.sampling_offsets.weight'].shape == (128, 256)
    assert sd['pts_bev_encoder.layers.0.attentions.0.sampling_offsets.weight'].shape == (64, 256)   # mmcv default 4 points
    assert sd['img_bev_encoder.layers.0.ffns.0.layers.0.0.weight'].shape == (512, 256)
    assert sd['pts_channel_weights'].shape == (256,)
    n_enc = sum(v.numel() for k, v in sd.items() if 'encoder' in k)
    assert n_enc == 3_609_792 - 2 * 3 * 0 or n_enc > 3_000_000
    camera_only, _ = synth.build_model('unibev_nus_C')       # type MSDeformableAttention3DUniQueryImg resolves
    assert not camera_only.with_pts_bev_encoder and camera_only.with_img_bev_encoder


def test_init_weights_semantics():
    import math
    from unibev_b200.plugin import MSDeformableAttention3DImg
    att = MSDeformableAttention3DImg(embed_dims=64, num_heads=8, num_levels=1, num_points=8)
    assert float(att.sampling_offsets.weight.abs().max()) == 0.0 and float(att.attention_weights.weight.abs().max()) == 0.0
    b = att.sampling_offsets.bias.view(8, 1, 8, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.arange(1, 9.0)) and torch.allclose(b[0, 0, :, 1], torch.zeros(8), atol=1e-6)
    assert torch.allclose(b[2, 0, 3], torch.tensor([0.0, 4.0]), atol=1e-5)        # head 2 looks along +y
    assert att.output_proj is None


def test_constructor_errors_match_reference():
    from unibev_b200.plugin import MSDeformableAttention3DPts, UniBEVTransformer
    with pytest.raises(ValueError, match='embed_dims must be divisible by num_heads'):
        MSDeformableAttention3DPts(embed_dims=30, num_heads=8)
    with pytest.raises(ValueError, match='Unrecognizable fusion method'):
        UniBEVTransformer(fusion_method='sum')


def test_modality_dropout_flag_draw_follows_numpy_stream():
    """transformer_fusion.py:227-228,465-477: two np.random draws decide (c_flag, l_flag) in training mode."""
    import numpy as np
    from unibev_b200 import synth
    model, _ = synth.build_model('unibev_nus_LC_cnw_256', num_layers=1)
    model.train()
    np.random.seed(3)
    draws = []
    for _ in range(50):
        model._draw_flags([0], [0])
        draws.append((model.c_flag, model.l_flag))
    np.random.seed(3)
    want = []
    for _ in range(50):
        c = l = 1
        if np.random.random() < 0.5:
            l = (np.random.random() < 0.5) * 1
            c = 1 - l
        want.append((c, l))
    assert draws == want and {(1, 1), (1, 0), (0, 1)} == set(draws)
    model.eval()
    model._draw_flags([0], None)
    assert (model.c_flag, model.l_flag) == (1, 0)


def test_shard_range_partitions():
    from unibev_b200.shard import shard_range
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 8):
            got = [i for r in range(world) for i in shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from unibev_b200.shard import shard_range, max_over_ranks, sum_over_ranks
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
mine = list(shard_range(5, r, w))
assert max_over_ranks(10.0 + r) == 10.0 + (w - 1)
assert sum_over_ranks(len(mine)) == 5
gathered = [None] * w
dist.all_gather_object(gathered, mine)
assert sorted(sum(gathered, [])) == list(range(5)), gathered
dist.barrier()
if r == 0:
    print('SHARD_OK', gathered)
dist.destroy_process_group()
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29531', str(script), ROOT],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'SHARD_OK' in out.stdout


_GRAD_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from unibev_b200.train import GradBuckets
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)                                   # same weights on every rank (the reference seeds ranks alike)
net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
unused = torch.nn.Linear(3, 3)                         # never part of rank 1's graph: contributes zeros there
never = torch.nn.Linear(2, 2)                          # no rank uses it: its grad must stay None (ADVICE r1)
params = list(net.parameters()) + list(unused.parameters()) + list(never.parameters())
buckets = GradBuckets(params, bucket_bytes=64)         # tiny buckets: several collectives, launched during backward
assert len(buckets.buckets) > 2
xs = [torch.randn(7, 6, generator=torch.Generator().manual_seed(10 + k)) for k in range(w)]
def loss_of(k, with_unused):
    y = net(xs[k])
    if with_unused:
        y = unused(y)
    return y.square().mean()
for step in range(2):                                  # hooks re-arm after finish()
    for p in params:
        p.grad = None
    loss_of(r, with_unused=(r == 0)).backward()
    buckets.finish()
    got = [p.grad.clone() if p.grad is not None else None for p in params]
    want_loss = sum(loss_of(k, with_unused=(k == 0)) for k in range(w)) / w
    want = torch.autograd.grad(want_loss, params, allow_unused=True)      # (no .grad accumulation: the hooks stay quiet)
    for p, g, wg in zip(params, got, want):
        if wg is None:
            assert g is None
        else:
            torch.testing.assert_close(g, wg, rtol=1e-6, atol=1e-7)
    assert all(p.grad is None for p in never.parameters())
# gradient views: prepare() points every p.grad at its bucket slot, autograd accumulates in place, finish() copies nothing
for step in range(2):
    buckets.prepare()
    flat_ptrs = {p: p.grad.data_ptr() for p in params}
    loss_of(r, with_unused=(r == 0)).backward()
    buckets.finish()
    want_loss = sum(loss_of(k, with_unused=(k == 0)) for k in range(w)) / w
    want = torch.autograd.grad(want_loss, params, allow_unused=True)
    for p, wg in zip(params, want):
        if wg is None:
            assert p.grad is None
        else:
            assert p.grad.data_ptr() == flat_ptrs[p]           # still the bucket slot: no copy back
            torch.testing.assert_close(p.grad, wg, rtol=1e-6, atol=1e-7)
# recorded pass (GraphedTrainStep, exchange='after'): the hooks put nothing on the wire, the book-keeping is dropped, and
# exchange_all() averages the bucket buffers -- which ARE the gradients -- in place
buckets.prepare()
buckets.defer_launch = True
loss_of(r, with_unused=(r == 0)).backward()
buckets.defer_launch = False
assert buckets._next == 0 and not buckets._works
buckets.reset_pass()
buckets.exchange_all()
want = torch.autograd.grad(sum(loss_of(k, with_unused=(k == 0)) for k in range(w)) / w, params, allow_unused=True)
for p, wg in zip(params, want):
    torch.testing.assert_close(p.grad, wg if wg is not None else torch.zeros_like(p), rtol=1e-6, atol=1e-7)
# uniform_usage=True (ranks seeded alike, every rank touches the same parameters): no flag exchange, same result
ub = GradBuckets(list(net.parameters()) + list(never.parameters()), bucket_bytes=64, uniform_usage=True)
buckets.remove()
ub.prepare()
loss_of(r, with_unused=False).backward()
ub.finish()
want = torch.autograd.grad(sum(loss_of(k, with_unused=False) for k in range(w)) / w, list(net.parameters()))
for p, wg in zip(net.parameters(), want):
    torch.testing.assert_close(p.grad, wg, rtol=1e-6, atol=1e-7)
assert all(p.grad is None for p in never.parameters())
ub.remove()
buckets = GradBuckets(params, bucket_bytes=64)
# two backward passes before finish(): no bucket is on the wire yet (bucket 0 holds the never-used parameters and waits for
# finish()), so the accumulated gradients are exchanged -- never a stale first-pass copy (ADVICE r1); had a bucket already
# been all-reduced, the second pass would raise instead
for p in params:
    p.grad = None
loss_of(r, with_unused=True).backward()
loss_of(r, with_unused=True).backward()
buckets.finish()
want = torch.autograd.grad(2 * sum(loss_of(k, with_unused=True) for k in range(w)) / w, params, allow_unused=True)
for p, wg in zip(params, want):
    if wg is not None:
        torch.testing.assert_close(p.grad, wg, rtol=1e-6, atol=1e-7)
dist.barrier()
if r == 0:
    print('GRAD_OK', len(buckets.buckets), buckets.nbytes())
dist.destroy_process_group()
'''


def test_gradient_buckets_average_over_two_gloo_ranks(tmp_path):
    """N > 1 training path (BASELINE configs[4]): bucketed gradient all-reduce == the gradient of the mean loss over
    the ranks' shards, including a parameter group one rank never touched (modality dropout)."""
    script = tmp_path / 'grad_worker.py'
    script.write_text(_GRAD_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29533', str(script), ROOT],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, (out.stdout[-1000:], out.stderr[-3000:])
    assert 'GRAD_OK' in out.stdout


def test_learned_positional_encoding_matches_oracle():
    """bev_pos producer (unibev_head.py:179-182): plugin class vs the oracle restatement of mmdet's module."""
    import torch
    from oracle.mmcv_semantics import learned_pos_encoding
    from unibev_b200.plugin import LearnedPositionalEncoding
    torch.manual_seed(0)
    enc = LearnedPositionalEncoding(num_feats=8, row_num_embed=6, col_num_embed=7)
    assert set(enc.state_dict()) == {'row_embed.weight', 'col_embed.weight'}
    pos = enc(torch.zeros(3, 5, 7))
    assert pos.shape == (3, 16, 5, 7)
    p = {'pe.' + k: v.detach() for k, v in enc.state_dict().items()}
    torch.testing.assert_close(pos, learned_pos_encoding(p, 'pe', 3, 5, 7), rtol=0, atol=0)
    assert torch.equal(pos[0], pos[2])                                   # batch-invariant
    assert torch.equal(pos[0, :8, 0], pos[0, :8, 4]) and torch.equal(pos[0, 8:, :, 0], pos[0, 8:, :, 6])
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 7, 7))
