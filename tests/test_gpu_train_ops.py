"""Streaming backward helpers of the training step (ub_colsum, ub_layernorm_bwd) and the autograd functions built on them
(ops.LinearFunction, ops.LayerNormFunction), through the C ABI, against torch's own fp32 autograd."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M,N', [(1, 4), (7, 96), (1000, 128), (80000, 128), (4097, 192), (333, 256), (50, 1024), (12345, 36)])
def test_colsum(M, N):
    from unibev_b200 import ops
    x = torch.randn(M, N, device='cuda', generator=torch.Generator('cuda').manual_seed(M + N))
    got = ops.colsum(x)
    want = x.double().sum(0)
    torch.testing.assert_close(got.double(), want, rtol=1e-5, atol=1e-4 * (M ** 0.5))
    acc = torch.ones(N, device='cuda')
    ops.colsum(x, out=acc)                                   # accumulates into `out`
    torch.testing.assert_close(acc.double(), want + 1.0, rtol=1e-5, atol=1e-4 * (M ** 0.5))


def test_colsum_rejects_uncovered_width():
    from unibev_b200 import _cabi, ops
    with pytest.raises(_cabi.UnsupportedShape):
        ops.colsum(torch.randn(5, 6, device='cuda'))
    assert not ops.train_ops_supported(6) and ops.train_ops_supported(128)


@pytest.mark.parametrize('with_residual', [False, True])
@pytest.mark.parametrize('rows,C', [(1, 128), (37, 128), (80000, 128), (5000, 256), (129, 96), (64, 512), (33, 1024), (10, 4)])
def test_layernorm_function_vs_torch(rows, C, with_residual):
    from unibev_b200 import ops
    g = torch.Generator('cuda').manual_seed(rows + C)
    x = (torch.randn(rows, C, device='cuda', generator=g) * 2 + 0.5)
    res = torch.randn(rows, C, device='cuda', generator=g) if with_residual else None
    gamma = torch.randn(C, device='cuda', generator=g)
    beta = torch.randn(C, device='cuda', generator=g)
    go = torch.randn(rows, C, device='cuda', generator=g)
    xa, ga, ba = (t.clone().requires_grad_() for t in (x, gamma, beta))
    ra = res.clone().requires_grad_() if with_residual else None
    ya = ops.LayerNormFunction.apply(xa, ra, ga, ba, 1e-5)
    ya.backward(go)
    xb, gb, bb = (t.double().clone().requires_grad_() for t in (x, gamma, beta))
    rb = res.double().clone().requires_grad_() if with_residual else None
    yb = F.layer_norm(xb + rb if with_residual else xb, (C,), gb, bb, 1e-5)
    yb.backward(go.double())
    torch.testing.assert_close(ya.detach().double(), yb.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(xa.grad.double(), xb.grad, rtol=1e-4, atol=1e-5)
    if with_residual:       # the gradient of the sum goes to both terms
        torch.testing.assert_close(ra.grad.double(), rb.grad, rtol=1e-4, atol=1e-5)
    scale = rows ** 0.5
    torch.testing.assert_close(ga.grad.double(), gb.grad, rtol=1e-4, atol=1e-5 * scale)
    torch.testing.assert_close(ba.grad.double(), bb.grad, rtol=1e-4, atol=1e-5 * scale)


@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('shape,N', [((2, 300, 128), 96), ((5000, 128), 256), ((3, 7, 11, 64), 128)])
def test_linear_function_vs_torch(shape, N, relu):
    from unibev_b200 import ops
    g = torch.Generator('cuda').manual_seed(N)
    x = torch.randn(*shape, device='cuda', generator=g)
    lin = torch.nn.Linear(shape[-1], N).cuda()
    go = torch.randn(*shape[:-1], N, device='cuda', generator=g)
    xa = x.clone().requires_grad_()
    ya = ops.linear_train(lin, xa, relu=relu)
    assert type(ya.grad_fn).__name__ == 'LinearFunctionBackward'
    ya.backward(go)
    got = (xa.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    lin.zero_grad()
    xb = x.clone().requires_grad_()
    yb = torch.relu(lin(xb)) if relu else lin(xb)
    yb.backward(go)
    torch.testing.assert_close(ya.detach(), yb.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(got[0], xb.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(got[1], lin.weight.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got[2], lin.bias.grad, rtol=1e-4, atol=1e-4)


def _keep_from_mask(mask, shape):
    bits = (mask[:, None] >> torch.arange(4, device=mask.device, dtype=torch.uint8)) & 1
    return bits.reshape(shape).bool()


@pytest.mark.parametrize('rows,C,p', [(4000, 128, 0.1), (513, 256, 0.25), (64, 96, 0.5), (20000, 128, 0.0)])
def test_dropout_add_layernorm_function(rows, C, p):
    """y = LayerNorm(dropout(x) + residual): the arithmetic in both directions against torch for the mask the kernel drew,
    the keep rate of that mask, and that masks change with the step and the call site but not on their own."""
    from unibev_b200 import ops
    g = torch.Generator('cuda').manual_seed(rows + C)
    x = torch.randn(rows, C, device='cuda', generator=g) * 1.5
    res = torch.randn(rows, C, device='cuda', generator=g)
    gamma = torch.randn(C, device='cuda', generator=g)
    beta = torch.randn(C, device='cuda', generator=g)
    go = torch.randn(rows, C, device='cuda', generator=g)
    rng = ops.DropoutRNG.get(x.device)
    rng.advance()
    xa, ra, ga, ba = (t.clone().requires_grad_() for t in (x, res, gamma, beta))
    ya = ops.DropoutAddLayerNormFunction.apply(xa, ra, ga, ba, 1e-5, p)
    mask = ya.grad_fn.saved_tensors[3]
    keep = _keep_from_mask(mask, (rows, C))
    thr = round(p * 65536)
    p_eff = thr / 65536
    n = rows * C
    assert abs(float(keep.float().mean()) - (1 - p_eff)) < 5 * (p_eff * (1 - p_eff) / n) ** 0.5 + 1e-12
    if p > 0:       # neighbouring elements / rows are not correlated in any obvious way
        k = keep.float() - (1 - p_eff)
        assert abs(float((k[:, 1:] * k[:, :-1]).mean())) < 5 * p_eff * (1 - p_eff) / n ** 0.5
        assert abs(float((k[1:] * k[:-1]).mean())) < 5 * p_eff * (1 - p_eff) / n ** 0.5
    ya.backward(go)
    xb, rb, gb, bb = (t.double().clone().requires_grad_() for t in (x, res, gamma, beta))
    yb = F.layer_norm(xb * keep.double() / (1 - p_eff) + rb, (C,), gb, bb, 1e-5)
    yb.backward(go.double())
    torch.testing.assert_close(ya.detach().double(), yb.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(xa.grad.double(), xb.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ra.grad.double(), rb.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ga.grad.double(), gb.grad, rtol=1e-4, atol=1e-5 * rows ** 0.5)
    torch.testing.assert_close(ba.grad.double(), bb.grad, rtol=1e-4, atol=1e-5 * rows ** 0.5)
    if p > 0:
        def draw():
            return ops.DropoutAddLayerNormFunction.apply(x, res, gamma, beta, 1e-5, p)
        rng.advance()
        y1 = draw()                       # step s, call site 1
        rng.site = 0
        y1_again = draw()                 # the same (step, site): the same mask
        y2 = draw()                       # call site 2
        rng.advance()
        y3 = draw()                       # step s + 1, call site 1
        assert torch.equal(y1, y1_again) and not torch.equal(y1, y2) and not torch.equal(y1, y3)
