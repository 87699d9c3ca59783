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
