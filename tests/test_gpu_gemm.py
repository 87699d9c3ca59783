"""GPU parity of ub_linear_tf32 (tcgen05 GEMM with fused epilogues) against torch.

Two kinds of checks: (1) inputs that are exactly representable in TF32 (multiples of 1/8 in a small range), so
the tensor-core product must equal the fp32 reference up to accumulation order -- this pins the operand layouts,
descriptors and every epilogue exactly; (2) Gaussian inputs with the stated TF32 tolerance (inputs are reduced to
10 mantissa bits: |err| <~ 2^-10 * sqrt(K) * |a||w|)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from unibev_b200 import ops as _ops
    return _ops


def _exact(shape, g, scale=8):
    return torch.randint(-16, 17, shape, generator=g).float() / scale


@pytest.mark.parametrize('M,N,K', [(128, 256, 256), (1000, 256, 256), (40000, 256, 256), (333, 96, 256),
                                   (700, 192, 256), (513, 512, 256), (260, 256, 512), (129, 128, 128), (64, 32, 32)])
@pytest.mark.parametrize('relu', [False, True])
def test_linear_plain_exact(ops, M, N, K, relu):
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = _exact((M, K), g), _exact((N, K), g), _exact((N,), g)
    want = F.linear(x.double(), w.double(), b.double())
    want = (want.relu() if relu else want).float()
    got = ops.linear_tf32(x.cuda(), w.cuda(), b.cuda(), relu=relu).cpu()
    torch.testing.assert_close(got, want, rtol=0, atol=1e-4)


def test_linear_residual_and_strided_out(ops):
    g = torch.Generator().manual_seed(3)
    M, N, K = 777, 96, 256
    x, w, b, r = _exact((M, K), g), _exact((N, K), g), _exact((N,), g), _exact((M, N), g)
    want = (F.linear(x.double(), w.double(), b.double()) + r.double()).float()
    got = ops.linear_tf32(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda()).cpu()
    torch.testing.assert_close(got, want, rtol=0, atol=1e-4)
    got2 = ops.linear_tf32(x.cuda(), w.cuda(), None).cpu()
    torch.testing.assert_close(got2, F.linear(x, w), rtol=0, atol=1e-4)


@pytest.mark.parametrize('M,N,K', [(1000, 256, 256), (40000, 256, 512), (200, 128, 128), (131, 32, 64)])
def test_linear_layernorm_exact_inputs(ops, M, N, K):
    g = torch.Generator().manual_seed(N + K)
    x, w, b, r = _exact((M, K), g), _exact((N, K), g, 64), _exact((N,), g), _exact((M, N), g)
    gam, bet = torch.randn(N, generator=g), torch.randn(N, generator=g)
    pre = F.linear(x.double(), w.double(), b.double()) + r.double()
    want = F.layer_norm(pre, (N,), gam.double(), bet.double(), 1e-5).float()
    got = ops.linear_tf32(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda(), ln=(gam.cuda(), bet.cuda(), 1e-5)).cpu()
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('G,Nv', [(1, 1000), (6, 1450), (2, 333)])
def test_linear_half_planes(ops, G, Nv):
    g = torch.Generator().manual_seed(Nv)
    N, K = 256, 256
    x, w, b = _exact((G * Nv, K), g), _exact((N, K), g, 64), _exact((N,), g)
    want = F.linear(x, w, b).view(G, Nv, 8, 32).permute(0, 2, 1, 3).half()
    got = ops.linear_tf32(x.cuda(), w.cuda(), b.cuda(), planes_nv=Nv).cpu()
    assert got.shape == (G, 8, Nv, 32)
    torch.testing.assert_close(got.float(), want.float(), rtol=1e-3, atol=1e-3)
    same = ops.value_to_half(F.linear(x, w, b).cuda(), G, Nv, 8).cpu()
    assert (got.float() - same.float()).abs().max() <= 1e-3 * want.float().abs().max()


def test_linear_tf32_tolerance_gaussian(ops):
    g = torch.Generator().manual_seed(0)
    M, N, K = 4096, 256, 256
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    want = F.linear(x.double(), w.double(), b.double()).float()
    got = ops.linear_tf32(x.cuda(), w.cuda(), b.cuda()).cpu()
    err = (got - want).abs()
    assert float(err.max()) < 8e-3 and float(err.mean()) < 1.5e-3, (float(err.max()), float(err.mean()))


def test_linear_rejects_uncovered_shapes(ops):
    from unibev_b200 import _cabi
    with pytest.raises(_cabi.UnsupportedShape):
        ops.linear_tf32(torch.zeros(10, 30, device='cuda'), torch.zeros(32, 30, device='cuda'))
    with pytest.raises(_cabi.UnsupportedShape):
        ops.linear_tf32(torch.zeros(10, 32, device='cuda'), torch.zeros(40, 32, device='cuda'))


@pytest.mark.parametrize('cs', [1, 2, 4])
def test_linear_cluster_sizes_agree(ops, cs):
    """The cluster size (TMA multicast of the weight tile) is a performance knob: results must not depend on it."""
    from unibev_b200 import _cabi
    g = torch.Generator().manual_seed(11)
    _cabi.check(_cabi.lib().ub_set_gemm_cluster(cs), 'ub_set_gemm_cluster')
    try:
        for M, N, K in ((40000, 256, 256), (1000, 96, 256), (650, 512, 256), (129, 256, 512), (5000, 192, 256)):
            x, w, b, r = _exact((M, K), g), _exact((N, K), g, 64), _exact((N,), g), _exact((M, N), g)
            want = F.linear(x.double(), w.double(), b.double()).float()
            got = ops.linear_tf32(x.cuda(), w.cuda(), b.cuda()).cpu()
            torch.testing.assert_close(got, want, rtol=0, atol=1e-4)
            if N <= 256:
                gam, bet = torch.randn(N, generator=g), torch.randn(N, generator=g)
                want = F.layer_norm(F.linear(x.double(), w.double(), b.double()) + r.double(), (N,), gam.double(),
                                    bet.double(), 1e-5).float()
                got = ops.linear_tf32(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda(),
                                      ln=(gam.cuda(), bet.cuda(), 1e-5)).cpu()
                torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    finally:
        _cabi.lib().ub_set_gemm_cluster(4)


# ---- fp16-operand variant (resident weight tile, optional fp16 copy of the result) ----------------------------

@pytest.mark.parametrize('M,N,K', [(1000, 256, 256), (40000, 256, 256), (333, 96, 256), (700, 192, 256),
                                   (260, 256, 512), (129, 128, 128), (5000, 512, 256), (64, 32, 64)])
def test_linear_f16_plain_exact(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = _exact((M, K), g), _exact((N, K), g), _exact((N,), g)
    want = F.linear(x.double(), w.double(), b.double()).relu().float()
    got, got16 = ops.linear_f16(x.half().cuda(), w.half().cuda(), b.cuda(), relu=True, f16_out=True)
    torch.testing.assert_close(got.cpu(), want, rtol=0, atol=1e-4)
    torch.testing.assert_close(got16.cpu().float(), want.half().float(), rtol=1e-3, atol=1e-3)
    only16 = ops.linear_f16(x.half().cuda(), w.half().cuda(), b.cuda(), relu=True, fp32_out=False, f16_out=True)
    assert only16[0] is None and torch.equal(only16[1].cpu(), got16.cpu())


@pytest.mark.parametrize('M,N,K', [(1000, 256, 256), (40000, 256, 512), (200, 128, 128)])
def test_linear_f16_layernorm(ops, M, N, K):
    g = torch.Generator().manual_seed(N + K + 1)
    x, w, b, r = _exact((M, K), g), _exact((N, K), g, 64), _exact((N,), g), _exact((M, N), g)
    gam, bet = torch.randn(N, generator=g), torch.randn(N, generator=g)
    pre = F.linear(x.double(), w.double(), b.double()) + r.double()
    want = F.layer_norm(pre, (N,), gam.double(), bet.double(), 1e-5).float()
    got, got16 = ops.linear_f16(x.half().cuda(), w.half().cuda(), b.cuda(), residual=r.cuda(),
                                ln=(gam.cuda(), bet.cuda(), 1e-5), f16_out=True)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got16.cpu().float(), want, rtol=2e-3, atol=2e-3)


def test_linear_f16_planes_and_strided_outputs(ops):
    g = torch.Generator().manual_seed(5)
    G, Nv, N, K = 3, 450, 256, 256
    x, w, b = _exact((G * Nv, K), g), _exact((N, K), g, 64), _exact((N,), g)
    want = F.linear(x, w, b).view(G, Nv, 8, 32).permute(0, 2, 1, 3).half()
    got = ops.linear_f16(x.half().cuda(), w.half().cuda(), b.cuda(), planes_nv=Nv).cpu()
    torch.testing.assert_close(got.float(), want.float(), rtol=1e-3, atol=1e-3)
    # two column halves of a wider output (how the 512-wide FFN hidden layer is produced)
    w2, b2 = _exact((512, K), g, 64), _exact((512,), g)
    buf = torch.empty(G * Nv, 512, device='cuda', dtype=torch.float16)
    for h in range(2):
        ops.linear_f16(x.half().cuda(), w2[h * 256:(h + 1) * 256].half().cuda(), b2[h * 256:(h + 1) * 256].cuda(),
                       relu=True, fp32_out=False, out16=buf[:, h * 256:(h + 1) * 256])
    torch.testing.assert_close(buf.cpu().float(), F.linear(x, w2, b2).relu().half().float(), rtol=1e-3, atol=1e-3)


def test_linear_f16_strided_residual_and_output_views(ops):
    """The self-attention offset|logit rows: residual = one column block of the per-frame positional buffer (row stride
    wider than N), output written into a column block of another buffer -- both ride TMA maps with the caller's stride."""
    g = torch.Generator().manual_seed(11)
    M, N, K = 3000, 96, 256
    x, w, b = _exact((M, K), g), _exact((N, K), g, 64), _exact((N,), g)
    pos_all = _exact((M, 6 * N), g).cuda()                   # six layers' blocks side by side
    out_all = torch.zeros(M, 2 * N, device='cuda')
    for blk in (0, 4):
        res = pos_all[:, blk * N:(blk + 1) * N]
        assert not res.is_contiguous()
        got, _ = ops.linear_f16(x.half().cuda(), w.half().cuda(), b.cuda(), residual=res, out=out_all[:, N:])
        want = (F.linear(x.double(), w.double(), b.double()) + res.cpu().double()).float()
        torch.testing.assert_close(got.cpu(), want, rtol=0, atol=1e-4)
        assert float(out_all[:, :N].abs().max()) == 0.0       # the neighbouring block is untouched


# ---------------------------------------------------------------------------------------------------------------------
# ub_linear_tf32x3: fp32-grade products from three TF32 MMAs.  Gaussian (NOT TF32-representable) operands, float64
# reference, fp32-class tolerance: |err| <= a few 2^-21 * sqrt(K) * |a||w| -- two orders below a single TF32 pass.
def _x3(ops, x, w, b=None, **kw):
    return ops.linear_tf32x3(x.cuda(), ops.split_tf32(w.cuda()), b.cuda() if b is not None else None, **kw).cpu()


def test_split_tf32_is_exact_hi_plus_lo(ops):
    g = torch.Generator().manual_seed(11)
    w = torch.randn(300, 256, generator=g) * torch.logspace(-6, 3, 256)
    hi, lo = (t.cpu() for t in ops.split_tf32(w.cuda()))
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0          # hi is what kind::tf32 reads of w
    assert int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0          # lo is a TF32 value too
    rel = ((hi.double() + lo.double() - w.double()).abs() / w.double().abs()).max()
    assert float(rel) < 2.0 ** -21


@pytest.mark.parametrize('M,N,K', [(128, 256, 256), (1000, 256, 256), (40000, 256, 256), (333, 96, 256),
                                   (700, 192, 256), (513, 512, 256), (260, 256, 512), (129, 128, 128), (64, 32, 32)])
@pytest.mark.parametrize('relu', [False, True])
def test_linear_x3_gaussian(ops, M, N, K, relu):
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    want = F.linear(x.double(), w.double(), b.double())
    want = (want.relu() if relu else want).float()
    got = _x3(ops, x, w, b, relu=relu)
    err = (got - want).abs()
    # measured on B200: mean 1.7e-6, max 1.1e-5 at K = 256 on O(1) outputs -- the floor is the tensor core's fp32
    # accumulation (3 K / 8 truncating adds per output), ~3x an FFMA GEMM's error and 60x below one TF32 pass
    assert float(err.max()) < 4e-5 and float(err.mean()) < 5e-6, (float(err.max()), float(err.mean()))
    one_pass = (ops.linear_tf32(x.cuda(), w.cuda(), b.cuda(), relu=relu).cpu() - want).abs()
    assert float(err.mean()) < float(one_pass.mean()) / 20


def test_linear_x3_large_dynamic_range(ops):
    """8-bit exponent: no saturation / flush where fp16 operands would fail (|x| up to 1e6, weights down to 1e-6)."""
    g = torch.Generator().manual_seed(4)
    M, N, K = 777, 256, 256
    x = torch.randn(M, K, generator=g) * torch.logspace(-3, 6, K)
    w = torch.randn(N, K, generator=g) * torch.logspace(-6, 0, N)[:, None]
    want = F.linear(x.double(), w.double()).float()
    got = _x3(ops, x, w)
    scale = F.linear(x.double().abs(), w.double().abs()).float()            # sum |a||w| per output
    assert float(((got - want).abs() / scale).max()) < 5e-6


def test_linear_x3_residual_strided_and_layernorm(ops):
    g = torch.Generator().manual_seed(3)
    M, N, K = 1777, 256, 512
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    gam, bet = torch.randn(N, generator=g), torch.randn(N, generator=g)
    pre = F.linear(x.double(), w.double(), b.double()) + r.double()
    got = _x3(ops, x, w, b, residual=r.cuda())
    torch.testing.assert_close(got, pre.float(), rtol=1e-5, atol=5e-5)
    want = F.layer_norm(pre, (N,), gam.double(), bet.double(), 1e-5).float()
    got = _x3(ops, x, w, b, residual=r.cuda(), ln=(gam.cuda(), bet.cuda(), 1e-5))
    torch.testing.assert_close(got, want, rtol=1e-5, atol=5e-5)
    # strided output / residual views (column blocks of wider matrices)
    wide_o, wide_r = torch.zeros(M, 3 * N).cuda(), torch.randn(M, 2 * N, generator=g).cuda()
    ops.linear_tf32x3(x.cuda(), ops.split_tf32(w.cuda()), b.cuda(), residual=wide_r[:, N:], out=wide_o[:, N:2 * N])
    want = (F.linear(x.double(), w.double(), b.double()) + wide_r[:, N:].cpu().double()).float()
    torch.testing.assert_close(wide_o[:, N:2 * N].cpu(), want, rtol=1e-5, atol=5e-5)
    assert float(wide_o[:, :N].abs().max()) == 0 and float(wide_o[:, 2 * N:].abs().max()) == 0


@pytest.mark.parametrize('G,Nv', [(1, 1000), (6, 1450), (2, 333)])
def test_linear_x3_fp32_half_head_planes(ops, G, Nv):
    g = torch.Generator().manual_seed(Nv)
    N, K = 256, 256
    x, w, b = torch.randn(G * Nv, K, generator=g), torch.randn(N, K, generator=g) / 16, torch.randn(N, generator=g)
    want = F.linear(x.double(), w.double(), b.double()).float().view(G, Nv, N // 16, 16).permute(0, 2, 1, 3)
    got = _x3(ops, x, w, b, planes_nv=Nv)
    assert got.shape == (G, N // 16, Nv, 16)
    torch.testing.assert_close(got, want.contiguous(), rtol=1e-5, atol=5e-5)


@pytest.mark.parametrize('M,N,K', [(100, 24, 32), (333, 48, 40), (1000, 96, 256), (65, 7, 3)])
def test_linear_simt_any_shape(ops, M, N, K):
    g = torch.Generator().manual_seed(M)
    x, w, b, r = (torch.randn(s, generator=g) for s in ((M, K), (N, K), (N,), (M, N)))
    want = (F.linear(x.double(), w.double(), b.double()) + r.double()).relu().float()
    got = ops.linear_simt(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda(), relu=True).cpu()
    torch.testing.assert_close(got, want, rtol=1e-5, atol=3e-5)      # fp32 FFMA accumulation over K
    torch.testing.assert_close(ops.linear_simt(x.cuda(), w.cuda()).cpu(), F.linear(x, w), rtol=1e-5, atol=3e-5)


def test_unsupported_shapes_are_counted(ops):
    from unibev_b200 import _cabi
    _cabi.reset_launch_count()
    x, w = torch.randn(64, 40).cuda(), torch.randn(24, 40).cuda()
    with pytest.raises(_cabi.UnsupportedShape):
        ops.linear_tf32x3(x, ops.split_tf32(w))
    assert _cabi.unsupported_count() == 1
    _cabi.reset_launch_count()
    assert _cabi.unsupported_count() == 0


# ---------------------------------------------------------------------------------------------------------------------
# ub_linear_f16x3: the same three-product scheme on kind::f16 MMAs (operands bounded below the fp16 range by the caller)
def _h3(ops, x, w, b=None, a_scale=64.0, **kw):
    return ops.linear_f16x3(x.cuda(), ops.split_f16(w.cuda(), a_scale), b.cuda() if b is not None else None, **kw).cpu()


@pytest.mark.parametrize('M,N,K', [(128, 256, 256), (1000, 256, 256), (40000, 256, 256), (333, 96, 256),
                                   (700, 192, 256), (513, 512, 256), (260, 256, 512), (129, 128, 128), (64, 32, 64)])
@pytest.mark.parametrize('relu', [False, True])
def test_linear_f16x3_gaussian(ops, M, N, K, relu):
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    want = F.linear(x.double(), w.double(), b.double())
    want = (want.relu() if relu else want).float()
    got = _h3(ops, x, w, b, relu=relu)
    err = (got - want).abs()
    # the same precision class as the 3xTF32 mode (2 x 11 significand bits per operand), half as many accumulation steps
    assert float(err.max()) < 4e-5 and float(err.mean()) < 5e-6, (float(err.max()), float(err.mean()))
    x3 = (_x3(ops, x, w, b, relu=relu) - want).abs()
    assert float(err.mean()) < 2 * float(x3.mean()) + 1e-7


def test_linear_f16x3_operand_range(ops):
    """Activations from 1e-4 to ~4e3 (a_scale = 8 keeps them inside the fp16 range the caller must guarantee) and weight rows
    from 1e-4 to 1 (row-wise power-of-two scaling): the error stays at the fp32 level relative to sum |a||w| of each output."""
    g = torch.Generator().manual_seed(4)
    M, N, K = 777, 256, 256
    x = torch.randn(M, K, generator=g) * torch.logspace(-4, 3, K)
    w = torch.randn(N, K, generator=g) * torch.logspace(-4, 0, N)[:, None]
    want = F.linear(x.double(), w.double()).float()
    got = _h3(ops, x, w, a_scale=8.0)
    scale = F.linear(x.double().abs(), w.double().abs()).float()
    assert float(((got - want).abs() / scale).max()) < 5e-6


def test_linear_f16x3_residual_layernorm_planes_scatter(ops):
    g = torch.Generator().manual_seed(3)
    M, N, K = 1777, 256, 512
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    gam, bet = torch.randn(N, generator=g), torch.randn(N, generator=g)
    pre = F.linear(x.double(), w.double(), b.double()) + r.double()
    torch.testing.assert_close(_h3(ops, x, w, b, residual=r.cuda()), pre.float(), rtol=1e-5, atol=5e-5)
    want = F.layer_norm(pre, (N,), gam.double(), bet.double(), 1e-5).float()
    torch.testing.assert_close(_h3(ops, x, w, b, residual=r.cuda(), ln=(gam.cuda(), bet.cuda(), 1e-5)), want, rtol=1e-5, atol=5e-5)
    # strided output view
    wide = torch.zeros(M, 3 * N).cuda()
    ops.linear_f16x3(x.cuda(), ops.split_f16(w.cuda(), 64.0), b.cuda(), out=wide[:, N:2 * N])
    torch.testing.assert_close(wide[:, N:2 * N].cpu(), F.linear(x.double(), w.double(), b.double()).float(), rtol=1e-5, atol=5e-5)
    assert float(wide[:, :N].abs().max()) == 0 and float(wide[:, 2 * N:].abs().max()) == 0
    # fp32 half-head planes
    G, Nv = 3, 500
    xp = torch.randn(G * Nv, 256, generator=g)
    wp = torch.randn(256, 256, generator=g) / 16
    wantp = F.linear(xp.double(), wp.double(), b.double()).float().view(G, Nv, 16, 16).permute(0, 2, 1, 3).contiguous()
    torch.testing.assert_close(_h3(ops, xp, wp, b, planes_nv=Nv), wantp, rtol=1e-5, atol=5e-5)
    # row scatter
    B, Nq, Ncam, Nout = 2, 1000, 6, 192
    xs = torch.randn(B * Nq, 256, generator=g)
    ws, bs = torch.randn(Nout, 256, generator=g) / 16, torch.randn(Nout, generator=g)
    q_dst = torch.full((Nq, Ncam), -1, dtype=torch.int32)
    perm = torch.randperm(Ncam * Nq, generator=g)
    k = 0
    for q in range(Nq):
        c = q % 4
        q_dst[q, :c] = perm[k:k + c].int()
        k += c
    out = torch.full((B, Ncam * Nq, Nout), float('nan')).cuda()
    ops.linear_f16x3(xs.cuda(), ops.split_f16(ws.cuda(), 64.0), bs.cuda(), out=out, scatter=(q_dst.cuda(), Nq))
    wants = F.linear(xs.double(), ws.double(), bs.double()).float().view(B, Nq, Nout)
    got = out.cpu()
    written = torch.zeros(Ncam * Nq, dtype=torch.bool)
    for q in range(Nq):
        for d in q_dst[q].tolist():
            if d < 0:
                break
            written[d] = True
            torch.testing.assert_close(got[:, d], wants[:, q], rtol=1e-5, atol=5e-5)
    assert bool(torch.isnan(got[:, ~written]).all())


def test_linear_f16x3_device_side_bound(ops):
    """ub_flatten_feats_max records the largest magnitude of the rows it writes; ub_linear_f16x3_dyn derives the activation
    scale from that device float (times the caller's proven factors) -- same result as the host-scaled call, at any
    input magnitude the bound covers (1e-3 .. 3e4 here)."""
    g = torch.Generator().manual_seed(8)
    G, C, H, W, N = 3, 256, 9, 11, 192
    for mag in (1e-3, 1.0, 3e4):
        feat = (torch.randn(G, C, H, W, generator=g) * mag / 4).clamp(-mag, mag)
        ea, eb = torch.randn(G, C, generator=g) * mag * 0.01, torch.randn(C, generator=g) * mag * 0.01
        mx = torch.zeros(1).cuda()
        rows = ops.flatten_feats_max(feat.cuda(), mx, ea.cuda(), eb.cuda())
        want_rows = feat.flatten(2).transpose(1, 2) + ea[:, None, :] + eb
        torch.testing.assert_close(rows.cpu(), want_rows, rtol=0, atol=0)
        assert float(mx) == float(want_rows.abs().max())
        w, b = torch.randn(N, C, generator=g) / 16, torch.randn(N, generator=g) * mag
        x = rows.view(-1, C)
        got = ops.linear_f16x3_dyn(x, (mx, 1.0, 0.0), ops.split_f16(w.cuda(), 1.0), b.cuda()).cpu()
        want = F.linear(want_rows.reshape(-1, C).double(), w.double(), b.double()).float()
        scale = F.linear(want_rows.reshape(-1, C).double().abs(), w.double().abs()).float() + b.abs()
        assert float(((got - want).abs() / scale).max()) < 5e-6
        # a looser (but still valid) bound -- what a projection's row sums give -- changes the scale, not the result class
        got2 = ops.linear_f16x3_dyn(x, (mx, 3.7, 0.5 * mag), ops.split_f16(w.cuda(), 1.0), b.cuda()).cpu()
        assert float(((got2 - want).abs() / scale).max()) < 5e-6
