"""Fused inference pipeline of the uniform BEV encoder (eval mode).

Reads the parameters held by a ``UniBEVTransformer`` module tree and runs, per
encoder layer:

    self-attention   V  = x Wv^T + bv                     GEMM  (N = C)
                     QP = (x + pos) [Woff;Watt]^T + b     GEMM  (N = H*P*3), pos part precomputed per frame
                     S  = sample(V, QP)                   fused ref-point/softmax/gather/reduce
                     x  = LN(S Wo^T + bo + x)             GEMM with residual + LayerNorm epilogue
    cross-attention  QP = x [Woff;Watt]^T + b             GEMM  (N = H*P*3)
                     S  = img_sample | bev_sample         (value projected once per layer)
                     x  = LN(S Wo^T + bo + x)
    FFN              x  = LN(relu(x W1^T + b1) W2^T + b2 + x)

No sampling_locations / attention_weights / rebatch tensors exist, nothing syncs with the host, and every GEMM and
sampling step is a libunibev_b200 kernel (no torch / cuBLAS matmul anywhere in this file).  Two precision classes:

* ``precision='fp32'`` (default): the arithmetic class of the reference (fp32 everywhere, no ``fp16`` key in any
  UniBEV config).  Projections run on the tensor cores with fp32-grade products: every a*w is evaluated as
  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo over hi / lo splits with 2 x 11 significand bits, fp32 accumulation, residual /
  LayerNorm epilogues -- as three kind::tf32 passes (``ub_linear_tf32x3``) in general, and as three kind::f16 passes
  (``ub_linear_f16x3``, twice the tensor-core rate) where the activation operand has a PROVEN magnitude bound derived from
  the weights (LayerNorm outputs and what is computed from them: ``_LayerWeights.bounds``), so that the power-of-two
  scaled operand stays inside the fp16 range.  Sampling in fp32 (fp32 value maps, fp32 bilinear x attention weights,
  expf / true division softmax).  Meets rtol 1e-3 / atol 1e-4 against the oracle.
* ``precision='fp16'`` (opt-in fast class): fp16 tensor-core operands (activations, weights and input tokens are
  rounded to an 11-bit significand and must stay below 65504), fp16-staged value maps and sampling weights, fp32
  accumulation / residual stream / LayerNorm.  max |err| ~3e-3 on O(1) outputs (tests use atol 5e-3).

``gemm=`` / ``sampling=`` override one half of a class (ablation: tools/ablation.py):
gemm in {'tf32x3', 'tf32', 'f16'}, sampling in {'fp32', 'win16'}.

Reference lines reproduced: transformer_fusion.py:231-278,463-538;
encoder_unibev_detr_img.py:189-289,413-479; encoder_unibev_detr_pts.py:129-209;
spatial_cross_attention_img.py:141-215,381-419; spatial_cross_attention_pts.py:159-206.
"""
import os

import numpy as np
import torch

from .. import _cabi, ops
from .attention import (MSDeformableAttention3DImg, MSDeformableAttention3DPts, MultiScaleDeformableAttention,
                        SpatialCrossAttentionImg, SpatialCrossAttentionPts)
from .encoder import anchor_heights

_ORDER = ('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')
PRECISIONS = {'fp32': ('tf32x3', 'fp32'), 'fp16': ('f16', 'win16')}     # class -> (gemm, sampling)


def _layer_ok(layer, cross_cls, inner_cls):
    if tuple(layer.operation_order) != _ORDER or len(layer.attentions) != 2:
        return False
    sa, ca = layer.attentions
    if type(sa) is not MultiScaleDeformableAttention or type(ca) is not cross_cls:
        return False
    da = ca.deformable_attention
    if not isinstance(da, inner_cls):
        return False
    head_dims_ok = all(m.embed_dims // m.num_heads in (8, 16, 32, 64) and m.embed_dims % m.num_heads == 0
                       for m in (sa, da))
    return (sa.num_levels == 1 and da.num_levels == 1 and sa.batch_first and da.batch_first
            and sa.num_points <= 16 and da.num_points <= 16 and head_dims_ok)


def fused_supported(model, img_feats, pts_feats):
    """The fused pipeline covers the shapes every shipped UniBEV config uses: one feature level,
    (self_attn, norm, cross_attn, norm, ffn, norm) layers.  Anything else takes the module path."""
    if img_feats is not None:
        if len(img_feats) != 1 or not all(_layer_ok(l, SpatialCrossAttentionImg, MSDeformableAttention3DImg)
                                          for l in model.img_bev_encoder.layers):
            return False
        if model.img_bev_encoder.return_intermediate:
            return False
    if pts_feats is not None:
        if len(pts_feats) != 1 or not all(_layer_ok(l, SpatialCrossAttentionPts, MSDeformableAttention3DPts)
                                          for l in model.pts_bev_encoder.layers):
            return False
        if model.pts_bev_encoder.return_intermediate:
            return False
    return True


class _LayerWeights:
    """Per-layer weights in the layout the fused pipeline wants (concatenated offset|logit linears)."""

    def __init__(self, layer):
        sa, ca = layer.attentions
        da = ca.deformable_attention
        f = layer.ffns[0]
        cat = torch.cat
        self.H_s, self.P_s = sa.num_heads, sa.num_points
        self.H_c, self.P_c = da.num_heads, da.num_points
        self.sa_wv, self.sa_bv = sa.value_proj.weight.detach(), sa.value_proj.bias.detach()
        self.sa_wq = cat((sa.sampling_offsets.weight, sa.attention_weights.weight), 0).detach().contiguous()
        self.sa_bq = cat((sa.sampling_offsets.bias, sa.attention_weights.bias), 0).detach().contiguous()
        self.sa_wo, self.sa_bo = sa.output_proj.weight.detach(), sa.output_proj.bias.detach()
        self.ca_wv, self.ca_bv = da.value_proj.weight.detach(), da.value_proj.bias.detach()
        self.ca_wq = cat((da.sampling_offsets.weight, da.attention_weights.weight), 0).detach().contiguous()
        self.ca_bq = cat((da.sampling_offsets.bias, da.attention_weights.bias), 0).detach().contiguous()
        self.ca_wo, self.ca_bo = ca.output_proj.weight.detach(), ca.output_proj.bias.detach()
        self.w1, self.b1 = f.layers[0][0].weight.detach(), f.layers[0][0].bias.detach()
        self.w2, self.b2 = f.layers[1].weight.detach(), f.layers[1].bias.detach()
        self.ln = [(n.weight.detach(), n.bias.detach(), n.eps) for n in layer.norms]
        self._half = None
        self._bounds = None

    def bounds(self, x_in):
        """Proven magnitude bounds of the activations this layer feeds to its projections, given the bound ``x_in`` of the
        layer input (None = unknown): what lets a projection run as fp16 x3 (``ub_linear_f16x3``) instead of 3xTF32.
        LayerNorm output: |y_i| <= sqrt(C - 1) |gamma_i| + |beta_i|; a projection: |x W^T + b| <= |x|_max max_n sum_k |W_nk| +
        |b|_max; sampled rows are convex combinations of value rows (softmax x bilinear weights sum to at most 1); ReLU
        shrinks.  One host read per weight version (the values are cached with the derived weights)."""
        if self._bounds is None or self._bounds[0] != x_in:
            C = self.ln[0][0].numel()

            def ln_bound(k):
                g, b, _ = self.ln[k]
                return float(((C - 1) ** 0.5 * g.abs() + b.abs()).max())

            def through(bound, w, b):
                return None if bound is None else bound * float(w.abs().sum(1).max()) + float(b.abs().max())
            b1, b2 = ln_bound(0), ln_bound(1)
            # cross-attention rows are sampled from value rows tokens Wv^T + bv: |.| <= |tokens|_max ca_mul + ca_add, with the
            # largest token only known on the device (FusedEncoder._tokens)
            self._bounds = (x_in, dict(x=x_in, sa_s=through(x_in, self.sa_wv, self.sa_bv), x1=b1, x2=b2,
                                       hid=through(b2, self.w1, self.b1), out=ln_bound(2),
                                       ca_mul=float(self.ca_wv.abs().sum(1).max()), ca_add=float(self.ca_bv.abs().max())))
        return self._bounds[1]

    def half(self):
        """fp16 copies of the projection weights read with fp16 operands (built once)."""
        if self._half is None:
            self._half = {k: getattr(self, k).half().contiguous()
                          for k in ('sa_wq', 'sa_wv', 'sa_wo', 'ca_wq', 'ca_wv', 'ca_wo', 'w1', 'w2')}
        return self._half


def weights_signature(model):
    """Identity + version of every parameter the fused pipeline derives copies from (concatenated / split / fp16
    weights): an in-place update (optimizer step, load_state_dict) bumps ``_version``, a ``.to()`` / ``.cuda()`` or a
    re-assignment changes ``data_ptr``.  Compared on every call; ~100 integers."""
    return tuple((p.data_ptr(), p._version) for p in model.parameters())


class FusedEncoder:
    def __init__(self, model, precision='fp32', gemm=None, sampling=None):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)} (got {precision!r}); 'fp32' is the class of "
                             "the reference, 'fp16' the opt-in fast class")
        self.m = model
        self.precision = precision
        self.gemm, self.sampling = PRECISIONS[precision]
        if gemm is not None:
            if gemm not in ('tf32x3', 'tf32', 'f16'):
                raise ValueError(f'unknown gemm class {gemm!r}')
            self.gemm = gemm
        if sampling is not None:
            if sampling not in ('fp32', 'win16'):
                raise ValueError(f'unknown sampling class {sampling!r}')
            self.sampling = sampling
        self.f16 = self.gemm == 'f16'
        # sampled rows leave the window kernels as fp16 (the A operand of the fp16 output projection)
        # fp32 sampling through the window-staged fp32 kernels where the shape is covered (UB_WIN32=0: tile kernels only)
        self.win32 = os.environ.get('UB_WIN32', '1') == '1'
        # fp32-grade projections whose activation operand has a proven bound run as fp16 x3 (UB_F16X3=0: 3xTF32 everywhere)
        self.f16x3 = os.environ.get('UB_F16X3', '1') == '1'
        self.half_samples = self.f16 and self.sampling == 'win16'
        self._signature = None
        self._qbound = None
        self._dyn, self._pos_dyn = {}, None
        self._drop_caches()

    # derived weight copies ---------------------------------------------------------------------------------
    def _drop_caches(self):
        self._w, self._split, self._split16, self._rn, self._pos_w = {}, {}, {}, {}, {}

    def refresh(self):
        """Drop every derived weight copy if any source parameter changed since they were built.  Returns True when
        the caches were dropped (a CUDA graph captured over the old copies is stale then)."""
        sig = weights_signature(self.m)
        if sig != self._signature:
            stale = self._signature is not None
            self._signature = sig
            self._drop_caches()
            return stale
        return False

    def _weights(self, name):
        if name not in self._w:
            self._w[name] = [_LayerWeights(l) for l in getattr(self.m, name).layers]
        return self._w[name]

    def _hi_lo(self, w):
        """(w_hi, w_lo) of the 3xTF32 projection (once per weight)."""
        key = (w.data_ptr(), tuple(w.shape))
        if key not in self._split:
            self._split[key] = ops.split_tf32(w.contiguous())
        return self._split[key]

    def _param_bound(self, t):
        if not isinstance(t, torch.nn.Parameter):
            return None
        key = (t.data_ptr(), t._version, tuple(t.shape))
        if self._qbound is None or self._qbound[0] != key:
            self._qbound = (key, float(t.detach().abs().max()))
        return self._qbound[1]

    @staticmethod
    def _a_scale(bound):
        """Power of two that brings activations bounded by ``bound`` just below 2^15 (half the fp16 range), or None when the
        operand cannot go through the fp16 split."""
        if bound is None or not (0.0 < bound < 3.0e7):
            return None
        import math
        return 2.0 ** max(-10, min(10, math.floor(math.log2(32768.0 / bound))))

    def _hi_lo16(self, w, a_scale):
        """(w16_hi, w16_lo, col_scale, a_scale) of the fp16 x3 projection (once per weight and activation scale)."""
        key = (w.data_ptr(), tuple(w.shape), a_scale)
        if key not in self._split16:
            self._split16[key] = ops.split_f16(w.contiguous(), a_scale)
        return self._split16[key]

    def _tf32(self, w):
        """Weight rounded to nearest TF32 (once): tcgen05 kind::tf32 truncates fp32 operands, which would bias every
        product low; a pre-rounded weight is read exactly."""
        key = (w.data_ptr(), tuple(w.shape))
        if key not in self._rn:
            bits = w.contiguous().view(torch.int32)
            self._rn[key] = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
        return self._rn[key]

    # dense projections ------------------------------------------------------------------------------
    # An activation travels as a pair (fp32 rows, fp16 copy or None).  In the fp16 class the LayerNorm epilogues emit
    # the fp16 copy next to the fp32 rows, and the projections read it as their A operand.
    def _lin(self, x, w, b, residual=None, relu=False, ln=None, out=None, w16=None, want16=False, only16=False, bound=None,
             dyn=None):
        """epilogue(x @ w^T) -> (fp32 rows or None, fp16 copy or None).  x: fp32 rows or a (fp32, fp16) pair.  ``bound``: a
        proven bound of |x| (``_LayerWeights.bounds``): the fp32-grade projection then runs as fp16 x3 instead of 3xTF32."""
        x32, x16 = x if isinstance(x, tuple) else (x, None)
        try:
            if self.gemm == 'tf32x3' and self.f16x3 and w.shape[1] % 64 == 0 and dyn is not None and not relu:
                # the operand's bound lives on the device: (one-float tensor, mul, add) states |x| <= t * mul + add
                try:
                    return ops.linear_f16x3_dyn(self._rows32(x), dyn, self._hi_lo16(w, 1.0), b, residual=residual, ln=ln,
                                                out=out), None
                except _cabi.UnsupportedShape:
                    pass
            if self.gemm == 'tf32x3' and self.f16x3 and w.shape[1] % 64 == 0 and self._a_scale(bound) is not None:
                try:
                    return ops.linear_f16x3(self._rows32(x), self._hi_lo16(w, self._a_scale(bound)), b, residual=residual,
                                            relu=relu, ln=ln, out=out), None
                except _cabi.UnsupportedShape:
                    pass
            if self.gemm == 'tf32x3':
                return ops.linear_tf32x3(self._rows32(x), self._hi_lo(w), b, residual=residual, relu=relu, ln=ln, out=out), None
            if self.f16 and x16 is not None and w16 is not None:
                N = w16.shape[0]
                if only16 and N == 512 and residual is None and ln is None and out is None:
                    # two column halves, each with its weight tile resident in shared memory (a 512-row W does
                    # not fit and would be re-streamed from L2 for every row tile)
                    o16 = torch.empty(x16.shape[0], N, device=x16.device, dtype=torch.float16)
                    for h0 in (0, 256):
                        ops.linear_f16(x16, w16[h0:h0 + 256], b[h0:h0 + 256], relu=relu, fp32_out=False,
                                       out16=o16[:, h0:h0 + 256])
                    return None, o16
                return ops.linear_f16(x16, w16, b, residual=residual, relu=relu, ln=ln, out=out,
                                      fp32_out=not only16, f16_out=want16 or only16)
            if self.gemm == 'tf32' and x32 is not None and not only16:      # ablation only: one TF32 pass
                o16 = torch.empty(x32.shape[0], w.shape[0], device=x32.device, dtype=torch.float16) if want16 else None
                return ops.linear_tf32(x32, self._tf32(w), b, residual=residual, relu=relu, ln=ln, out=out,
                                       out16=o16), o16
            if self.f16:        # channel counts the fp16 GEMM does not take (C % 64 != 0): fp32-grade products instead
                return ops.linear_tf32x3(self._rows32(x), self._hi_lo(w), b, residual=residual, relu=relu, ln=ln, out=out), None
        except _cabi.UnsupportedShape:
            pass        # counted by the library (ub_unsupported_count); bench.py asserts the benchmarked shapes have none
        # generic fp32 FFMA projection + (for 'norm' steps) one streaming residual / LayerNorm pass
        x32 = self._rows32(x)
        if ln is not None:
            o = ops.linear_simt(x32, w, b, out=out)
            o = ops.add_layernorm(o, ln[0], ln[1], residual=residual, eps=ln[2], out=o)
        else:
            o = ops.linear_simt(x32, w, b, residual=residual, relu=relu, out=out)
        return o, None

    @staticmethod
    def _rows(s, rows, C):
        """sampled (B, Nq, C) -> the (fp32 rows, fp16 rows) pair `_lin` takes"""
        s = s.view(rows, C)
        return (None, s) if s.dtype == torch.float16 else s

    def _project_value(self, x, w, b, G, Nv, H, P, w16=None, planes32=False, bound=None, dyn=None):
        """value_proj of the rows x (G*Nv, C) -> (value planes for the window kernels or None, fp32 rows or None).
        Planes are fp16 head-major (G, H, Nv, 32) in the 'win16' sampling class and, with ``planes32``, fp32 half-head
        planes (G, 2H, Nv, 16) in the 'fp32' class; they come straight out of the projection's epilogue."""
        x32, x16 = x if isinstance(x, tuple) else (x, None)
        C = w.shape[0]
        if planes32 and self.win32 and self.sampling == 'fp32' and ops.window_supported(C // H, P):
            try:
                if self.f16x3 and w.shape[1] % 64 == 0 and dyn is not None:
                    return ops.linear_f16x3_dyn(self._rows32(x), dyn, self._hi_lo16(w, 1.0), b, planes_nv=Nv), None
                if self.f16x3 and w.shape[1] % 64 == 0 and self._a_scale(bound) is not None:
                    return ops.linear_f16x3(self._rows32(x), self._hi_lo16(w, self._a_scale(bound)), b, planes_nv=Nv), None
                return ops.linear_tf32x3(self._rows32(x), self._hi_lo(w), b, planes_nv=Nv), None
            except _cabi.UnsupportedShape:
                pass
        if self.sampling == 'win16' and ops.window_supported(C // H, P):
            try:
                if self.f16 and x16 is not None and w16 is not None:
                    return ops.linear_f16(x16, w16, b, planes_nv=Nv), None
                if self.gemm == 'tf32' and x32 is not None:
                    return ops.linear_tf32(x32, self._tf32(w), b, planes_nv=Nv), None
            except _cabi.UnsupportedShape:
                pass
            rows, _ = self._lin(x, w, b, w16=w16, bound=bound, dyn=dyn)
            return ops.value_to_half(rows, G, Nv, H), rows
        return None, self._lin(x, w, b, w16=w16, bound=bound, dyn=dyn)[0]

    def _bev_sample(self, x, w, b, qp, B, bev_h, bev_w, fh, fw, H, P, w16=None, bound=None, dyn=None):
        """value_proj + BEV-grid sampling: rows x (B*fh*fw, C) un-projected -> sampled (B, Nq, C)."""
        C = w.shape[0]
        planes, rows = self._project_value(x, w, b, B, fh * fw, H, P, w16, planes32=qp.shape[2] % 4 == 0, bound=bound, dyn=dyn)
        if planes is not None and planes.dtype == torch.float32:
            try:
                return ops.bev_sample_win32(planes, qp, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2, workspace=self._counter())
            except _cabi.UnsupportedShape:
                rows = ops.planes32_to_rows(planes)
        elif planes is not None and qp.shape[2] % 4 == 0:
            try:
                half = self.half_samples and w16 is not None      # the output projection takes fp16 operands
                return ops.bev_sample_win(planes, qp, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2, workspace=self._counter(),
                                          out_dtype=torch.float16 if half else torch.float32,
                                          # fp32 rows that feed a plain TF32 projection: round instead of truncating
                                          round_tf32=self.gemm == 'tf32')
            except _cabi.UnsupportedShape:
                pass
        if rows is None:
            rows = self._lin(x, w, b, w16=w16, bound=bound, dyn=dyn)[0]
        return ops.bev_sample(rows.view(B, fh * fw, C), qp, bev_h, bev_w, fh, fw, H, P, 0, H * P * 2)

    def _counter(self):
        """The next 2-int work-counter slot of this call's workspace (zeroed once per call, re-armed by each kernel)."""
        k = self._n_counters
        self._n_counters += 1
        if k >= self._counters.shape[0]:
            return None        # more window launches than slots: the op allocates its own
        return self._counters[k]

    def _pos_rows(self, names, pos, rows):
        """Positional part of every layer's self-attention offset|logit rows, once per frame and for all encoders
        together: -> {encoder name: [per-layer (rows, n_q) views]}.  pos: fp32 rows or a (fp32, fp16) pair."""
        if pos is None:
            return {n: None for n in names}
        blocks = [(n, lw) for n in names for lw in self._weights(n)]
        widths = [lw.sa_wq.shape[0] for _, lw in blocks]
        dev = (pos[1] if isinstance(pos, tuple) and pos[0] is None else (pos[0] if isinstance(pos, tuple) else pos)).device
        buf = torch.empty(rows, sum(widths), device=dev, dtype=torch.float32)
        views = buf.split(widths, dim=1)
        done = False
        if self.f16 and isinstance(pos, tuple) and pos[1] is not None and all(w == widths[0] and w % 32 == 0 for w in widths):
            # fp16 operands, as many layers per GEMM as fit one 256-column tile with its weights resident
            key = tuple(names)
            if key not in self._pos_w:
                self._pos_w[key] = torch.cat([lw.sa_wq for _, lw in blocks], 0).half().contiguous()
            w_all, per = self._pos_w[key], max(1, 256 // widths[0])
            try:
                for b0 in range(0, len(blocks), per):
                    c0, c1 = b0 * widths[0], min(len(blocks), b0 + per) * widths[0]
                    ops.linear_f16(pos[1], w_all[c0:c1], None, out=buf[:, c0:c1])
                done = True
            except _cabi.UnsupportedShape:
                pass
        if not done and self.gemm == 'tf32x3' and self._pos_dyn is not None and all(w == widths[0] and w % 32 == 0 for w in widths):
            # bev_pos is an input: its largest magnitude was found on the device while flattening it
            key = ('h3',) + tuple(names)
            per = max(1, 256 // widths[0])
            if key not in self._pos_w:
                self._pos_w[key] = [ops.split_f16(torch.cat([lw.sa_wq for _, lw in blocks[b0:b0 + per]], 0).contiguous(), 1.0)
                                    for b0 in range(0, len(blocks), per)]
            try:
                for k, b0 in enumerate(range(0, len(blocks), per)):
                    c0, c1 = b0 * widths[0], min(len(blocks), b0 + per) * widths[0]
                    ops.linear_f16x3_dyn(self._rows32(pos), self._pos_dyn, self._pos_w[key][k], None, out=buf[:, c0:c1])
                done = True
            except _cabi.UnsupportedShape:
                pass
        if not done and self.gemm == 'tf32x3' and all(w == widths[0] and w % 32 == 0 for w in widths):
            # as many layers per launch as fit one 256-column tile: the A side of the 3xTF32 projection (TMA, hi / lo
            # split, operand reads) is paid once per launch
            key = ('x3',) + tuple(names)
            per = max(1, 256 // widths[0])
            if key not in self._pos_w:
                self._pos_w[key] = [ops.split_tf32(torch.cat([lw.sa_wq for _, lw in blocks[b0:b0 + per]], 0).contiguous())
                                    for b0 in range(0, len(blocks), per)]
            try:
                for k, b0 in enumerate(range(0, len(blocks), per)):
                    c0, c1 = b0 * widths[0], min(len(blocks), b0 + per) * widths[0]
                    ops.linear_tf32x3(self._rows32(pos), self._pos_w[key][k], None, out=buf[:, c0:c1])
                done = True
            except _cabi.UnsupportedShape:
                pass
        if not done:
            for (_, lw), dst in zip(blocks, views):
                self._lin(pos, lw.sa_wq, None, out=dst)
        out, k = {}, 0
        for n in names:
            nl = len(self._weights(n))
            out[n] = views[k:k + nl]
            k += nl
        return out

    # one BEV encoder ------------------------------------------------------------------------------
    def _use_f16(self, C):
        return self.f16 and C % 64 == 0

    def _run_encoder(self, name, queries, B, pos_q, value_tokens, sample_cross, bev_h, bev_w):
        """queries (Nq, C) the BEV query table (every sample starts from it, transformer_fusion.py:493-498); pos_q: per-layer
        positional offset|logit rows or None; value_tokens: un-projected feature rows, fp32 or a (fp32 | None, fp16)
        pair; sample_cross(lw, value_tokens, x, w16, x_bound) -> sampled (B, Nq, C) (it runs the offset|logit projection of
        the query rows x, whose proven magnitude bound is x_bound, itself)."""
        layers = self._weights(name)
        Nq, C = queries.shape
        f16 = self._use_f16(C)
        if C % 8 == 0:
            x32, x16 = ops.broadcast_rows(queries.detach(), B, fp32=True, fp16=f16)
        else:
            x32 = queries.detach().unsqueeze(0).expand(B, Nq, C).contiguous()
            x16 = None
        x = (x32.view(B * Nq, C), x16.view(B * Nq, C) if f16 else None)
        if f16 and not isinstance(value_tokens, tuple):
            value_tokens = (value_tokens, value_tokens.half())   # fp16 copy once per frame, read by every layer
        # the query table is an input: its bound is its own largest magnitude, read once per version when it is a Parameter
        # (the head's bev_embedding.weight, unibev_head.py:126-133,172-177); a plain tensor stays unbounded (no host sync per call)
        x_bound = self._param_bound(queries) if (self.gemm == 'tf32x3' and self.f16x3) else None
        for i, lw in enumerate(layers):
            h = lw.half() if f16 else None
            bd = (lw.bounds(x_bound) if (self.gemm == 'tf32x3' and self.f16x3) else
                  dict.fromkeys(('x', 'sa_s', 'x1', 'x2', 'hid', 'out', 'ca_mul', 'ca_add')))
            # --- BEV self-attention (mmcv MultiScaleDeformableAttention, value = query, 1 level)
            qp, _ = self._lin(x, lw.sa_wq, lw.sa_bq, residual=pos_q[i] if pos_q is not None else None,
                              w16=h and h['sa_wq'], bound=bd['x'])
            s = self._bev_sample(x, lw.sa_wv, lw.sa_bv, qp.view(B, Nq, -1), B, bev_h, bev_w, bev_h, bev_w, lw.H_s, lw.P_s,
                                 w16=h and h['sa_wv'], bound=bd['x'])
            x = self._lin(self._rows(s, B * Nq, C), lw.sa_wo, lw.sa_bo, residual=x[0], ln=lw.ln[0], want16=f16,
                          w16=h and h['sa_wo'], bound=bd['sa_s'])
            if f16 and x[1] is None:
                x = (x[0], x[0].half())
            # --- spatial cross-attention (query_pos is None for attentions[1])
            s = sample_cross(lw, value_tokens, x, h and h['ca_wq'], bd['x1'])
            tok_max = self._dyn.get(id(value_tokens))      # largest input token (device): sampled rows <= max * ca_mul + ca_add
            x = self._lin(self._rows(s, B * Nq, C), lw.ca_wo, lw.ca_bo, residual=x[0], ln=lw.ln[1], want16=f16,
                          w16=h and h['ca_wo'], dyn=(tok_max, bd['ca_mul'], bd['ca_add']) if tok_max is not None else None)
            if f16 and x[1] is None:
                x = (x[0], x[0].half())
            # --- FFN
            hid = self._lin(x, lw.w1, lw.b1, relu=True, w16=h and h['w1'], only16=f16, bound=bd['x2'])
            x = self._lin(hid, lw.w2, lw.b2, residual=x[0], ln=lw.ln[2], w16=h and h['w2'], want16=f16 and i + 1 < len(layers),
                          bound=bd['hid'])
            x_bound = bd['out']
            if f16 and x[1] is None and i + 1 < len(layers):
                x = (x[0], x[0].half())
        return x[0].view(B, Nq, C)

    def __call__(self, img_feats, pts_feats, bev_queries, bev_h, bev_w, bev_pos, img_metas, lidar2img=None,
                 img_shape=None):
        """``lidar2img`` (B, N, 4, 4) fp32 already on the device and ``img_shape`` (h, w) replace the per-call
        host->device copy of ``img_metas[i]['lidar2img']`` (encoder_unibev_detr_img.py:115-124) when the caller
        stages calibration itself (``unibev_b200.pipeline.FramePipeline``)."""
        m = self.m
        self.refresh()
        ref = (img_feats or pts_feats)[0]
        B, dev = ref.size(0), ref.device
        Nq, C = bev_h * bev_w, m.embed_dims
        if m.dual_queries:
            q_img, q_pts = bev_queries
        else:
            q_img = q_pts = bev_queries
        f16 = self._use_f16(C)
        names = (['img_bev_encoder'] if img_feats is not None else []) + (['pts_bev_encoder'] if pts_feats is not None else [])
        with torch.no_grad(), torch.cuda.device(dev):
            # work counters of this call's window-kernel launches: one memset per call, distinct memory per call
            # (and so per captured CUDA graph)
            n_win = sum(2 * len(self._weights(n)) for n in names)
            ws = torch.zeros(max(n_win, 1) + 2, 2, device=dev, dtype=torch.int32)
            self._counters, self._n_counters = ws[:-2], 0
            # ... and of the device-side operand bounds (largest magnitudes of bev_pos / image tokens / LiDAR tokens)
            maxes = ws[-2:].view(torch.float32).reshape(-1)
            dyn_ok = self.gemm == 'tf32x3' and self.f16x3
            self._dyn, self._pos_dyn = {}, None
            pos = None
            if bev_pos is not None and dyn_ok:
                pos = ops.flatten_feats_max(bev_pos, maxes[0:1]).view(B * Nq, C)
                self._pos_dyn = (maxes[0:1], 1.0, 0.0)
            elif bev_pos is not None:
                pos = ops.flatten_feats(bev_pos, fp32=not f16, fp16=f16)
                pos = tuple(t.view(B * Nq, C) if t is not None else None for t in pos) if f16 else pos.view(B * Nq, C)
            img = pts = None
            pos_q = self._pos_rows(names, pos, B * Nq)
            if img_feats is not None:
                feat = img_feats[0]
                _, N, _, fh, fw = feat.shape
                enc = m.img_bev_encoder
                tokens = self._tokens(feat, m.cams_embeds if m.use_cams_embeds else None, m.img_level_embeds[0], f16,
                                      B * N * fh * fw, C, maxes[1:2] if dyn_ok else None)
                if lidar2img is not None:
                    l2i = lidar2img
                    ih, iw = img_shape
                else:
                    l2i = torch.from_numpy(np.asarray([mt['lidar2img'] for mt in img_metas], dtype=np.float32)).to(dev)
                    ih, iw = img_metas[0]['img_shape'][0][0], img_metas[0]['img_shape'][0][1]
                D = enc.num_points_in_pillar
                zs = anchor_heights(enc.pc_range[5] - enc.pc_range[2], D).tolist()
                ref_cam, mask = ops.project_points(l2i, zs, enc.pc_range, ih, iw, bev_h, bev_w)
                hits, order = [], []

                def cross(lw, tokens, x, wq16, x_bound):
                    if (self.win32 and self.sampling == 'fp32' and self.gemm == 'tf32x3' and D % 2 == 0
                            and ops.window_supported(C // lw.H_c, lw.P_c)):
                        # fp32 window kernel: value planes from the projection's epilogue, offset|logit rows written
                        # in hit-list order by theirs
                        if not hits:
                            hits.append(ops.build_hits(mask))
                        if not order:
                            order.append(ops.hit_order(mask, ref_cam, hits[0]))
                        q_dst = order[0][0]
                        try:
                            tmax = self._dyn.get(id(tokens))
                            if self.f16x3 and tmax is not None and lw.ca_wv.shape[1] % 64 == 0:
                                planes = ops.linear_f16x3_dyn(self._rows32(tokens), (tmax, 1.0, 0.0), self._hi_lo16(lw.ca_wv, 1.0),
                                                              lw.ca_bv, planes_nv=fh * fw)
                            else:
                                planes = ops.linear_tf32x3(self._rows32(tokens), self._hi_lo(lw.ca_wv), lw.ca_bv, planes_nv=fh * fw)
                            qp_hit = torch.empty(B, N * Nq, lw.ca_wq.shape[0], device=dev, dtype=torch.float32)
                            if self.f16x3 and lw.ca_wq.shape[1] % 64 == 0 and self._a_scale(x_bound) is not None:
                                ops.linear_f16x3(self._rows32(x), self._hi_lo16(lw.ca_wq, self._a_scale(x_bound)), lw.ca_bq,
                                                 out=qp_hit, scatter=(q_dst, Nq))
                            else:
                                ops.linear_tf32x3_scatter(self._rows32(x), self._hi_lo(lw.ca_wq), lw.ca_bq, q_dst, Nq, qp_hit)
                            return ops.img_sample_win32(planes, qp_hit, order[0], hits[0], bev_h, bev_w, fh, fw, lw.H_c,
                                                        lw.P_c, 0, lw.H_c * lw.P_c * 2)
                        except _cabi.UnsupportedShape:
                            pass
                    qp = self._lin(x, lw.ca_wq, lw.ca_bq, w16=wq16, bound=x_bound)[0].view(B, Nq, -1)
                    planes, rows = self._project_value(tokens, lw.ca_wv, lw.ca_bv, B * N, fh * fw, lw.H_c, lw.P_c,
                                                       w16=lw.half()['ca_wv'] if isinstance(tokens, tuple) else None)
                    if planes is not None and (fh + 2) * (fw + 2) * 64 <= 150 * 1024 and qp.shape[2] % 4 == 0:
                        if not hits:
                            hits.append(ops.build_hits(mask))
                        try:
                            return ops.img_sample_win(planes.view(B, N, lw.H_c, fh * fw, -1), qp, ref_cam, hits[0], bev_h,
                                                      bev_w, fh, fw, lw.H_c, lw.P_c, 0, lw.H_c * lw.P_c * 2,
                                                      out_dtype=torch.float16 if self.half_samples and
                                                      isinstance(tokens, tuple) else torch.float32)
                        except _cabi.UnsupportedShape:
                            pass
                    if rows is None:
                        rows = self._lin(tokens, lw.ca_wv, lw.ca_bv,
                                         w16=lw.half()['ca_wv'] if isinstance(tokens, tuple) else None)[0]
                    return ops.img_sample(rows.view(B, N, fh * fw, C), qp, ref_cam, mask, bev_h, bev_w, fh, fw,
                                          lw.H_c, lw.P_c, 0, lw.H_c * lw.P_c * 2)
                img = self._run_encoder('img_bev_encoder', q_img, B, pos_q['img_bev_encoder'], tokens, cross, bev_h, bev_w)
            if pts_feats is not None:
                feat = pts_feats[0]
                _, _, fh, fw = feat.shape
                tokens = self._tokens(feat, None, m.pts_level_embeds[0], f16, B * fh * fw, C, maxes[2:3] if dyn_ok else None)

                def cross(lw, tokens, x, wq16, x_bound, fh=fh, fw=fw):
                    qp = self._lin(x, lw.ca_wq, lw.ca_bq, w16=wq16, bound=x_bound)[0].view(B, Nq, -1)
                    tmax = self._dyn.get(id(tokens)) if not isinstance(tokens, tuple) else None
                    return self._bev_sample(tokens, lw.ca_wv, lw.ca_bv, qp, B, bev_h, bev_w, fh, fw, lw.H_c, lw.P_c,
                                            w16=lw.half()['ca_wv'] if isinstance(tokens, tuple) else None,
                                            dyn=(tmax, 1.0, 0.0) if tmax is not None else None)
                pts = self._run_encoder('pts_bev_encoder', q_pts, B, pos_q['pts_bev_encoder'], tokens, cross, bev_h, bev_w)
            return m.fuse(img, pts)

    def _tokens(self, feat, embed_a, embed_b, f16, rows, C, absmax=None):
        """backbone map -> token rows: fp32, or (None, fp16) when every reader takes fp16 operands.  With ``absmax`` (one
        zeroed device float) the flatten kernel also records the largest token magnitude: the device-side operand bound
        that lets the value projection (and, through its row sums, the output projection) run as fp16 x3."""
        if f16:
            _, t16 = ops.flatten_feats(feat, embed_a, embed_b, fp32=False, fp16=True)
            return (None, t16.view(rows, C))
        if absmax is not None:
            t = ops.flatten_feats_max(feat, absmax, embed_a, embed_b).view(rows, C)
            self._dyn[id(t)] = absmax
            return t
        return ops.flatten_feats(feat, embed_a, embed_b).view(rows, C)

    @staticmethod
    def _rows32(x):
        if isinstance(x, tuple):
            return x[0] if x[0] is not None else x[1].float()
        return x
