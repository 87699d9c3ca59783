"""Torch-facing wrappers of the C ABI: tensors in, tensors out, current CUDA stream.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); the
arithmetic happens in libunibev_b200.so.  Every op requires CUDA fp32 tensors and
raises otherwise -- there is no CPU path.
"""
import ctypes
import os

import torch

from . import _cabi

FUSE_MODES = {'linear': 0, 'avg': 1, 'cat': 2}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)   # of the current device: see _call


def _call(fn, t, *args):
    """One C-ABI call on the device of tensor ``t`` (made current for the call: kernel attributes, the current stream and
    cudaMalloc'ed scratch are per device -- mmcv's op wraps its launches in a CUDAGuard the same way) and on that device's
    current stream."""
    if t.device.index == torch.cuda.current_device():
        _cabi.check(getattr(_cabi.lib(), fn)(*args, _stream()), fn)
    else:
        with torch.cuda.device(t.device):
            _cabi.check(getattr(_cabi.lib(), fn)(*args, _stream()), fn)


_CONST = {}


def const_tensor(values, dtype, device):
    """A small constant tensor (shape lists, start indexes) built ONCE per (values, dtype, device): creating it from a
    Python list is a pageable host -> device copy, which stalls the stream and is not allowed while a CUDA graph is being
    captured (``unibev_b200.train.GraphedTrainStep``).  Callers must not write to the result."""
    key = (repr(values), dtype, str(device))
    t = _CONST.get(key)
    if t is None:
        t = _CONST[key] = torch.tensor(values, dtype=dtype, device=device)
    return t


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _need(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f'unibev_b200: `{name}` must be a CUDA tensor (no CPU fallback exists)')
    if t.dtype != dtype:
        raise TypeError(f'unibev_b200: `{name}` must be {dtype}, got {t.dtype}')
    return t if t.is_contiguous() else t.contiguous()


def msda_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    value = _need(value, 'value')
    loc = _need(sampling_locations, 'sampling_locations')
    w = _need(attention_weights, 'attention_weights')
    shapes = _need(spatial_shapes.to(value.device), 'spatial_shapes', torch.int64)
    lsi = _need(level_start_index.to(value.device), 'level_start_index', torch.int64)
    B, Nv, H, D = value.shape
    _, Nq, _, L, P, _ = loc.shape
    if loc.shape != (B, Nq, H, L, P, 2) or w.shape != (B, Nq, H, L, P) or shapes.shape != (L, 2):
        raise ValueError('msda: inconsistent shapes '
                         f'value{tuple(value.shape)} loc{tuple(loc.shape)} w{tuple(w.shape)} shapes{tuple(shapes.shape)}')
    out = torch.empty(B, Nq, H * D, device=value.device, dtype=torch.float32)
    _call('ub_msda_fwd', value, _ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(w), _ptr(out),
                                        B, Nv, H, D, Nq, L, P)
    return out


class MultiScaleDeformableAttnFunction(torch.autograd.Function):
    """Same call signature and gradient tuple as mmcv's op of the same name
    (reference call site spatial_cross_attention_img.py:433-435)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step=64):
        out = msda_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                           attention_weights)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, w = ctx.saved_tensors
        value, loc, w = _need(value, 'value'), _need(loc, 'loc'), _need(w, 'w')
        shapes = _need(shapes.to(value.device), 'spatial_shapes', torch.int64)
        lsi = _need(lsi.to(value.device), 'level_start_index', torch.int64)
        go = _need(grad_output, 'grad_output')
        B, Nv, H, D = value.shape
        _, Nq, _, L, P, _ = loc.shape
        g_value = torch.zeros_like(value)
        g_loc = torch.empty_like(loc)
        g_w = torch.empty_like(w)
        _call('ub_msda_bwd', value, _ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(w), _ptr(go),
                                            _ptr(g_value), _ptr(g_loc), _ptr(g_w), B, Nv, H, D, Nq, L, P)
        return g_value, None, None, g_loc, g_w, None


def project_points(lidar2img, zs, pc_range, img_h, img_w, bev_h, bev_w):
    """lidar2img (B, N, 4, 4) cuda fp32; zs: D floats (host); -> ref_cam (B, Nq, N, D, 2), mask (B, Nq, N) uint8."""
    l2i = _need(lidar2img, 'lidar2img')
    B, N = l2i.shape[:2]
    D = len(zs)
    zs_c = (ctypes.c_float * D)(*[float(z) for z in zs])
    pc_c = (ctypes.c_float * 6)(*[float(v) for v in pc_range])
    Nq = bev_h * bev_w
    ref = torch.empty(B, Nq, N, D, 2, device=l2i.device, dtype=torch.float32)
    mask = torch.empty(B, Nq, N, device=l2i.device, dtype=torch.uint8)
    _call('ub_project_points', l2i, _ptr(l2i), zs_c, pc_c, float(img_h), float(img_w), _ptr(ref), _ptr(mask),
                                              B, N, bev_h, bev_w, D)
    return ref, mask


def bev_sample(value, qproj, bev_h, bev_w, fH, fW, H, P, off_col, logit_col, out=None):
    """value (B, fH*fW, C); qproj (B, Nq, ld) -> (B, Nq, C)."""
    value, qproj = _need(value, 'value'), _need(qproj, 'qproj')
    B, Nv, C = value.shape
    if Nv != fH * fW or qproj.shape[0] != B or qproj.shape[1] != bev_h * bev_w or C % H:
        raise ValueError(f'bev_sample: inconsistent shapes value{tuple(value.shape)} qproj{tuple(qproj.shape)}')
    if out is None:
        out = torch.empty(B, bev_h * bev_w, C, device=value.device, dtype=torch.float32)
    _call('ub_bev_sample_fwd', value, _ptr(value), _ptr(qproj), _ptr(out), B, bev_h, bev_w, fH, fW, H, C // H, P,
                                              qproj.shape[2], off_col, logit_col)
    return out


def img_sample(value, qproj, ref_cam, mask, bev_h, bev_w, fH, fW, H, P, off_col, logit_col, out=None):
    """value (B, N, fH*fW, C); qproj (B, Nq, ld); ref_cam (B, Nq, N, D, 2); mask (B, Nq, N) uint8 -> (B, Nq, C)."""
    value, qproj, ref_cam = _need(value, 'value'), _need(qproj, 'qproj'), _need(ref_cam, 'ref_cam')
    mask = _need(mask, 'mask', torch.uint8)
    B, N, Nv, C = value.shape
    D = ref_cam.shape[3]
    Nq = bev_h * bev_w
    if Nv != fH * fW or qproj.shape[:2] != (B, Nq) or ref_cam.shape != (B, Nq, N, D, 2) or mask.shape != (B, Nq, N) or C % H:
        raise ValueError('img_sample: inconsistent shapes')
    if out is None:
        out = torch.empty(B, Nq, C, device=value.device, dtype=torch.float32)
    _call('ub_img_sample_fwd', value, _ptr(value), _ptr(qproj), _ptr(ref_cam), _ptr(mask), _ptr(out), B, N,
                                              bev_h, bev_w, fH, fW, H, C // H, P, D, qproj.shape[2], off_col, logit_col)
    return out


class BevSampleFunction(torch.autograd.Function):
    """Autograd twin of ``bev_sample``: forward ``ub_bev_sample_fwd``, backward ``ub_bev_sample_bwd`` (gradients with
    respect to the projected value rows and the RAW offset | logit rows; no loc / weight tensors in either direction).
    ``FusedSample.bev(value, qproj, bev_h, bev_w, fH, fW, H, P, off_col, logit_col)``."""

    @staticmethod
    def forward(ctx, value, qproj, bev_h, bev_w, fH, fW, H, P, off_col, logit_col):
        value, qproj = _need(value, 'value'), _need(qproj, 'qproj')
        ctx.save_for_backward(value, qproj)
        ctx.geo = (bev_h, bev_w, fH, fW, H, P, off_col, logit_col)
        return bev_sample(value, qproj, bev_h, bev_w, fH, fW, H, P, off_col, logit_col)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        value, qproj = ctx.saved_tensors
        bev_h, bev_w, fH, fW, H, P, off_col, logit_col = ctx.geo
        go = _need(grad_out, 'grad_out')
        B, _, C = value.shape
        g_value = torch.zeros_like(value)
        g_qproj = torch.empty_like(qproj) if qproj.shape[2] == 3 * H * P else torch.zeros_like(qproj)
        _call('ub_bev_sample_bwd', value, _ptr(value), _ptr(qproj), _ptr(go), _ptr(g_value), _ptr(g_qproj), B, bev_h, bev_w,
              fH, fW, H, C // H, P, qproj.shape[2], off_col, logit_col)
        return (g_value, g_qproj) + (None,) * 8


class ImgSampleFunction(torch.autograd.Function):
    """Autograd twin of ``img_sample`` (camera cross-attention incl. the batch-0 hit quirk and the count division):
    forward ``ub_img_sample_fwd``, backward ``ub_img_sample_bwd``."""

    @staticmethod
    def forward(ctx, value, qproj, ref_cam, mask, bev_h, bev_w, fH, fW, H, P, off_col, logit_col):
        value, qproj, ref_cam = _need(value, 'value'), _need(qproj, 'qproj'), _need(ref_cam, 'ref_cam')
        mask = _need(mask, 'mask', torch.uint8)
        ctx.save_for_backward(value, qproj, ref_cam, mask)
        ctx.geo = (bev_h, bev_w, fH, fW, H, P, off_col, logit_col)
        return img_sample(value, qproj, ref_cam, mask, bev_h, bev_w, fH, fW, H, P, off_col, logit_col)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        value, qproj, ref_cam, mask = ctx.saved_tensors
        bev_h, bev_w, fH, fW, H, P, off_col, logit_col = ctx.geo
        go = _need(grad_out, 'grad_out')
        B, N, _, C = value.shape
        g_value = torch.zeros_like(value)
        g_qproj = torch.empty_like(qproj) if qproj.shape[2] == 3 * H * P else torch.zeros_like(qproj)
        _call('ub_img_sample_bwd', value, _ptr(value), _ptr(qproj), _ptr(ref_cam), _ptr(mask), _ptr(go), _ptr(g_value),
              _ptr(g_qproj), B, N, bev_h, bev_w, fH, fW, H, C // H, P, ref_cam.shape[3], qproj.shape[2], off_col, logit_col)
        return (g_value, g_qproj) + (None,) * 10


def fused_sample_supported(Dh, P):
    """Shapes the generic fused sampling kernels cover in both directions (``ub_bev/img_sample_fwd`` + ``_bwd``)."""
    return Dh in (8, 16, 32, 64) and 0 < P <= 16


def value_to_half(value, G, Nv, H, out=None):
    """value (G*Nv, C) fp32 token-major -> (G, H, Nv, C // H) fp16 head-major planes for the window kernels."""
    value = _need(value, 'value')
    C = value.shape[-1]
    if value.numel() != G * Nv * C or C % H or (C // H) % 8:
        raise ValueError(f'value_to_half: inconsistent shapes value{tuple(value.shape)} G={G} Nv={Nv} H={H}')
    if out is None:
        out = torch.empty(G, H, Nv, C // H, device=value.device, dtype=torch.float16)
    _call('ub_value_to_half', value, _ptr(value), _ptr(out), G, Nv, H, C // H)
    return out


def window_supported(Dh, P):
    """Shapes the window-staged (fp16) sampling kernels are built for."""
    return Dh == 32 and P in (4, 8)


def bev_sample_win(value16, qproj, bev_h, bev_w, fH, fW, H, P, off_col, logit_col, out=None, out_dtype=torch.float32,
                   workspace=None, round_tf32=False):
    """value16 (B, H, fH*fW, 32) fp16 from value_to_half; qproj (B, Nq, ld) -> (B, Nq, H*32) fp32 (or fp16).
    ``workspace``: 2 zeroed int32 owned by the caller (left zero by the call); calls that may overlap in time need
    distinct ones.  Default: a fresh one per call."""
    value16, qproj = _need(value16, 'value16', torch.float16), _need(qproj, 'qproj')
    B, Hh, Nv, Dh = value16.shape
    if Hh != H or Nv != fH * fW or qproj.shape[0] != B or qproj.shape[1] != bev_h * bev_w:
        raise ValueError(f'bev_sample_win: inconsistent shapes value16{tuple(value16.shape)} qproj{tuple(qproj.shape)}')
    if out is None:
        out = torch.empty(B, bev_h * bev_w, H * Dh, device=qproj.device, dtype=out_dtype)
    if workspace is None:
        workspace = torch.zeros(2, device=qproj.device, dtype=torch.int32)
    elif workspace.dtype != torch.int32 or workspace.numel() < 2 or workspace.device != qproj.device:
        raise ValueError('bev_sample_win: `workspace` must be 2 int32 on the device of the inputs')
    _call('ub_bev_sample_win_fwd', value16, _ptr(value16), _ptr(qproj), _ptr(out), int(out.dtype == torch.float16),
          B, bev_h, bev_w, fH, fW, H, Dh, P, qproj.shape[2], off_col, logit_col, _ptr(workspace), int(round_tf32))
    return out


def bev_sample_win32(planes32, qproj, bev_h, bev_w, fH, fW, H, P, off_col, logit_col, out=None, workspace=None):
    """fp32 twin of ``bev_sample_win``: planes32 (B, 2H, fH*fW, 16) fp32 half-head planes (``linear_tf32x3(planes_nv=)``
    or ``value_to_planes32``); qproj (B, Nq, ld) -> (B, Nq, H*32) fp32."""
    planes32, qproj = _need(planes32, 'planes32'), _need(qproj, 'qproj')
    B, H2, Nv, Dh2 = planes32.shape
    if H2 != 2 * H or Dh2 != 16 or Nv != fH * fW or qproj.shape[0] != B or qproj.shape[1] != bev_h * bev_w:
        raise ValueError(f'bev_sample_win32: inconsistent shapes planes32{tuple(planes32.shape)} qproj{tuple(qproj.shape)}')
    if out is None:
        out = torch.empty(B, bev_h * bev_w, H * 32, device=qproj.device, dtype=torch.float32)
    if workspace is None:
        workspace = torch.zeros(2, device=qproj.device, dtype=torch.int32)
    elif workspace.dtype != torch.int32 or workspace.numel() < 2 or workspace.device != qproj.device:
        raise ValueError('bev_sample_win32: `workspace` must be 2 int32 on the device of the inputs')
    _call('ub_bev_sample_win32_fwd', planes32, _ptr(planes32), _ptr(qproj), _ptr(out), B, bev_h, bev_w, fH, fW, H, 32, P,
          qproj.shape[2], off_col, logit_col, _ptr(workspace))
    return out


def value_to_planes32(value, G, Nv, H):
    """value (G*Nv, H*32) fp32 token-major rows -> (G, 2H, Nv, 16) half-head planes (torch glue for tests / ablations;
    the product path gets the planes from the value projection's epilogue)."""
    return value.view(G, Nv, 2 * H, 16).permute(0, 2, 1, 3).contiguous()


def hit_order(mask, ref_cam, hits):
    """-> q_dst (Nq, N) int32 (scatter map of ``linear_tf32x3_scatter``), hit_ref (B, N, Nq, 2 D) fp32, hit_meta (B, N, Nq, 4)
    fp32 records {query index bits, 1 / #cameras, 0, 0}: the hit-list-ordered inputs of ``img_sample_win32``; once per frame."""
    mask, ref_cam = _need(mask, 'mask', torch.uint8), _need(ref_cam, 'ref_cam')
    B, Nq, N = mask.shape
    D = ref_cam.shape[3]
    hit_idx, hit_cnt, inv_cnt = hits[0], hits[1], hits[2]
    q_dst = torch.empty(Nq, N, device=mask.device, dtype=torch.int32)
    hit_ref = torch.empty(B, N, Nq, 2 * D, device=mask.device, dtype=torch.float32)
    hit_meta = torch.empty(B, N, Nq, 4, device=mask.device, dtype=torch.float32)
    _call('ub_hit_order', mask, _ptr(mask), _ptr(ref_cam), _ptr(hit_idx), _ptr(hit_cnt), _ptr(inv_cnt), _ptr(q_dst),
          _ptr(hit_ref), _ptr(hit_meta), B, N, Nq, D)
    return q_dst, hit_ref, hit_meta


def linear_tf32x3_scatter(x, w_split, bias, q_dst, rows_per_item, out):
    """``linear_tf32x3`` whose row b * rows_per_item + q is written to the rows b * out.shape[1] + q_dst[q, j] (j up to the
    first negative entry) of ``out`` (B, rows, N): the offset|logit rows in hit-list order."""
    x = _need(x, 'x')
    w_hi, w_lo = _need(w_split[0], 'w_hi'), _need(w_split[1], 'w_lo')
    q_dst = _need(q_dst, 'q_dst', torch.int32)
    M, K = x.shape
    N = w_hi.shape[0]
    bias = _need(bias, 'bias') if bias is not None else None
    if (w_hi.shape != (N, K) or out.dim() != 3 or out.shape[2] != N or not out.is_contiguous() or out.dtype != torch.float32
            or M % rows_per_item or out.shape[0] != M // rows_per_item or q_dst.shape[0] != rows_per_item):
        raise ValueError('linear_tf32x3_scatter: inconsistent shapes')
    _call('ub_linear_tf32x3_scatter', x, _ptr(x), _ptr(w_hi), _ptr(w_lo), _ptr(bias), _ptr(out), N, _ptr(q_dst), q_dst.shape[1],
          rows_per_item, out.shape[1], M, N, K)
    return out


def img_sample_win32(planes32, qp_hit, order, hits, bev_h, bev_w, fH, fW, H, P, off_col, logit_col, out=None):
    """fp32 twin of ``img_sample_win``: planes32 (B*N, 2H, fH*fW, 16); qp_hit (B, N*Nq, ld) offset|logit rows in hit-list
    order; order = hit_order(mask, ref_cam, hits); hits = build_hits(mask) -> (B, Nq, H*32) fp32."""
    planes32, qp_hit = _need(planes32, 'planes32'), _need(qp_hit, 'qp_hit')
    hit_ref, hit_meta = _need(order[1], 'hit_ref'), _need(order[2], 'hit_meta')
    hit_idx, hit_cnt = hits[0], hits[1]
    B, N, Nq, D2 = hit_ref.shape
    if (planes32.shape != (B * N, 2 * H, fH * fW, 16) or Nq != bev_h * bev_w or qp_hit.shape[:2] != (B, N * Nq)
            or hit_idx.shape != (N + 1, Nq) or hit_meta.shape != (B, N, Nq, 4)):
        raise ValueError('img_sample_win32: inconsistent shapes')
    if out is None:
        out = torch.empty(B, Nq, H * 32, device=qp_hit.device, dtype=torch.float32)
    _call('ub_img_sample_win32_fwd', planes32, _ptr(planes32), _ptr(qp_hit), _ptr(hit_ref), _ptr(hit_meta), _ptr(hit_idx),
          _ptr(hit_cnt), _ptr(out), B, N, bev_h, bev_w, fH, fW, H, 32, P, D2 // 2, qp_hit.shape[2], off_col, logit_col)
    return out


def planes32_to_rows(planes32):
    """(G, 2H, Nv, 16) half-head planes -> (G*Nv, H*32) token-major rows (torch glue; inverse of value_to_planes32)."""
    G, H2, Nv, _ = planes32.shape
    return planes32.permute(0, 2, 1, 3).reshape(G * Nv, H2 * 16)


def build_hits(mask):
    """mask (B, Nq, N) uint8 -> hit_idx (N + 1, Nq) int32 (rank-split lists, row N = unseen queries), hit_cnt (2N + 1)
    int32 (first counts, later counts, unseen count), inv_cnt (B, Nq) fp32, hit_ic (B, N, Nq) fp32 = inv_cnt in hit-list
    order (all on the device; see ub_build_hits)."""
    mask = _need(mask, 'mask', torch.uint8)
    B, Nq, N = mask.shape
    hit_idx = torch.empty(N + 1, Nq, device=mask.device, dtype=torch.int32)
    hit_cnt = torch.empty(2 * N + 1, device=mask.device, dtype=torch.int32)
    inv_cnt = torch.empty(B, Nq, device=mask.device, dtype=torch.float32)
    hit_ic = torch.empty(B, N, Nq, device=mask.device, dtype=torch.float32)      # inv_cnt in hit-list order
    _call('ub_build_hits', mask, _ptr(mask), _ptr(hit_idx), _ptr(hit_cnt), _ptr(inv_cnt), _ptr(hit_ic), B, N, Nq)
    return hit_idx, hit_cnt, inv_cnt, hit_ic


def img_sample_win(value16, qproj, ref_cam, hits, bev_h, bev_w, fH, fW, H, P, off_col, logit_col, out=None,
                   out_dtype=torch.float32):
    """value16 (B, N, H, fH*fW, 32) fp16; hits = build_hits(mask); -> (B, Nq, H*32) fp32."""
    value16, qproj, ref_cam = _need(value16, 'value16', torch.float16), _need(qproj, 'qproj'), _need(ref_cam, 'ref_cam')
    hit_idx, hit_cnt, inv_cnt = hits[:3]
    hit_ic = hits[3] if len(hits) > 3 else None
    B, N, Hh, Nv, Dh = value16.shape
    D = ref_cam.shape[3]
    Nq = bev_h * bev_w
    if (Hh != H or Nv != fH * fW or qproj.shape[:2] != (B, Nq) or ref_cam.shape != (B, Nq, N, D, 2)
            or hit_idx.shape != (N + 1, Nq) or hit_cnt.shape != (2 * N + 1,) or inv_cnt.shape != (B, Nq)):
        raise ValueError('img_sample_win: inconsistent shapes')
    if out is None:
        out = torch.empty(B, Nq, H * Dh, device=qproj.device, dtype=out_dtype)
    _call('ub_img_sample_win_fwd', value16, _ptr(value16), _ptr(qproj), _ptr(ref_cam), _ptr(hit_idx), _ptr(hit_cnt),
                                                  _ptr(inv_cnt), _ptr(hit_ic), _ptr(out), int(out.dtype == torch.float16), B, N,
                                                  bev_h,
                                                  bev_w, fH, fW, H, Dh, P, D,
                                                  qproj.shape[2], off_col, logit_col)
    return out


def linear_tf32(x, weight, bias=None, residual=None, relu=False, ln=None, out=None, planes_nv=None, out16=None):
    """x (M, K) @ weight (N, K)^T on the tcgen05 tensor cores (TF32) with the epilogue fused:
    + bias, + residual (M, N), ReLU, LayerNorm (ln = (gamma, beta, eps)); ``planes_nv`` = rows per value map:
    return fp16 head-major planes (M // planes_nv, N // 32, planes_nv, 32) instead of an fp32 (M, N) matrix."""
    x, weight = _need(x, 'x'), _need(weight, 'weight')
    M, K = x.shape
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise ValueError(f'linear_tf32: x{tuple(x.shape)} vs weight{tuple(weight.shape)}')
    bias = _need(bias, 'bias') if bias is not None else None
    flags = (1 if relu else 0) | (2 if ln is not None else 0)
    gamma = beta = None
    eps = 0.0
    if ln is not None:
        gamma, beta, eps = _need(ln[0], 'gamma'), _need(ln[1], 'beta'), float(ln[2])
    ldr = 0
    if residual is not None:
        if residual.shape != (M, N):
            raise ValueError('linear_tf32: residual shape mismatch')
        if not (residual.is_cuda and residual.dtype == torch.float32 and residual.stride(1) == 1):
            residual = _need(residual, 'residual')
        ldr = residual.stride(0)
    planes = None
    if planes_nv is not None:
        if M % planes_nv or N % 32:
            raise ValueError('linear_tf32: planes need M % Nv == 0 and N % 32 == 0')
        planes = out if out is not None else torch.empty(M // planes_nv, N // 32, planes_nv, 32, device=x.device,
                                                         dtype=torch.float16)
        res = planes
        out_ptr, ldc = None, 0
    else:
        if out is None:
            out = torch.empty(M, N, device=x.device, dtype=torch.float32)
        elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError('linear_tf32: `out` must be an fp32 CUDA (M, N) matrix with unit column stride')
        res = out
        out_ptr, ldc = _ptr(out), out.stride(0)
    if out16 is not None:
        if planes is not None or out16.shape != (M, N) or out16.dtype != torch.float16 or out16.stride(1) != 1:
            raise ValueError('linear_tf32: `out16` must be an fp16 CUDA (M, N) matrix (not with planes)')
        _call('ub_linear_tf32_dual', x, _ptr(x), _ptr(weight), _ptr(bias), _ptr(residual), ldr, _ptr(gamma),
                                                    _ptr(beta), eps, out_ptr, ldc, _ptr(out16), out16.stride(0), M, N, K,
                                                    flags)
        return res
    _call('ub_linear_tf32', x, _ptr(x), _ptr(weight), _ptr(bias), _ptr(residual), ldr, _ptr(gamma), _ptr(beta),
                                           eps, out_ptr, ldc, _ptr(planes), planes_nv or 0, M, N, K, flags)
    return res


def linear_simt(x, weight, bias=None, residual=None, relu=False, out=None):
    """Generic fp32 (FFMA) projection ``x (M, K) @ weight (N, K)^T`` for shapes the tensor-core entry points reject."""
    x, weight = _need(x, 'x'), _need(weight, 'weight')
    M, K = x.shape
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise ValueError(f'linear_simt: x{tuple(x.shape)} vs weight{tuple(weight.shape)}')
    bias = _need(bias, 'bias') if bias is not None else None
    ldr = 0
    if residual is not None:
        if residual.shape != (M, N):
            raise ValueError('linear_simt: residual shape mismatch')
        if not (residual.is_cuda and residual.dtype == torch.float32 and residual.stride(1) == 1):
            residual = _need(residual, 'residual')
        ldr = residual.stride(0)
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32 or not out.is_cuda:
        raise ValueError('linear_simt: `out` must be an fp32 CUDA (M, N) matrix with unit column stride')
    _call('ub_linear_simt', x, _ptr(x), _ptr(weight), _ptr(bias), _ptr(residual), ldr, _ptr(out),
                                           out.stride(0), M, N, K, int(relu))
    return out


def split_tf32(weight):
    """weight fp32 -> (hi, lo): hi = what a TF32 MMA reads of weight (13 low mantissa bits cleared), lo = tf32(weight - hi)."""
    weight = _need(weight, 'weight')
    hi, lo = torch.empty_like(weight), torch.empty_like(weight)
    _call('ub_split_tf32', weight, _ptr(weight), _ptr(hi), _ptr(lo), weight.numel())
    return hi, lo


def linear_tf32x3(x, w_split, bias=None, residual=None, relu=False, ln=None, out=None, planes_nv=None):
    """fp32-grade ``x (M, K) @ W (N, K)^T`` on the tcgen05 tensor cores (three TF32 MMAs per product, see
    ``ub_linear_tf32x3``); ``w_split = split_tf32(W)``.  Same fused epilogues as ``linear_tf32``; ``planes_nv``: return fp32
    half-head planes (M // planes_nv, N // 16, planes_nv, 16) for the fp32 window-staged sampling kernels."""
    x = _need(x, 'x')
    w_hi, w_lo = _need(w_split[0], 'w_hi'), _need(w_split[1], 'w_lo')
    M, K = x.shape
    N = w_hi.shape[0]
    if w_hi.shape != (N, K) or w_lo.shape != (N, K):
        raise ValueError(f'linear_tf32x3: x{tuple(x.shape)} vs weight{tuple(w_hi.shape)}')
    bias = _need(bias, 'bias') if bias is not None else None
    flags = (1 if relu else 0) | (2 if ln is not None else 0)
    gamma = beta = None
    eps = 0.0
    if ln is not None:
        gamma, beta, eps = _need(ln[0], 'gamma'), _need(ln[1], 'beta'), float(ln[2])
    ldr = 0
    if residual is not None:
        if residual.shape != (M, N):
            raise ValueError('linear_tf32x3: residual shape mismatch')
        if not (residual.is_cuda and residual.dtype == torch.float32 and residual.stride(1) == 1):
            residual = _need(residual, 'residual')
        ldr = residual.stride(0)
    planes = None
    if planes_nv is not None:
        if M % planes_nv or N % 16:
            raise ValueError('linear_tf32x3: planes need M % Nv == 0 and N % 16 == 0')
        planes = out if out is not None else torch.empty(M // planes_nv, N // 16, planes_nv, 16, device=x.device,
                                                         dtype=torch.float32)
        res, out_ptr, ldc = planes, None, 0
    else:
        if out is None:
            out = torch.empty(M, N, device=x.device, dtype=torch.float32)
        elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError('linear_tf32x3: `out` must be an fp32 CUDA (M, N) matrix with unit column stride')
        res, out_ptr, ldc = out, _ptr(out), out.stride(0)
    _call('ub_linear_tf32x3', x, _ptr(x), _ptr(w_hi), _ptr(w_lo), _ptr(bias), _ptr(residual), ldr, _ptr(gamma),
                                             _ptr(beta), eps, out_ptr, ldc, _ptr(planes), planes_nv or 0, M, N, K, flags)
    return res


def split_f16(weight, a_scale=1.0):
    """weight (N, K) fp32 -> (hi16, lo16, col_scale, a_scale): fp16 hi / lo parts of the row-scaled weight and the epilogue
    factors 1 / (s_n a_scale) -- the weight operand of ``linear_f16x3`` for activations that are multiplied by ``a_scale``
    (a power of two chosen so that |activation| * a_scale stays below the fp16 range) before their own split."""
    weight = _need(weight, 'weight')
    if weight.dim() != 2:
        raise ValueError('split_f16: weight must be (N, K)')
    hi = torch.empty(weight.shape, device=weight.device, dtype=torch.float16)
    lo = torch.empty(weight.shape, device=weight.device, dtype=torch.float16)
    cs = torch.empty(weight.shape[0], device=weight.device, dtype=torch.float32)
    _call('ub_split_f16', weight, _ptr(weight), _ptr(hi), _ptr(lo), _ptr(cs), weight.shape[0], weight.shape[1], float(a_scale))
    return hi, lo, cs, float(a_scale)


def linear_f16x3(x, w16_split, bias=None, residual=None, relu=False, ln=None, out=None, planes_nv=None, scatter=None):
    """``linear_tf32x3`` on fp16 tensor-core passes (``ub_linear_f16x3``): same precision class, twice the rate -- ONLY for
    operands the caller has bounded below the fp16 range.  ``w16_split = split_f16(W)``.  ``scatter=(q_dst, rows_per_item)``
    with a 3-D ``out`` (B, rows, N): rows leave in hit-list order as in ``linear_tf32x3_scatter``."""
    x = _need(x, 'x')
    w_hi, w_lo = _need(w16_split[0], 'w16_hi', torch.float16), _need(w16_split[1], 'w16_lo', torch.float16)
    col_scale, a_scale = _need(w16_split[2], 'col_scale'), float(w16_split[3])
    M, K = x.shape
    N = w_hi.shape[0]
    if w_hi.shape != (N, K) or w_lo.shape != (N, K):
        raise ValueError(f'linear_f16x3: x{tuple(x.shape)} vs weight{tuple(w_hi.shape)}')
    bias = _need(bias, 'bias') if bias is not None else None
    flags = (1 if relu else 0) | (2 if ln is not None else 0)
    gamma = beta = None
    eps = 0.0
    if ln is not None:
        gamma, beta, eps = _need(ln[0], 'gamma'), _need(ln[1], 'beta'), float(ln[2])
    ldr = 0
    if residual is not None:
        if residual.shape != (M, N):
            raise ValueError('linear_f16x3: residual shape mismatch')
        if not (residual.is_cuda and residual.dtype == torch.float32 and residual.stride(1) == 1):
            residual = _need(residual, 'residual')
        ldr = residual.stride(0)
    planes, sc, sc_r, sc_rows, sc_dst = None, None, 0, 0, 0
    if planes_nv is not None:
        if M % planes_nv or N % 16:
            raise ValueError('linear_f16x3: planes need M % Nv == 0 and N % 16 == 0')
        planes = out if out is not None else torch.empty(M // planes_nv, N // 16, planes_nv, 16, device=x.device,
                                                         dtype=torch.float32)
        res, out_ptr, ldc = planes, None, 0
    elif scatter is not None:
        sc, sc_rows = _need(scatter[0], 'q_dst', torch.int32), int(scatter[1])
        if (out is None or out.dim() != 3 or out.shape[2] != N or not out.is_contiguous() or out.dtype != torch.float32
                or M % sc_rows or out.shape[0] != M // sc_rows or sc.shape[0] != sc_rows):
            raise ValueError('linear_f16x3: scatter needs a contiguous fp32 `out` (B, rows, N) and q_dst (rows_per_item, R)')
        sc_r, sc_dst = sc.shape[1], out.shape[1]
        res, out_ptr, ldc = out, _ptr(out), N
    else:
        if out is None:
            out = torch.empty(M, N, device=x.device, dtype=torch.float32)
        elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError('linear_f16x3: `out` must be an fp32 CUDA (M, N) matrix with unit column stride')
        res, out_ptr, ldc = out, _ptr(out), out.stride(0)
    _call('ub_linear_f16x3', x, _ptr(x), a_scale, _ptr(w_hi), _ptr(w_lo), _ptr(col_scale), _ptr(bias), _ptr(residual), ldr,
          _ptr(gamma), _ptr(beta), eps,
          out_ptr, ldc, _ptr(planes), planes_nv or 0, _ptr(sc), sc_r, sc_rows, sc_dst, M, N, K, flags)
    return res


def linear_f16x3_dyn(x, bound, w16_split, bias=None, residual=None, ln=None, out=None, planes_nv=None):
    """``linear_f16x3`` whose activation bound lives on the device: ``bound = (tensor with one float, mul, add)`` states
    |x| <= tensor[0] * mul + add; ``w16_split = split_f16(W, 1.0)``."""
    x = _need(x, 'x')
    bt, mul, add = _need(bound[0], 'bound'), float(bound[1]), float(bound[2])
    w_hi, w_lo = _need(w16_split[0], 'w16_hi', torch.float16), _need(w16_split[1], 'w16_lo', torch.float16)
    col_scale = _need(w16_split[2], 'col_scale')
    if float(w16_split[3]) != 1.0:
        raise ValueError('linear_f16x3_dyn: the weight split must be made with a_scale = 1')
    M, K = x.shape
    N = w_hi.shape[0]
    if w_hi.shape != (N, K) or w_lo.shape != (N, K):
        raise ValueError(f'linear_f16x3_dyn: x{tuple(x.shape)} vs weight{tuple(w_hi.shape)}')
    bias = _need(bias, 'bias') if bias is not None else None
    flags = 2 if ln is not None else 0
    gamma = beta = None
    eps = 0.0
    if ln is not None:
        gamma, beta, eps = _need(ln[0], 'gamma'), _need(ln[1], 'beta'), float(ln[2])
    ldr = 0
    if residual is not None:
        if residual.shape != (M, N):
            raise ValueError('linear_f16x3_dyn: residual shape mismatch')
        if not (residual.is_cuda and residual.dtype == torch.float32 and residual.stride(1) == 1):
            residual = _need(residual, 'residual')
        ldr = residual.stride(0)
    planes = None
    if planes_nv is not None:
        if M % planes_nv or N % 16:
            raise ValueError('linear_f16x3_dyn: planes need M % Nv == 0 and N % 16 == 0')
        planes = out if out is not None else torch.empty(M // planes_nv, N // 16, planes_nv, 16, device=x.device,
                                                         dtype=torch.float32)
        res, out_ptr, ldc = planes, None, 0
    else:
        if out is None:
            out = torch.empty(M, N, device=x.device, dtype=torch.float32)
        elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError('linear_f16x3_dyn: `out` must be an fp32 CUDA (M, N) matrix with unit column stride')
        res, out_ptr, ldc = out, _ptr(out), out.stride(0)
    _call('ub_linear_f16x3_dyn', x, _ptr(x), _ptr(bt), mul, add, _ptr(w_hi), _ptr(w_lo), _ptr(col_scale), _ptr(bias),
          _ptr(residual), ldr, _ptr(gamma), _ptr(beta), eps, out_ptr, ldc, _ptr(planes), planes_nv or 0, M, N, K, flags)
    return res


def linear_f16(x16, w16, bias=None, residual=None, relu=False, ln=None, out=None, out16=None, fp32_out=True,
               f16_out=False, planes_nv=None):
    """fp16-operand variant of ``linear_tf32``: x16 (M, K) fp16 @ w16 (N, K)^T fp16, fp32 accumulation and epilogue.
    Returns ``planes`` (with ``planes_nv``) or the pair (fp32 result or None, fp16 copy or None)."""
    x16, w16 = _need(x16, 'x16', torch.float16), _need(w16, 'w16', torch.float16)
    M, K = x16.shape
    N = w16.shape[0]
    if w16.shape[1] != K:
        raise ValueError(f'linear_f16: x{tuple(x16.shape)} vs weight{tuple(w16.shape)}')
    bias = _need(bias, 'bias') if bias is not None else None
    flags = (1 if relu else 0) | (2 if ln is not None else 0)
    gamma = beta = None
    eps = 0.0
    if ln is not None:
        gamma, beta, eps = _need(ln[0], 'gamma'), _need(ln[1], 'beta'), float(ln[2])
    ldr = 0
    if residual is not None:
        if residual.shape != (M, N):
            raise ValueError('linear_f16: residual shape mismatch')
        if not (residual.is_cuda and residual.dtype == torch.float32 and residual.stride(1) == 1):
            residual = _need(residual, 'residual')
        ldr = residual.stride(0)
    planes = None
    if planes_nv is not None:
        if M % planes_nv or N % 32:
            raise ValueError('linear_f16: planes need M % Nv == 0 and N % 32 == 0')
        planes = torch.empty(M // planes_nv, N // 32, planes_nv, 32, device=x16.device, dtype=torch.float16)
        out = out16 = None
    else:
        if out is None and fp32_out:
            out = torch.empty(M, N, device=x16.device, dtype=torch.float32)
        if out16 is None and f16_out:
            out16 = torch.empty(M, N, device=x16.device, dtype=torch.float16)
        for t, dt in ((out, torch.float32), (out16, torch.float16)):
            if t is not None and (t.shape != (M, N) or t.stride(1) != 1 or t.dtype != dt or not t.is_cuda):
                raise ValueError('linear_f16: outputs must be CUDA (M, N) matrices with unit column stride')
    _call('ub_linear_f16', x16, _ptr(x16), _ptr(w16), _ptr(bias), _ptr(residual), ldr, _ptr(gamma), _ptr(beta),
                                          eps, _ptr(out), out.stride(0) if out is not None else 0, _ptr(out16),
                                          out16.stride(0) if out16 is not None else 0, _ptr(planes), planes_nv or 0,
                                          M, N, K, flags)
    return planes if planes is not None else (out, out16)


def add_layernorm(x, gamma, beta, bias=None, residual=None, eps=1e-5, out=None, out16=None):
    x = _need(x, 'x')
    C = x.shape[-1]
    rows = x.numel() // C
    gamma, beta = _need(gamma, 'gamma'), _need(beta, 'beta')
    bias = _need(bias, 'bias') if bias is not None else None
    residual = _need(residual, 'residual') if residual is not None else None
    if residual is not None and residual.shape != x.shape:
        raise ValueError('add_layernorm: residual shape mismatch')
    if out is None:
        out = torch.empty_like(x)
    if out16 is not None:
        if out16.shape != x.shape or out16.dtype != torch.float16 or not out16.is_cuda or not out16.is_contiguous():
            raise ValueError('add_layernorm: `out16` must be a contiguous fp16 CUDA tensor of the shape of x')
        _call('ub_add_layernorm16', x, _ptr(x), _ptr(bias), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(out),
                                                   _ptr(out16), rows, C, float(eps))
        return out
    _call('ub_add_layernorm', x, _ptr(x), _ptr(bias), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(out),
                                             rows, C, float(eps))
    return out


def colsum(x, out=None):
    """x (..., N) -> (N) column sums over all leading dims (``ub_colsum``): the bias gradient of a projection."""
    x = _need(x, 'x')
    N = x.shape[-1]
    if out is None:
        out = torch.zeros(N, device=x.device, dtype=torch.float32)
    _call('ub_colsum', x, _ptr(x), _ptr(out), x.numel() // N, N)
    return out


def train_ops_supported(C):
    """Last-dim sizes ``ub_colsum`` / ``ub_layernorm_bwd`` / ``ub_add_layernorm`` cover."""
    return C % 4 == 0 and 0 < C <= 1024


def _wgrad(g2, x2):
    """g2 (M, N)^T @ x2 (M, K) -> (N, K): the weight gradient of a projection.  M is the number of BEV queries / tokens
    (tens of thousands), N and K a few hundred: one skinny product with a very long reduction, for which cuBLAS picks a
    single-wave kernel (72 us at M = 80 000, N = K = 128).  Splitting the reduction into S batches (one bmm, S x N x K partial
    sums, one small sum) fills the GPU: ~3x faster."""
    M = g2.shape[0]
    for S in (64, 50, 40, 32, 25, 20, 16, 10, 8):
        if M % S == 0 and M // S >= 512:
            part = torch.bmm(g2.view(S, M // S, -1).transpose(1, 2), x2.reshape(S, M // S, -1))
            return part.sum(0)
    return g2.t() @ x2


class LinearFunction(torch.autograd.Function):
    """y = [relu](x W^T + b) with the library's column-sum kernel for the bias gradient (torch's generic column reduction is
    the single largest item of the training step's backward: ~50 us per projection against ~10 us here) and a
    split-reduction weight gradient; the matrix products stay with cuBLAS (TF32 when
    ``torch.backends.cuda.matmul.allow_tf32`` is set, as the reference's stack did; ReLU in its epilogue)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu=False):
        ctx.has_bias = bias is not None
        ctx.relu = bool(relu) and bias is not None
        x2 = x.reshape(-1, x.shape[-1])
        y = x.new_empty(*x.shape[:-1], weight.shape[0])      # (returned as is, not as a view: callers may apply in-place ReLU)
        if ctx.relu:
            torch._addmm_activation(bias, x2, weight.t(), use_gelu=False, out=y.view(-1, weight.shape[0]))
        elif bias is not None:
            torch.addmm(bias, x2, weight.t(), out=y.view(-1, weight.shape[0]))
        else:
            torch.mm(x2, weight.t(), out=y.view(-1, weight.shape[0]))
        ctx.save_for_backward(x, weight, y if ctx.relu else None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if ctx.relu:
            g2 = torch.ops.aten.threshold_backward(g2, y.view(-1, y.shape[-1]), 0)
        g2 = g2 if g2.is_contiguous() else g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (g2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            gw = _wgrad(g2, x.reshape(-1, x.shape[-1]))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = colsum(g2)
        return gx, gw, gb, None


class LayerNormFunction(torch.autograd.Function):
    """y = LayerNorm(x [+ residual]) * gamma + beta: forward ``ub_add_layernorm``, backward ``ub_layernorm_bwd`` (one pass
    each; the sum x + residual is never written: the backward pass rebuilds it from its two terms)."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps):
        x = _need(x, 'x')
        residual = _need(residual, 'residual') if residual is not None else None
        ctx.save_for_backward(x, residual, gamma)
        ctx.eps = float(eps)
        return add_layernorm(x, gamma.detach(), beta.detach(), residual=residual, eps=eps)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, residual, gamma = ctx.saved_tensors
        gy = _need(gy, 'grad_out')
        C = x.shape[-1]
        gx = torch.empty_like(x)
        gg = torch.zeros(2, C, device=x.device, dtype=torch.float32)
        _call('ub_layernorm_bwd', x, _ptr(x), _ptr(residual), _ptr(gy), _ptr(_need(gamma.detach(), 'gamma')), _ptr(gx), _ptr(gg[0]),
              _ptr(gg[1]), x.numel() // C, C, ctx.eps)
        return gx, (gx if residual is not None else None), gg[0], gg[1], None


# UB_FUSED_TRAIN=0 keeps the module (autograd) path on plain torch modules + the op-level ub_msda_fwd / ub_msda_bwd
TRAIN_KERNELS = os.environ.get('UB_FUSED_TRAIN', '1') == '1'


def linear_train(module, x, relu=False):
    """``nn.Linear`` (+ ReLU) forward of the module (autograd) path: ``LinearFunction`` for fp32 CUDA activations, else the
    module."""
    if (TRAIN_KERNELS and x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and module.bias is not None
            and train_ops_supported(module.out_features)):
        return LinearFunction.apply(x, module.weight, module.bias, relu)
    return torch.relu(module(x)) if relu else module(x)


class AddRowVectorFunction(torch.autograd.Function):
    """x (..., C) + v (C) with the library's column sum for the gradient of v (the level / camera embeddings added to every
    token, transformer_fusion.py:241-253,266-270: torch reduces their gradient with its generic column reduction)."""

    @staticmethod
    def forward(ctx, x, v):
        return x + v

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        gc = g if g.is_contiguous() else g.contiguous()
        return g, (colsum(gc) if ctx.needs_input_grad[1] else None)


def add_row_vector(x, v):
    if TRAIN_KERNELS and x.is_cuda and x.dtype == torch.float32 and v.dim() == 1 and train_ops_supported(v.numel()):
        return AddRowVectorFunction.apply(x, v)
    return x + v


class DropoutRNG:
    """Per-device state of the in-kernel dropout generator: a device tensor {seed, step} plus the host-side numbering of
    the call sites of a step.  ``advance()`` (one tiny kernel: also recorded by CUDA-graph capture, so every replay moves on)
    starts a new step; ``UniBEVTransformer.encode`` calls it at the top of every training-mode forward."""
    _by_device = {}

    def __init__(self, device):
        self.state = torch.tensor([torch.initial_seed() & 0x7fffffffffffffff, 0], dtype=torch.int64, device=device)
        self._one = torch.tensor([0, 1], dtype=torch.int64, device=device)
        self.site = 0

    @classmethod
    def get(cls, device):
        key = str(device)
        if key not in cls._by_device:
            cls._by_device[key] = cls(device)
        return cls._by_device[key]

    def advance(self):
        self.state.add_(self._one)
        self.site = 0

    def next_site(self):
        self.site += 1
        return self.site


class DropoutAddLayerNormFunction(torch.autograd.Function):
    """y = LayerNorm(dropout(x, p) + residual) * gamma + beta: ``ub_dropout_add_layernorm_fwd`` / ``_bwd``."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps, p):
        x, residual = _need(x, 'x'), _need(residual, 'residual')
        C = x.shape[-1]
        rng = DropoutRNG.get(x.device)
        mask = torch.empty(x.numel() // 4, dtype=torch.uint8, device=x.device)
        out = torch.empty_like(x)
        _call('ub_dropout_add_layernorm_fwd', x, _ptr(x), _ptr(residual), _ptr(_need(gamma.detach(), 'gamma')),
              _ptr(_need(beta.detach(), 'beta')), _ptr(out), _ptr(mask), x.numel() // C, C, float(eps), float(p), _ptr(rng.state),
              rng.next_site())
        ctx.save_for_backward(x, residual, gamma, mask)
        ctx.eps, ctx.p = float(eps), float(p)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, residual, gamma, mask = ctx.saved_tensors
        gy = _need(gy, 'grad_out')
        C = x.shape[-1]
        gx, gr = torch.empty_like(x), torch.empty_like(x)
        gg = torch.zeros(2, C, device=x.device, dtype=torch.float32)
        _call('ub_dropout_add_layernorm_bwd', x, _ptr(x), _ptr(residual), _ptr(mask), _ptr(gy), _ptr(_need(gamma.detach(), 'gamma')),
              _ptr(gx), _ptr(gr), _ptr(gg[0]), _ptr(gg[1]), x.numel() // C, C, ctx.eps, ctx.p)
        return gx, gr, gg[0], gg[1], None, None


class Deferred:
    """``dropout(out) + identity`` not yet computed: what an attention / FFN of the module path hands to a 'norm' step that
    folds dropout and addition into its kernel (``layer_norm_train``).  ``p`` = drop probability still to be applied to
    ``out`` (0: none).  ``materialize()`` is the plain expression for every other consumer."""
    __slots__ = ('out', 'identity', 'p')

    def __init__(self, out, identity, p=0.0):
        self.out, self.identity, self.p = out, identity, float(p)

    def materialize(self):
        out = torch.nn.functional.dropout(self.out, self.p, True) if self.p > 0.0 else self.out
        return out + self.identity


def add_identity(out, identity, defer, dropout=None):
    """``dropout(out) + identity`` of the attentions / FFN, or the unevaluated triple when the caller's next step is a
    'norm' that drops and adds while normalising."""
    if (defer and TRAIN_KERNELS and out.is_cuda and out.dtype == torch.float32 and out.shape == identity.shape
            and train_ops_supported(out.shape[-1])):
        p = dropout.p if (isinstance(dropout, torch.nn.Dropout) and dropout.training) else 0.0
        if dropout is None or isinstance(dropout, (torch.nn.Dropout, torch.nn.Identity)):
            return Deferred(out, identity, p)
    return (dropout(out) if dropout is not None else out) + identity


def layer_norm_train(module, x):
    """``nn.LayerNorm`` forward of the module (autograd) path through the library's kernels where the shape is covered.
    ``x`` may be a ``Deferred`` dropout + sum."""
    if isinstance(x, Deferred):
        ok = (TRAIN_KERNELS and module.elementwise_affine and len(module.normalized_shape) == 1
              and train_ops_supported(x.out.shape[-1]))
        if ok and x.p > 0.0:
            return DropoutAddLayerNormFunction.apply(x.out, x.identity, module.weight, module.bias, module.eps, x.p)
        if ok:
            return LayerNormFunction.apply(x.out, x.identity, module.weight, module.bias, module.eps)
        return module(x.materialize())
    if (TRAIN_KERNELS and x.is_cuda and x.dtype == torch.float32 and module.elementwise_affine and len(module.normalized_shape) == 1
            and train_ops_supported(x.shape[-1])):
        return LayerNormFunction.apply(x, None, module.weight, module.bias, module.eps)
    return module(x)




def cnw_fuse(img, pts, w_img, w_pts, mode, c_flag, l_flag, s_img=None, s_pts=None, modal_embed=None):
    ref = img if img is not None else pts
    if ref is None:
        raise ValueError('cnw_fuse: both modalities are None')
    img = _need(img, 'img') if img is not None else None
    pts = _need(pts, 'pts') if pts is not None else None
    B, Nq, C = ref.shape
    opt = [(_need(t, n) if t is not None else None) for t, n in
           ((w_img, 'w_img'), (w_pts, 'w_pts'), (s_img, 's_img'), (s_pts, 's_pts'), (modal_embed, 'modal_embed'))]
    out = torch.empty(B, Nq, C * (2 if mode == 'cat' else 1), device=ref.device, dtype=torch.float32)
    _call('ub_cnw_fuse', ref, _ptr(img), _ptr(pts), *[_ptr(t) for t in opt], _ptr(out), B * Nq, Nq, C,
                                        FUSE_MODES[mode], int(c_flag), int(l_flag))
    return out


def flatten_feats(feat, embed_a=None, embed_b=None, fp32=True, fp16=False):
    """feat (..., C, h, w) with G = prod(leading dims) -> (G, h*w, C) + embed_a[g % len(embed_a)] + embed_b.
    Returns the fp32 tensor, or with ``fp16=True`` the pair (fp32 or None, fp16 copy)."""
    feat = _need(feat, 'feat')
    C, h, w = feat.shape[-3:]
    G = feat.numel() // (C * h * w)
    embed_a = _need(embed_a, 'embed_a') if embed_a is not None else None
    embed_b = _need(embed_b, 'embed_b') if embed_b is not None else None
    n_a = embed_a.shape[0] if embed_a is not None else 0
    out = torch.empty(G, h * w, C, device=feat.device, dtype=torch.float32) if fp32 or not fp16 else None
    out16 = torch.empty(G, h * w, C, device=feat.device, dtype=torch.float16) if fp16 else None
    _call('ub_flatten_feats16', feat, _ptr(feat), _ptr(embed_a), n_a, _ptr(embed_b), _ptr(out), _ptr(out16), G, C,
                                               h * w)
    return (out, out16) if fp16 else out


def flatten_feats_max(feat, absmax, embed_a=None, embed_b=None):
    """``flatten_feats`` (fp32 rows) that also raises ``absmax`` (one zeroed fp32 on the device) to the largest magnitude it
    wrote: the device-side operand bound of ``linear_f16x3_dyn``."""
    feat = _need(feat, 'feat')
    absmax = _need(absmax, 'absmax')
    C, h, w = feat.shape[-3:]
    G = feat.numel() // (C * h * w)
    embed_a = _need(embed_a, 'embed_a') if embed_a is not None else None
    embed_b = _need(embed_b, 'embed_b') if embed_b is not None else None
    out = torch.empty(G, h * w, C, device=feat.device, dtype=torch.float32)
    _call('ub_flatten_feats_max', feat, _ptr(feat), _ptr(embed_a), embed_a.shape[0] if embed_a is not None else 0,
          _ptr(embed_b), _ptr(out), _ptr(absmax), G, C, h * w)
    return out


def broadcast_rows(src, B, fp32=True, fp16=True):
    """src (rows, C) -> (fp32 (B, rows, C) or None, fp16 (B, rows, C) or None): the query table repeated per sample."""
    src = _need(src, 'src')
    rows, C = src.shape
    out32 = torch.empty(B, rows, C, device=src.device, dtype=torch.float32) if fp32 else None
    out16 = torch.empty(B, rows, C, device=src.device, dtype=torch.float16) if fp16 else None
    _call('ub_broadcast_rows', src, _ptr(src), rows, C, B, _ptr(out32), _ptr(out16))
    return out32, out16


# ---------------------------------------------------------------------------------------------- [R8] voxelize
_vox_ws = {}


def hard_voxelize(points, voxel_size, pc_range, max_points, max_voxels):
    """points (N, C) fp32 on the device -> (voxels (max_voxels, max_points, C), coors (max_voxels, 3) int32 (z, y, x),
    num_points_per_voxel (max_voxels) int32, voxel_num (1,) int32 ON THE DEVICE).  Rows >= voxel_num are padding.
    Asynchronous: the caller decides when (if ever) to read voxel_num back."""
    points = _need(points, 'points')
    if points.dim() != 2 or points.size(1) < 3:
        raise ValueError(f'hard_voxelize: points must be (N, C>=3), got {tuple(points.shape)}')
    N, C = points.shape
    dev = points.device
    voxels = torch.empty(max_voxels, max_points, C, device=dev, dtype=torch.float32)
    coors = torch.zeros(max_voxels, 3, device=dev, dtype=torch.int32)
    num = torch.empty(max_voxels, device=dev, dtype=torch.int32)
    voxel_num = torch.zeros(1, device=dev, dtype=torch.int32)
    if N == 0:
        voxels.zero_()
        num.zero_()
        return voxels, coors, num, voxel_num
    nbytes = ctypes.c_size_t(0)
    _cabi.check(_cabi.lib().ub_voxelize_workspace_bytes(N, ctypes.byref(nbytes)), 'ub_voxelize_workspace_bytes')
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _vox_ws.get(key)
    if ws is None or ws.numel() < nbytes.value:
        ws = _vox_ws[key] = torch.empty(nbytes.value, device=dev, dtype=torch.uint8)
    vs = (ctypes.c_float * 3)(*[float(v) for v in voxel_size])
    pr = (ctypes.c_float * 6)(*[float(v) for v in pc_range])
    _call('ub_hard_voxelize', points, _ptr(points), N, C, vs, pr, int(max_points), int(max_voxels), _ptr(voxels),
                                             _ptr(coors), _ptr(num), _ptr(voxel_num), _ptr(ws), ws.numel())
    return voxels, coors, num, voxel_num


def voxel_mean(voxels, num_points, num_features):
    """HardSimpleVFE: (M, max_points, C), (M,) int32 -> (M, num_features)."""
    voxels = _need(voxels, 'voxels')
    num_points = _need(num_points, 'num_points', torch.int32)
    M, T, C = voxels.shape
    out = torch.empty(M, num_features, device=voxels.device, dtype=torch.float32)
    _call('ub_voxel_mean', voxels, _ptr(voxels), _ptr(num_points), M, T, C, int(num_features), _ptr(out))
    return out
