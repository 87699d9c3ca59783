"""CPU restatement of UniBEV's object-query decoder.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Functional, parameterised by a state dict with the reference's
key names; follows

* ``DetectionTransformerDecoder.forward`` -- decoder.py:65-128 (layer loop, ``inverse_sigmoid`` refinement :33-48,
  ``return_intermediate`` stacking);
* ``CustomMSDeformableAttention.forward`` -- decoder.py:230-338 (identical arithmetic to mmcv's
  ``MultiScaleDeformableAttention``: ``oracle.mmcv_semantics.mmcv_msda_forward``);
* the un-vendored mmcv-full 1.3.17 ``DetrTransformerDecoderLayer`` / ``BaseTransformerLayer.forward`` operation dispatch
  and ``MultiheadAttention`` wrapper (``nn.MultiheadAttention`` with ``query_pos`` added to query and key, value without;
  ``identity + dropout(out)``), restated from their published behaviour -- parity UNPINNED for that part (the reference
  ships no test for it); the pinned part is the reference's own decoder.py run over those stubs (tests/golden/decoder.npz).
"""
import torch
import torch.nn.functional as F

from . import mmcv_semantics as ms


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def multihead_attention(p, prefix, query, query_pos, num_heads):
    """mmcv MultiheadAttention as the decoder's self-attention: q = k = query + query_pos, v = query; sequence-first."""
    q = query + query_pos if query_pos is not None else query
    out, _ = F.multi_head_attention_forward(
        q, q, query, query.shape[-1], num_heads, p[prefix + '.attn.in_proj_weight'], p[prefix + '.attn.in_proj_bias'],
        None, None, False, 0.0, p[prefix + '.attn.out_proj.weight'], p[prefix + '.attn.out_proj.bias'], training=False,
        need_weights=False)
    return query + out


def decoder_layer(p, prefix, query, value, query_pos, reference_points, spatial_shapes, num_heads, num_points):
    """('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'), sequence-first tensors (num_query, bs, C)."""
    query = multihead_attention(p, prefix + '.attentions.0', query, query_pos, num_heads)
    query = ms.layer_norm(p, prefix + '.norms.0', query)
    query = ms.mmcv_msda_forward(p, prefix + '.attentions.1', query, value=value, query_pos=query_pos,
                                 reference_points=reference_points, spatial_shapes=spatial_shapes, num_heads=num_heads,
                                 num_levels=1, num_points=num_points, batch_first=False)
    query = ms.layer_norm(p, prefix + '.norms.1', query)
    query = ms.ffn_forward(p, prefix + '.ffns.0', query)
    return ms.layer_norm(p, prefix + '.norms.2', query)


def decoder_forward(p, query, value, query_pos, reference_points, bev_hw, num_layers, num_heads=8, num_points=4,
                    reg_branches=None, prefix='layers'):
    """-> (inter_states (L, num_query, bs, C), inter_references (L, bs, num_query, 3)).
    ``reg_branches``: list of callables (one per layer) or None."""
    shapes = [tuple(int(v) for v in bev_hw)]
    output, inter, inter_ref = query, [], []
    for lid in range(num_layers):
        ref_in = reference_points[..., :2].unsqueeze(2)
        output = decoder_layer(p, f'{prefix}.{lid}', output, value, query_pos, ref_in, shapes, num_heads, num_points)
        if reg_branches is not None:
            tmp = reg_branches[lid](output.permute(1, 0, 2))
            new_ref = torch.zeros_like(reference_points)
            new_ref[..., :2] = tmp[..., :2] + inverse_sigmoid(reference_points[..., :2])
            new_ref[..., 2:3] = tmp[..., 4:5] + inverse_sigmoid(reference_points[..., 2:3])
            reference_points = new_ref.sigmoid()
        inter.append(output)
        inter_ref.append(reference_points)
    return torch.stack(inter), torch.stack(inter_ref)


def reg_branches_from(p, num_layers, prefix='reg'):
    """The seeded Linear-ReLU-Linear branches stored with the golden vector."""
    def branch(i):
        def f(x):
            return F.linear(F.relu(F.linear(x, p[f'{prefix}.{i}.0.weight'], p[f'{prefix}.{i}.0.bias'])),
                            p[f'{prefix}.{i}.2.weight'], p[f'{prefix}.{i}.2.bias'])
        return f
    return [branch(i) for i in range(num_layers)]
