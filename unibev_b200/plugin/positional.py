"""Producer of ``bev_pos`` (input side of the hot path).

The reference head builds ``bev_mask = zeros(bs, bev_h, bev_w)`` and ``bev_pos = self.positional_encoding(bev_mask)``
(unibev_head.py:179-182) with mmdet's ``LearnedPositionalEncoding(num_feats, row_num_embed, col_num_embed)``
(config unibev_nus_LC_cnw_256_modality_dropout.py:356-361; un-vendored mmdet 2.19, restated from its published
behaviour: column embedding of x and row embedding of y concatenated on the channel axis, repeated for the batch).
Same constructor keywords and parameter names (``row_embed.weight``, ``col_embed.weight``), same output
``(bs, 2 * num_feats, h, w)``."""
import torch
import torch.nn as nn

from ..registry import HAVE_MMDET, POSITIONAL_ENCODING


@POSITIONAL_ENCODING.register_module(name=None if not HAVE_MMDET else 'UBLearnedPositionalEncoding')
class LearnedPositionalEncoding(nn.Module):
    def __init__(self, num_feats, row_num_embed=50, col_num_embed=50, init_cfg=None):
        super().__init__()
        self.row_embed = nn.Embedding(row_num_embed, num_feats)
        self.col_embed = nn.Embedding(col_num_embed, num_feats)
        self.num_feats, self.row_num_embed, self.col_num_embed = num_feats, row_num_embed, col_num_embed
        self.init_weights()

    def init_weights(self):          # mmdet init_cfg: Uniform over nn.Embedding
        nn.init.uniform_(self.row_embed.weight)
        nn.init.uniform_(self.col_embed.weight)

    def forward(self, mask):
        """mask (bs, h, w), only its shape and device are used -> (bs, 2 * num_feats, h, w)."""
        h, w = mask.shape[-2:]
        if h > self.row_num_embed or w > self.col_num_embed:
            raise ValueError(f'mask {h} x {w} exceeds the {self.row_num_embed} x {self.col_num_embed} embedding tables')
        x = torch.arange(w, device=mask.device)
        y = torch.arange(h, device=mask.device)
        x_embed, y_embed = self.col_embed(x), self.row_embed(y)
        pos = torch.cat((x_embed.unsqueeze(0).repeat(h, 1, 1), y_embed.unsqueeze(1).repeat(1, w, 1)), dim=-1)
        return pos.permute(2, 0, 1).unsqueeze(0).repeat(mask.shape[0], 1, 1, 1)

    def __repr__(self):
        return (f'{type(self).__name__}(num_feats={self.num_feats}, row_num_embed={self.row_num_embed}, '
                f'col_num_embed={self.col_num_embed})')
