"""Importing this package registers the hot-path modules under the reference's names
(ATTENTION / TRANSFORMER_LAYER / TRANSFORMER_LAYER_SEQUENCE / TRANSFORMER registries)."""
from .attention import (MSDeformableAttention3DImg, MSDeformableAttention3DPts, MultiScaleDeformableAttention,
                        SpatialCrossAttentionImg, SpatialCrossAttentionPts)
from .decoder import (CustomMSDeformableAttention, DetectionTransformerDecoder, DetrTransformerDecoderLayer,
                      MultiheadAttention)
from .coder import NMSFreeCoder, denormalize_bbox
from .encoder import FFN, ImgEncoder, ImgLayer, PtsEncoder, PtsLayer
from .head import UniBEVHead
from .positional import LearnedPositionalEncoding
from .transformer import UniBEVTransformer
from .voxelize import HardSimpleVFE, Voxelization, voxelize

__all__ = ['MSDeformableAttention3DImg', 'MSDeformableAttention3DPts', 'MultiScaleDeformableAttention',
           'SpatialCrossAttentionImg', 'SpatialCrossAttentionPts', 'FFN', 'ImgEncoder', 'ImgLayer', 'PtsEncoder',
           'PtsLayer', 'UniBEVTransformer', 'HardSimpleVFE', 'Voxelization', 'voxelize', 'CustomMSDeformableAttention',
           'DetectionTransformerDecoder', 'DetrTransformerDecoderLayer', 'MultiheadAttention', 'LearnedPositionalEncoding',
           'NMSFreeCoder', 'denormalize_bbox', 'UniBEVHead']
