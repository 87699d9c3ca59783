"""Registries for the plugin surface.

With mmcv installed the modules register into mmcv's own ATTENTION /
TRANSFORMER_LAYER / TRANSFORMER_LAYER_SEQUENCE registries and mmdet's TRANSFORMER
registry (exactly what the reference does, e.g. spatial_cross_attention_img.py:23,
transformer_fusion.py:49), so `type=` strings in the reference configs resolve to
the B200 implementations.  Without mmcv (this image) a minimal stand-in with the
same `register_module` / `build` behaviour is used, so the config subtrees still
build standalone.
"""
import copy


class _Registry:
    def __init__(self, name):
        self.name = name
        self._table = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self._table and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._table[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self._table.get(key)

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or ('type' not in cfg and not (default_args and 'type' in default_args)):
            raise TypeError(f'cfg must be a dict with a "type" key, got {cfg!r}')
        args = copy.copy(dict(cfg))
        for k, v in (default_args or {}).items():
            args.setdefault(k, v)
        typ = args.pop('type')
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f'{typ} is not in the {self.name} registry')
        return cls(**args)


try:  # pragma: no cover - mmcv is absent from the build image
    from mmcv.cnn.bricks.registry import (ATTENTION, FEEDFORWARD_NETWORK, TRANSFORMER_LAYER,
                                          TRANSFORMER_LAYER_SEQUENCE)
    from mmdet.models.utils.builder import TRANSFORMER
    HAVE_MMCV = True
except Exception:  # noqa: BLE001
    ATTENTION = _Registry('attention')
    FEEDFORWARD_NETWORK = _Registry('feed-forward network')
    TRANSFORMER_LAYER = _Registry('transformerLayer')
    TRANSFORMER_LAYER_SEQUENCE = _Registry('transformer-layers sequence')
    TRANSFORMER = _Registry('Transformer')
    HAVE_MMCV = False

try:  # pragma: no cover - mmdet is absent from the build image
    from mmdet.core.bbox.builder import BBOX_CODERS
    from mmdet.models import HEADS
    from mmdet.models.utils.builder import POSITIONAL_ENCODING
    HAVE_MMDET = True
except Exception:  # noqa: BLE001
    BBOX_CODERS = _Registry('bbox_coder')
    HEADS = _Registry('head')
    POSITIONAL_ENCODING = _Registry('position encoding')
    HAVE_MMDET = False


def build_attention(cfg):
    return ATTENTION.build(cfg)


def build_feedforward_network(cfg, default_args=None):
    return FEEDFORWARD_NETWORK.build(cfg, default_args)


def build_transformer_layer(cfg):
    return TRANSFORMER_LAYER.build(cfg)


def build_transformer_layer_sequence(cfg):
    return TRANSFORMER_LAYER_SEQUENCE.build(cfg)


def build_transformer(cfg):
    return TRANSFORMER.build(cfg)


def build_bbox_coder(cfg):
    return BBOX_CODERS.build(cfg)


def build_positional_encoding(cfg):
    return POSITIONAL_ENCODING.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)
