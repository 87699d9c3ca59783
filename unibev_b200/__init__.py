"""unibev_b200 -- B200-native implementation of UniBEV's uniform BEV encoder hot path.

``unibev_b200.plugin`` mirrors the reference plugin surface (registered module names, config
keys, state-dict keys); ``unibev_b200.ops`` wraps the C ABI of ``libunibev_b200.so``
(``include/unibev_b200.h``).  There is no CPU implementation in this package.
"""
from . import _cabi, ops, registry  # noqa: F401
from .registry import build_transformer  # noqa: F401

__version__ = '0.1.0'
