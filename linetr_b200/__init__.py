"""linetr_b200 - B200-native LineTR line-descriptor forward + mutual-NN matcher.

Drop-in names (same call surface as yosungho/LineTR `models/`):
    LineTransformer, get_dist_matrix, nn_matcher, nn_matcher_distmat
Batched front-end:
    PairEngine, LineBatch
`install_as_reference_models()` registers this package's modules under the reference's
module names (`models.line_transformer`, `models.nn_matcher`, `models.line_process`) so that
the reference's own `models/matching.py` / `match_line_pairs.py` run unchanged on top of it.
"""
from __future__ import annotations

import sys
import types

__version__ = "0.1.0"

_LAZY = {
    "LineTransformer": ("line_transformer", "LineTransformer"),
    "get_dist_matrix": ("line_process", "get_dist_matrix"),
    "nn_matcher": ("nn_matcher", "nn_matcher"),
    "nn_matcher_distmat": ("nn_matcher", "nn_matcher_distmat"),
    "PairEngine": ("engine", "PairEngine"),
    "LineBatch": ("engine", "LineBatch"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)


def install_as_reference_models(package: str = "models"):
    """Make `from models.line_transformer import LineTransformer, get_dist_matrix` and
    `from models.nn_matcher import nn_matcher, nn_matcher_distmat` (reference
    models/matching.py:5-6) resolve to this package.  Call before importing `models.matching`
    from a reference checkout; the remaining reference modules (superpoint, line_detector,
    matching) are left untouched."""
    from . import line_process, line_transformer, nn_matcher as nnm
    if package not in sys.modules:
        try:
            __import__(package)
        except Exception:
            pkg = types.ModuleType(package)
            pkg.__path__ = []
            sys.modules[package] = pkg
    sys.modules[f"{package}.line_transformer"] = line_transformer
    sys.modules[f"{package}.nn_matcher"] = nnm
    sys.modules[f"{package}.line_process"] = line_process
    pkg = sys.modules[package]
    pkg.line_transformer, pkg.nn_matcher, pkg.line_process = line_transformer, nnm, line_process
