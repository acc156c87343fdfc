"""ORACLE helper (test infrastructure only): import the reference's own EasyAnimateTransformer3DModel from
/root/reference (easyanimate/models/{transformer3d,attention,processor,norm,...}.py, executed unmodified) with the
third-party `diffusers` primitives it needs supplied by oracle/_refshim (see its docstring).

Works only where /root/reference exists (the authoring container); used to validate oracle/dit.py
(tests/test_oracle_cpu.py) and to mint tests/golden/dit_*.safetensors.  Never imported on the GPU box or by the
product path."""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "easyanimate", "models", "transformer3d.py"))


def _reference_models():
    if not available():
        raise RuntimeError("/root/reference is not present here")
    try:
        import diffusers  # noqa: F401  (a real install wins over the shim)
    except ImportError:
        shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_refshim")
        if shim not in sys.path:
            sys.path.insert(0, shim)
    # Synthetic parent packages so that `easyanimate/__init__.py` and `easyanimate/models/__init__.py` (which import
    # the text encoders, pipelines, ...) are NOT executed: only the files on the transformer path are.
    if "easyanimate" not in sys.modules:
        pkg = types.ModuleType("easyanimate")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "easyanimate")]
        sys.modules["easyanimate"] = pkg
    if "easyanimate.models" not in sys.modules:
        sub = types.ModuleType("easyanimate.models")
        sub.__path__ = [os.path.join(REFERENCE_ROOT, "easyanimate", "models")]
        sys.modules["easyanimate.models"] = sub
    return importlib.import_module("easyanimate.models.transformer3d")


def reference_transformer(**config):
    """The reference's `EasyAnimateTransformer3DModel(**config)` (transformer3d.py:1347)."""
    return _reference_models().EasyAnimateTransformer3DModel(**config)
