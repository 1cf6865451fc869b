"""Reference `libs/model.py` (its model assembly code) over the B200 operators (see package docstring)."""
import os

import galerkin_transformer.layers  # noqa: F401  (patched operator namespace first: libs/model.py:1-6 imports it)
from galerkin_transformer import export, load_reference_module
from galerkin_transformer_b200 import model as _b200
from galerkin_transformer_b200.dropin import LAYER_CLASSES, MODEL_CLASSES, patch

_ref = load_reference_module("model")
_NATIVE = ("FourierTransformer2D", "SimpleTransformer", "FourierTransformer2DLite")
if _ref is not None:
    patch(_ref)                    # the assembly code now builds the B200 operators / fusion units
    export(globals(), _ref)
    if os.environ.get("GALERKIN_B200_NATIVE_MODELS", "0") == "1":
        for _n in _NATIVE:
            globals()[_n] = getattr(_b200, _n)
else:
    for _n in MODEL_CLASSES + _NATIVE + ("DownScaler", "UpScaler"):
        globals()[_n] = getattr(_b200, _n)
    from galerkin_transformer_b200.layers import *  # noqa: F401,F403
