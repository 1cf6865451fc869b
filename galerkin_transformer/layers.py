"""Reference `libs/layers.py` namespace with the hot-path operators served by libgalerkin_b200 (see package docstring)."""
from galerkin_transformer import export, load_reference_module
from galerkin_transformer_b200 import layers as _b200
from galerkin_transformer_b200.dropin import LAYER_CLASSES

_ref = load_reference_module("layers")
if _ref is not None:
    export(globals(), _ref)
for _n in LAYER_CLASSES + ("Identity",):
    globals()[_n] = getattr(_b200, _n)
if _ref is not None:               # the reference's own helper classes that build these operators see the B200 ones too
    for _n in LAYER_CLASSES:
        setattr(_ref, _n, getattr(_b200, _n))
