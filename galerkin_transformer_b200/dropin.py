"""Drop-in seam: rebind the reference's hot-path classes to the B200 implementations.

The reference's scripts do `from libs import *`, and `libs/__init__.py` / `libs/model.py:1-6`
resolve `layers`, `model`, ... either as top-level modules (with `libs/` on sys.path) or as the
installed package `galerkin_transformer`.  `patch()` takes those already-imported module objects and
replaces, in each namespace, the operator classes named by BASELINE north_star
(`SimpleAttention`, `SpectralConv1d/2d`, plus the fusion units `FeedForward`,
`SimpleTransformerEncoderLayer`, `SpectralRegressor`, `PointwiseRegressor`).  Model assembly
(`FourierTransformer2D`, `SimpleTransformer`, ...), datasets, losses and the training loop stay
the reference's own code; constructor arguments and state_dict keys are identical, so released
checkpoints load unchanged.

    import layers, model                     # the reference's modules (libs/ on sys.path)
    from galerkin_transformer_b200.dropin import patch
    patch(layers, model)                     # examples/ex{1,2,3,4}_*.py then run unchanged
"""
from . import layers as _L
from . import model as _M

LAYER_CLASSES = ("SimpleAttention", "FeedForward", "SpectralConv1d", "SpectralConv2d")
MODEL_CLASSES = ("SimpleTransformerEncoderLayer", "SpectralRegressor", "PointwiseRegressor")


def patch(*namespaces):
    """Rebind the hot-path classes in every given module namespace; returns {module: [names]}."""
    done = {}
    for ns in namespaces:
        names = []
        for name in LAYER_CLASSES:
            if hasattr(ns, name):
                setattr(ns, name, getattr(_L, name))
                names.append(name)
        for name in MODEL_CLASSES:
            if hasattr(ns, name):
                setattr(ns, name, getattr(_M, name))
                names.append(name)
        done[getattr(ns, "__name__", repr(ns))] = names
    return done
