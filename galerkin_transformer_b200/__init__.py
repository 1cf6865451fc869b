"""galerkin_transformer_b200 -- B200-native (sm_100a) Galerkin/Fourier attention encoder and
spectral-convolution decoder, drop-in for the hot path of scaomath/galerkin-transformer.

    from galerkin_transformer_b200 import SimpleAttention, SpectralConv2d, FourierTransformer2D

The CUDA library (csrc/libgalerkin_b200.so, C ABI in include/galerkin_b200.h) is loaded lazily on
the first operator call; it is required -- there is no CPU or eager-PyTorch fallback.
"""
from .layers import FeedForward, Identity, SimpleAttention, SpectralConv1d, SpectralConv2d  # noqa: F401
from .model import (DownScaler, FourierTransformer2D, FourierTransformer2DLite,  # noqa: F401
                    PointwiseRegressor, SimpleTransformer, SimpleTransformerEncoderLayer,
                    SpectralRegressor, UpScaler)
from .functional import get_precision, set_precision  # noqa: F401
from .utils import scaler_sizes, set_attn_dropout  # noqa: F401
from .train import FusedAdam, WeightedL2Loss2d, one_cycle, train_batch_darcy  # noqa: F401

__version__ = "0.1.0"
