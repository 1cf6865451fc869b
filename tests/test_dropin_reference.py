"""Build container only (skipped where /root/reference is absent, e.g. the GPU box): patching the
reference's own namespaces makes ITS model assembly build our operators, with an unchanged
state_dict layout."""
import os
import sys

import pytest
import yaml

REF = os.environ.get("GALERKIN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "libs")), reason="reference not mounted")


def test_patch_rebinds_reference_namespaces():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import _import_reference, scaler_sizes_ref
    layers, model = _import_reference()
    import galerkin_transformer_b200 as G
    from galerkin_transformer_b200.dropin import patch

    with open(os.path.join(REF, "config.yml")) as f:
        cfg = dict(yaml.full_load(f)["ex2_darcy"])
    down, up = scaler_sizes_ref(29, 10)
    cfg.update(downscaler_size=down, upscaler_size=up, n_hidden=32, dim_feedforward=64, freq_dim=8,
               fourier_modes=4, num_encoder_layers=2)
    before = model.FourierTransformer2D(**cfg)
    saved = {n: getattr(model, n) for n in ("SimpleAttention", "SpectralConv2d", "SimpleTransformerEncoderLayer",
                                            "SpectralRegressor", "FeedForward", "SpectralConv1d",
                                            "PointwiseRegressor")}
    try:
        done = patch(layers, model)
        assert "SimpleAttention" in done[layers.__name__] and "SpectralRegressor" in done[model.__name__]
        after = model.FourierTransformer2D(**cfg)          # the REFERENCE's assembly code
        assert isinstance(after.encoder_layers[0], G.SimpleTransformerEncoderLayer)
        assert isinstance(after.encoder_layers[0].attn, G.SimpleAttention)
        assert isinstance(after.regressor.spectral_conv[0], G.SpectralConv2d)
        assert list(after.state_dict().keys()) == list(before.state_dict().keys())
        after.load_state_dict(before.state_dict())
    finally:
        for n, c in saved.items():
            setattr(model, n, c)
            if hasattr(layers, n):
                setattr(layers, n, c)
