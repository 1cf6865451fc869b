"""Small host-side helpers."""
import math

from torch import nn


def scaler_sizes(n_f, n_c, scale_factor=True):
    """Interpolation sizes for the down/up-scalers given fine/coarse grid sizes
    (same numbers as DarcyDataset.get_scaler_sizes, libs/ft.py:698-714):
    141, 43 -> ((0.555, 0.555), ((77, 77), (141, 141)))."""
    factor = round(math.sqrt(n_c / n_f), 4)
    last_digit = float(str(factor)[-1])
    factor = round(factor, 3)
    if last_digit < 5:
        factor += 5e-3
    factor = int(factor / 5e-3 + 5e-1) * 5e-3
    n_m = round(n_f * factor) - 1
    up = ((n_m, n_m), (n_f, n_f))
    if scale_factor:
        return (factor, factor), up
    return ((n_m, n_m), (n_c, n_c)), up


def set_attn_dropout(module: nn.Module, mode: str):
    """Set every SimpleAttention under `module` to 'reference' (always-on p=0.5 dropout on the
    attention matrix, as libs/layers.py:730-731) or 'off'."""
    from .layers import SimpleAttention
    assert mode in ("reference", "off")
    for m in module.modules():
        if isinstance(m, SimpleAttention):
            m.attn_dropout = mode
    return module
