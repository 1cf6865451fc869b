"""The reference's own `libs/utils_ft.py` (out of the hot path; re-exported when the reference sources are available)."""
from galerkin_transformer import export, load_reference_module

_ref = load_reference_module("utils_ft")
if _ref is not None:
    export(globals(), _ref)
