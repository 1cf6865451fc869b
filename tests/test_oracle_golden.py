"""CPU: the oracle restatement reproduces the reference's recorded outputs and
gradients (fixtures made by tests/golden/make_golden.py from the reference itself)."""
import pytest
import torch

from helpers import golden_names, load_golden, oracle_grads, rel_l2

FWD_TOL = 2e-6     # fp32 restatement vs fp32 reference, same op order up to fusion
GRAD_TOL = 2e-5


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_fixture(name):
    fix = load_golden(name)
    outs, gi, gp = oracle_grads(fix)
    assert len(outs) >= len(fix["outputs"]) or name.startswith("attn_")
    for o, ref in zip(outs, fix["outputs"]):
        assert o.shape == ref.shape
        assert rel_l2(o, ref) <= FWD_TOL, (name, rel_l2(o, ref))
    for k, ref in fix["grad_inputs"].items():
        assert rel_l2(gi[k], ref) <= GRAD_TOL, (name, k, rel_l2(gi[k], ref))
    for k, ref in fix["grad_params"].items():
        assert gp[k] is not None, (name, k)
        assert rel_l2(gp[k], ref) <= GRAD_TOL, (name, k, rel_l2(gp[k], ref))


def test_oracle_fp64_agrees_with_fp32_fixture():
    """The oracle in fp64 stays within fp32 round-off of the recorded fp32 reference."""
    fix = load_golden("enc_galerkin_attnnorm")
    outs, _, _ = oracle_grads(fix, dtype=torch.float64)
    assert rel_l2(outs[0], fix["outputs"][0]) <= 5e-6
