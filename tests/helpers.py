"""Shared helpers for the parity tests: fixture loading and oracle dispatch."""
import glob
import os

import torch

from oracle import galerkin_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.pt")))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def oracle_forward(fix, sd, inputs):
    """Run the oracle for a golden case.  Returns a tuple of output tensors."""
    name, cfg = fix["name"], fix["config"]
    mask = fix["masks"][0] if fix.get("masks") else None
    if mask is not None:
        mask = mask.to(next(iter(inputs.values())).device)
    if name.startswith("attn_"):
        kw = dict(n_head=cfg["n_head"], attention_type=cfg["attention_type"],
                  norm=cfg.get("norm", False), eps=cfg.get("eps", 1e-5),
                  pos_dim=cfg["pos_dim"], attn_mask=mask)
        if "x" in inputs:
            q = k = v = inputs["x"]
        else:
            q, k, v = inputs["q"], inputs["k"], inputs["v"]
        return O.simple_attention(sd, "", q, k, v, inputs["pos"], **kw)
    if name.startswith("enc_"):
        return (O.encoder_layer(sd, "", inputs["x"], inputs["pos"], n_head=cfg["n_head"],
                                attention_type=cfg["attention_type"],
                                layer_norm=cfg["layer_norm"], attn_norm=cfg["attn_norm"],
                                norm_eps=cfg.get("norm_eps"), pos_dim=cfg["pos_dim"],
                                residual_type=cfg.get("residual_type", "add")),)
    if name.startswith("sc1d_"):
        return (O.spectral_conv1d(sd, "", inputs["x"], modes=cfg["modes"],
                                  activation=cfg.get("activation", "silu")),)
    if name.startswith("sc2d_"):
        out = O.spectral_conv2d(sd, "", inputs["x"], modes=cfg["modes"],
                                activation=cfg.get("activation", "silu"),
                                return_freq=cfg.get("return_freq", False))
        if cfg.get("return_freq"):
            return (out[0], torch.view_as_real(out[1]))
        return (out,)
    if name.startswith("model_ft2d_"):
        return (O.fourier_transformer_2d(sd, cfg, inputs["node"], inputs["pos"], inputs["grid"]),)
    if name.startswith("model_simple_"):
        return (O.simple_transformer(sd, cfg, inputs["node"], inputs["pos"]),)
    if name.startswith("model_ft2dlite_"):
        return (O.fourier_transformer_2d_lite(sd, cfg, inputs["node"], inputs["pos"], inputs["grid"]),)
    raise KeyError(name)


def oracle_grads(fix, dtype=torch.float32, device="cpu"):
    """Forward + backward of the oracle on a golden case with the recorded cotangent.

    Returns (outputs, grad_inputs dict, grad_params dict)."""
    sd = {k: v.to(device=device, dtype=dtype).requires_grad_(v.is_floating_point())
          for k, v in fix["state_dict"].items()}
    inputs = {k: v.to(device=device, dtype=dtype) for k, v in fix["inputs"].items()}
    gnames = list(fix["grad_inputs"].keys())
    for k in gnames:
        inputs[k].requires_grad_(True)
    outs = oracle_forward(fix, sd, inputs)
    gi, gp = {}, {}
    if fix.get("cotangent") is not None:
        cot = fix["cotangent"].to(device=device, dtype=dtype)
        pnames = list(fix["grad_params"].keys())
        grads = torch.autograd.grad((outs[0] * cot).sum(),
                                    [inputs[k] for k in gnames] + [sd[k] for k in pnames],
                                    allow_unused=True)
        gi = dict(zip(gnames, grads[:len(gnames)]))
        gp = dict(zip(pnames, grads[len(gnames):]))
    return outs, gi, gp
