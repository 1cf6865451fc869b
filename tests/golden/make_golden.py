"""Generate the golden fixtures that pin ``oracle/galerkin_oracle.py`` (and, on the GPU
box, the CUDA path) to the reference.

Runs ONLY in the build container, where the reference checkout is mounted read-only at
/root/reference; the GPU box never sees it.  The reference modules are imported as-is
(libs/layers.py, libs/model.py), constructed with small configurations of every hot-path
operator, given randomly perturbed parameters (so LayerNorm affine / bias bugs are
visible), run forward + backward in fp32 with the always-on attention dropout
(libs/layers.py:700-701, 730-731) either neutralised or replaced by an explicit,
recorded keep-mask, and the inputs / state_dict / outputs / gradients are saved as
``tests/golden/<case>.pt``.

    python tests/golden/make_golden.py            # rewrites every fixture
"""
import os
import sys
import types
import contextlib

import torch
import yaml

REF = os.environ.get("GALERKIN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    class _Stub(types.ModuleType):
        """Absent plotting / notebook dependencies: any attribute is a no-op callable."""
        __path__ = []

        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return lambda *a, **k: None

    for name in ("torchinfo", "matplotlib", "matplotlib.pyplot", "IPython", "h5py",
                 "IPython.display", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    sys.path.insert(0, os.path.join(REF, "libs"))
    import layers  # noqa
    import model   # noqa
    return layers, model


@contextlib.contextmanager
def attn_dropout_as(layers, masks=None):
    """Replace the reference's F.dropout by identity, or by recorded keep-masks
    (consumed in call order; value kept is scaled by 1/(1-p) = 2 as F.dropout does)."""
    F = layers.F
    orig = F.dropout
    queue = list(masks) if masks is not None else None

    def fake(x, *a, **k):
        if queue is None:
            return x
        m = queue.pop(0)
        assert m.shape == x.shape, (m.shape, x.shape)
        return x * (m.to(x.dtype) * 2.0)
    F.dropout = fake
    try:
        yield
    finally:
        F.dropout = orig


def perturb(module, gen, scale=0.3):
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.ndim == 1:                       # biases, LayerNorm affine
                p.add_(scale * torch.randn(p.shape, generator=gen))
            elif "fourier_weight" in name:        # tiny by init (gain 1/(in*out))
                p.add_(0.2 * torch.randn(p.shape, generator=gen))
            else:
                p.add_(0.05 * torch.randn(p.shape, generator=gen))


def zero_dropouts(module):
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0


def run_case(module, fwd, inputs, grad_inputs, layers, masks=None):
    """fwd(module, **inputs) -> tensor or tuple of tensors (first one is differentiated)."""
    for k in grad_inputs:
        inputs[k].requires_grad_(True)
    with attn_dropout_as(layers, masks):
        out = fwd(module, **inputs)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    outs = [o for o in outs if torch.is_tensor(o)]
    gen = torch.Generator().manual_seed(7)
    cot = torch.randn(outs[0].shape, generator=gen)
    params = dict(module.named_parameters())
    grads = torch.autograd.grad((outs[0] * cot).sum(),
                                [inputs[k] for k in grad_inputs] + list(params.values()),
                                allow_unused=True)
    gi = {k: g.detach() for k, g in zip(grad_inputs, grads[:len(grad_inputs)])}
    gp = {k: g.detach() for k, g in zip(params.keys(), grads[len(grad_inputs):])
          if g is not None}
    return dict(outputs=[o.detach() for o in outs], cotangent=cot,
                grad_inputs=gi, grad_params=gp)


def save(name, cfg, module, inputs, result, masks=None):
    blob = dict(name=name, config=cfg,
                state_dict={k: v.detach().clone() for k, v in module.state_dict().items()},
                inputs={k: v.detach() for k, v in inputs.items()}, masks=masks, **result)
    path = os.path.join(HERE, name + ".pt")
    torch.save(blob, path)
    print(f"{name:34s} {os.path.getsize(path)/1024:8.1f} KiB  out {tuple(result['outputs'][0].shape)}")


def mesh_pos(b, n, dim):
    g = torch.linspace(0, 1, n)
    if dim == 1:
        return g[None, :, None].repeat(b, 1, 1)
    xx, yy = torch.meshgrid(g, g, indexing="ij")
    return torch.stack([xx, yy], -1).reshape(1, n * n, 2).repeat(b, 1, 1)


def scaler_sizes_ref(n_f, n_c):
    """DarcyDataset.get_scaler_sizes (libs/ft.py:698-714) evaluated without importing ft."""
    import numpy as np
    factor = np.round(np.sqrt(n_c / n_f), 4)
    last_digit = float(str(factor)[-1])
    factor = np.round(factor, 3)
    if last_digit < 5:
        factor += 5e-3
    factor = int(factor / 5e-3 + 5e-1) * 5e-3
    n_m = round(n_f * factor) - 1
    return (factor, factor), ((n_m, n_m), (n_f, n_f))


def main():
    layers, model = _import_reference()
    torch.manual_seed(1127802)
    gen = torch.Generator().manual_seed(1127802)
    rn = lambda *s: torch.randn(*s, generator=gen)

    # ---------------- SimpleAttention (layers.py:764-951) ----------------
    attn_cases = [
        ("attn_galerkin_p2", dict(n_head=4, d_model=32, pos_dim=2, attention_type="galerkin",
                                  norm=True, eps=1e-7), 2, 49, False),
        ("attn_galerkin_p1_mask", dict(n_head=2, d_model=32, pos_dim=1, attention_type="galerkin",
                                       norm=True, eps=1e-5), 2, 64, True),
        ("attn_galerkin_nonorm_h1", dict(n_head=1, d_model=48, pos_dim=2, attention_type="galerkin",
                                         norm=False), 2, 36, False),
        ("attn_fourier_p2", dict(n_head=4, d_model=32, pos_dim=2, attention_type="fourier",
                                 norm=True, eps=1e-7), 2, 49, False),
        ("attn_fourier_p1_mask", dict(n_head=2, d_model=16, pos_dim=1, attention_type="fourier",
                                      norm=True, eps=1e-5), 2, 40, True),
    ]
    for name, cfg, b, n, use_mask in attn_cases:
        m = layers.SimpleAttention(**cfg)
        perturb(m, gen)
        zero_dropouts(m)
        x = rn(b, n, cfg["d_model"])
        pos = mesh_pos(b, int(round(n ** (1 / cfg["pos_dim"]))), cfg["pos_dim"])[:, :n]
        d = cfg["d_model"] // cfg["n_head"] + cfg["pos_dim"]
        masks = None
        if use_mask:
            shape = (b, cfg["n_head"], d, d) if cfg["attention_type"] == "galerkin" \
                else (b, cfg["n_head"], n, n)
            masks = [(torch.rand(shape, generator=gen) > 0.5).to(torch.uint8)]
        inputs = dict(x=x, pos=pos)
        res = run_case(m, lambda mod, x, pos: mod(x, x, x, pos=pos), inputs, ["x"], layers, masks)
        save(name, cfg, m, inputs, res, masks)

    # cross-attention style call: distinct query / key / value tensors
    cfg = dict(n_head=2, d_model=16, pos_dim=1, attention_type="galerkin", norm=True, eps=1e-5)
    m = layers.SimpleAttention(**cfg); perturb(m, gen); zero_dropouts(m)
    inputs = dict(q=rn(2, 30, 16), k=rn(2, 30, 16), v=rn(2, 30, 16), pos=mesh_pos(2, 30, 1))
    res = run_case(m, lambda mod, q, k, v, pos: mod(q, k, v, pos=pos), inputs, ["q", "k", "v"], layers)
    save("attn_galerkin_qkv_distinct", cfg, m, inputs, res)

    # ---------------- SimpleTransformerEncoderLayer (model.py:33-140) ----------------
    enc_cases = [
        ("enc_galerkin_attnnorm", dict(d_model=32, n_head=4, pos_dim=2, dim_feedforward=64,
                                       attention_type="galerkin", layer_norm=False, attn_norm=True,
                                       norm_eps=1e-7, dropout=0.0, ffn_dropout=0.0), 2, 49),
        ("enc_galerkin_layernorm_h1", dict(d_model=48, n_head=1, pos_dim=2, dim_feedforward=96,
                                           attention_type="galerkin", layer_norm=True, attn_norm=False,
                                           dropout=0.0, ffn_dropout=0.0), 2, 36),
        ("enc_fourier_minus", dict(d_model=32, n_head=2, pos_dim=1, dim_feedforward=48,
                                   attention_type="fourier", layer_norm=False, attn_norm=True,
                                   residual_type="minus", dropout=0.0, ffn_dropout=0.0), 2, 40),
    ]
    for name, cfg, b, n in enc_cases:
        m = model.SimpleTransformerEncoderLayer(**cfg)
        perturb(m, gen)
        zero_dropouts(m)
        inputs = dict(x=rn(b, n, cfg["d_model"]),
                      pos=mesh_pos(b, int(round(n ** (1 / cfg["pos_dim"]))), cfg["pos_dim"])[:, :n])
        res = run_case(m, lambda mod, x, pos: mod(x, pos), inputs, ["x"], layers)
        save(name, cfg, m, inputs, res)

    # ---------------- SpectralConv1d / 2d (layers.py:1040-1197) ----------------
    for name, cfg, shape in [
        ("sc1d_n64", dict(in_dim=8, out_dim=6, modes=5, dropout=0.0), (2, 64, 8)),
        ("sc1d_n45_relu", dict(in_dim=4, out_dim=4, modes=7, dropout=0.0, activation="relu"), (3, 45, 4)),
    ]:
        m = layers.SpectralConv1d(**cfg); perturb(m, gen); zero_dropouts(m)
        inputs = dict(x=rn(*shape))
        res = run_case(m, lambda mod, x: mod(x), inputs, ["x"], layers)
        save(name, cfg, m, inputs, res)
    for name, cfg, shape in [
        ("sc2d_n15", dict(in_dim=6, out_dim=5, modes=4, dropout=0.0), (2, 15, 15, 6)),
        ("sc2d_n16_flat", dict(in_dim=4, out_dim=4, modes=3, dropout=0.0), (2, 256, 4)),
        ("sc2d_n21_m8", dict(in_dim=3, out_dim=7, modes=8, dropout=0.0, activation="relu"), (1, 21, 21, 3)),
    ]:
        m = layers.SpectralConv2d(**cfg); perturb(m, gen); zero_dropouts(m)
        inputs = dict(x=rn(*shape))
        res = run_case(m, lambda mod, x: mod(x), inputs, ["x"], layers)
        save(name, cfg, m, inputs, res)
    # return_freq=True: also pins out_ft
    cfg = dict(in_dim=4, out_dim=3, modes=3, dropout=0.0, return_freq=True)
    m = layers.SpectralConv2d(**cfg); perturb(m, gen); zero_dropouts(m)
    inputs = dict(x=rn(2, 9, 9, 4))
    with attn_dropout_as(layers):
        y, out_ft = m(inputs["x"])
    save("sc2d_n9_freq", cfg, m, inputs,
         dict(outputs=[y.detach(), torch.view_as_real(out_ft.detach())], cotangent=None,
              grad_inputs={}, grad_params={}))

    # ---------------- full models (model.py:752-1283) ----------------
    with open(os.path.join(REF, "config.yml")) as f:
        yml = yaml.full_load(f)

    # FourierTransformer2D, ex2-like: fine 31x31, coarse 11x11, interp scalers
    n_f, n_c, b = 29, 10, 2
    down, up = scaler_sizes_ref(n_f, n_c)
    cfg = dict(yml["ex2_darcy"])
    cfg.update(n_hidden=32, n_head=4, dim_feedforward=64, num_encoder_layers=2, freq_dim=8,
               fourier_modes=4, downscaler_size=down, upscaler_size=up, attn_norm=True,
               norm_eps=1e-7)
    m = model.FourierTransformer2D(**cfg); perturb(m, gen); zero_dropouts(m)
    inputs = dict(node=rn(b, n_f, n_f, 1), pos=mesh_pos(b, n_c, 2),
                  grid=mesh_pos(b, n_f, 2).reshape(b, n_f, n_f, 2))
    res = run_case(m, lambda mod, node, pos, grid: mod(node, None, pos, grid)["preds"],
                   inputs, ["node"], layers)
    save("model_ft2d_darcy_small", cfg, m, inputs, res)

    # FourierTransformer2D, ex3-like: fourier attention, pointwise decoder, no up/down interp sizes changed
    n_f, n_c = 21, 8
    down, _ = scaler_sizes_ref(n_f, n_c)
    cfg = dict(yml["ex3_darcy_inv"])
    cfg.update(n_hidden=32, n_head=4, dim_feedforward=48, num_encoder_layers=2,
               attention_type="fourier", downscaler_size=down, upscaler_size=((n_c, n_c), (n_c, n_c)),
               attn_norm=True, norm_eps=1e-7)
    m = model.FourierTransformer2D(**cfg); perturb(m, gen); zero_dropouts(m)
    inputs = dict(node=rn(b, n_f, n_f, 1), pos=mesh_pos(b, n_c, 2),
                  grid=mesh_pos(b, n_c, 2).reshape(b, n_c, n_c, 2))
    res = run_case(m, lambda mod, node, pos, grid: mod(node, None, pos, grid)["preds"],
                   inputs, ["node"], layers)
    save("model_ft2d_darcyinv_small", cfg, m, inputs, res)

    # SimpleTransformer, ex1-like (Burgers): n=96, Galerkin, 4 heads
    cfg = dict(yml["ex1_burgers"])
    cfg.update(attention_type="galerkin", n_hidden=32, n_head=4, dim_feedforward=64,
               num_encoder_layers=2, freq_dim=12, fourier_modes=6)
    m = model.SimpleTransformer(**cfg); perturb(m, gen); zero_dropouts(m)
    inputs = dict(node=rn(b, 96, 1), pos=mesh_pos(b, 96, 1))
    res = run_case(m, lambda mod, node, pos: mod(node, None, pos)["preds"], inputs, ["node"], layers)
    save("model_simple_burgers_small", cfg, m, inputs, res)

    # FourierTransformer2DLite, ex4-like (Navier-Stokes): 12x12 grid, 1 head, post-LN
    cfg = dict(node_feats=4 + 2, pos_dim=2, n_targets=1, n_hidden=24, num_feat_layers=0,
               num_encoder_layers=2, n_head=1, dim_feedforward=48, attention_type="galerkin",
               feat_extract_type=None, xavier_init=0.01, diagonal_weight=0.01, layer_norm=True,
               attn_norm=False, return_attn_weight=False, return_latent=False, decoder_type="ifft",
               freq_dim=10, num_regressor_layers=2, fourier_modes=4, spacial_dim=2,
               spacial_fc=False, dropout=0.0, encoder_dropout=0.0, decoder_dropout=0.0,
               ffn_dropout=0.0, debug=False)
    m = model.FourierTransformer2DLite(**cfg); perturb(m, gen); zero_dropouts(m)
    ng = 12
    inputs = dict(node=rn(b, ng, ng, 4), pos=mesh_pos(b, ng, 2),
                  grid=mesh_pos(b, ng, 2).reshape(b, ng, ng, 2))
    res = run_case(m, lambda mod, node, pos, grid: mod(node, None, pos, grid)["preds"],
                   inputs, ["node"], layers)
    save("model_ft2dlite_ns_small", cfg, m, inputs, res)


if __name__ == "__main__":
    main()
