"""One C3-sized encoder layer + one SpectralConv2d, fwd+bwd, a few iterations (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import galerkin_transformer_b200 as G

dev = "cuda"
torch.manual_seed(0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
layer = G.SimpleTransformerEncoderLayer(d_model=128, n_head=4, pos_dim=2, dim_feedforward=256, attention_type="galerkin",
                                        layer_norm=False, attn_norm=True, norm_eps=1e-7, dropout=0.05,
                                        ffn_dropout=0.05).to(dev)
conv = G.SpectralConv2d(32, 32, 12, dropout=0.0).to(dev)
B, n = 8, 1849
g = torch.linspace(0, 1, 43, device=dev)
pos = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(1, -1, 2).repeat(B, 1, 1)
x = torch.randn(B, n, 128, device=dev, requires_grad=True)
xs = torch.randn(B, 141, 141, 32, device=dev, requires_grad=True)
for _ in range(iters):
    y = layer(x, pos)
    y.square().mean().backward()
    z = conv(xs)
    z.square().mean().backward()
torch.cuda.synchronize()
print("done")
