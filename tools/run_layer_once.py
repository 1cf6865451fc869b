"""One fused encoder layer (C3 shape: batch 8, 1849 tokens, d_model 128) forward + backward through the module API, twice,
and one scaler convolution block (128 -> 128 on 8 x 77 x 77) forward + backward: the launch sequence the ncu `--set full`
capture of tools/final_profile.sh records.  x3 precision, reference dropouts on."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import galerkin_transformer_b200 as G                                            # noqa: E402
from galerkin_transformer_b200 import functional as GF                           # noqa: E402

G.set_precision("x3")
dev = "cuda"
torch.manual_seed(0)
layer = G.SimpleTransformerEncoderLayer(d_model=128, n_head=4, pos_dim=2, dim_feedforward=256, attention_type="galerkin",
                                        layer_norm=False, attn_norm=True, norm_eps=1e-7, dropout=0.05, ffn_dropout=0.05).to(dev)
layer.train()
x = torch.randn(8, 1849, 128, device=dev, requires_grad=True)
g = torch.linspace(0, 1, 43, device=dev)
pos = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(1, -1, 2).repeat(8, 1, 1)
for _ in range(2):
    y = layer(x, pos)
    y.square().mean().backward()
xc = torch.randn(8, 77, 77, 128, device=dev, requires_grad=True)
w = (torch.randn(128, 128, 3, 3, device=dev) / 34).requires_grad_(True)
yc = GF.conv3x3_block(xc, w, act="silu", drop_p=0.0)
yc.square().mean().backward()
torch.cuda.synchronize()
print("done")
