"""Stage-by-stage check of the fused encoder kernels against a torch fp64 restatement (GPU box diagnostics).

    python tools/debug_fused.py --stage {pack,1,2,3,all,layer} [--B 2 --n 300]

Each stage is fed REFERENCE inputs, so a failure is local to one kernel.  Prints relative L2 errors."""
import argparse
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import galerkin_transformer_b200 as G                                           # noqa: E402
from galerkin_transformer_b200 import _lib, functional as GF                    # noqa: E402

DM, H, DK, DFF, HP = 128, 4, 32, 256, 48


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def make(B, n, p, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc)
    d = DK + p
    P = dict(wq=rn(DM, DM, sc=0.1), wk=rn(DM, DM, sc=0.1), wv=rn(DM, DM, sc=0.1), bq=rn(DM, sc=0.1), bk=rn(DM, sc=0.1),
             bv=rn(DM, sc=0.1), gk=[1 + rn(DK, sc=0.2) for _ in range(H)], bek=[rn(DK, sc=0.2) for _ in range(H)],
             gv=[1 + rn(DK, sc=0.2) for _ in range(H)], bev=[rn(DK, sc=0.2) for _ in range(H)],
             wfc=rn(DM, H * d, sc=0.1), bfc=rn(DM, sc=0.1), w1=rn(DFF, DM, sc=0.1), b1=rn(DFF, sc=0.1),
             w2=rn(DM, DFF, sc=0.1), b2=rn(DM, sc=0.1))
    x = rn(B, n, DM)
    pos = torch.rand(B, n, p, device="cuda", generator=g)
    keep = (torch.rand(B, H, d, d, device="cuda", generator=g) > 0.5).to(torch.uint8)
    return P, x, pos, keep


def reference(P, x, pos, keep, eps, sign=1.0):
    """fp64 layer; returns every intermediate the fused kernels write."""
    D = lambda t: t.double()
    B, n, _ = x.shape
    p = pos.shape[-1]
    d = DK + p
    X = D(x).reshape(B * n, DM)
    q = X @ D(P["wq"]).t() + D(P["bq"])
    k = X @ D(P["wk"]).t() + D(P["bk"])
    v = X @ D(P["wv"]).t() + D(P["bv"])

    def hn(t, gam, bet):
        t = t.view(B * n, H, DK)
        mu = t.mean(-1, keepdim=True)
        var = ((t - mu) ** 2).mean(-1, keepdim=True)
        rs = 1.0 / torch.sqrt(var + eps)
        xh = (t - mu) * rs
        aff = xh * torch.stack([D(g_) for g_ in gam])[None] + torch.stack([D(b_) for b_ in bet])[None]
        return xh.reshape(B * n, DM), rs.squeeze(-1), aff

    kh, rk, ka = hn(k, P["gk"], P["bek"])
    vh, rv, va = hn(v, P["gv"], P["bev"])
    pp = D(pos).reshape(B, n, 1, p).expand(B, n, H, p)
    Kt = torch.cat([pp, ka.view(B, n, H, DK)], -1).permute(0, 2, 1, 3)       # B,H,n,d
    Vt = torch.cat([pp, va.view(B, n, H, DK)], -1).permute(0, 2, 1, 3)
    Qt = torch.cat([pp, q.view(B, n, H, DK)], -1).permute(0, 2, 1, 3)
    Araw = Kt.transpose(-1, -2) @ Vt                                          # B,H,d,d
    A = Araw / n * (2.0 * D(keep))
    heads = (Qt @ A).permute(0, 2, 1, 3).reshape(B * n, H * d)
    x1 = X + sign * (heads @ D(P["wfc"]).t() + D(P["bfc"]))
    hid = torch.relu(x1 @ D(P["w1"]).t() + D(P["b1"]))
    x2 = x1 + hid @ D(P["w2"]).t() + D(P["b2"])
    return dict(qkv=torch.cat([q, kh, vh], 1), rk=rk, rv=rv, Araw=Araw, A=A, heads=heads, x1=x1, hid=hid, x2=x2)


def reference_backward(P, x, pos, keep, eps, dy, sign=1.0):
    """fp64 backward of `reference` for the cotangent dy (B n, 128): every buffer the fused backward kernels write.
    Parameter gradients come from autograd on the same fp64 graph."""
    D = lambda t: t.double()
    B, n, _ = x.shape
    p = pos.shape[-1]
    d = DK + p
    leaves = {k: (D(v).clone().requires_grad_(True) if torch.is_tensor(v) else [D(t).clone().requires_grad_(True) for t in v])
              for k, v in P.items()}
    X = D(x).reshape(B * n, DM).clone().requires_grad_(True)
    q = X @ leaves["wq"].t() + leaves["bq"]
    k = X @ leaves["wk"].t() + leaves["bk"]
    v = X @ leaves["wv"].t() + leaves["bv"]
    for t in (q, k, v):
        t.retain_grad()

    def hn(t, gam, bet):
        t = t.view(B * n, H, DK)
        mu = t.mean(-1, keepdim=True)
        var = ((t - mu) ** 2).mean(-1, keepdim=True)
        xh = (t - mu) / torch.sqrt(var + eps)
        return xh * torch.stack(gam)[None] + torch.stack(bet)[None]

    ka, va = hn(k, leaves["gk"], leaves["bek"]), hn(v, leaves["gv"], leaves["bev"])
    pp = D(pos).reshape(B, n, 1, p).expand(B, n, H, p)
    Kt = torch.cat([pp, ka.view(B, n, H, DK)], -1).permute(0, 2, 1, 3)
    Vt = torch.cat([pp, va.view(B, n, H, DK)], -1).permute(0, 2, 1, 3)
    Qt = torch.cat([pp, q.view(B, n, H, DK)], -1).permute(0, 2, 1, 3)
    A = (Kt.transpose(-1, -2) @ Vt) / n * (2.0 * D(keep))
    heads = (Qt @ A).permute(0, 2, 1, 3).reshape(B * n, H * d)
    heads.retain_grad()
    x1 = X + sign * (heads @ leaves["wfc"].t() + leaves["bfc"])
    x1.retain_grad()
    z1 = x1 @ leaves["w1"].t() + leaves["b1"]
    z1.retain_grad()
    x2 = x1 + torch.relu(z1) @ leaves["w2"].t() + leaves["b2"]
    flat = [leaves[k_] for k_ in ("wq", "wk", "wv", "bq", "bk", "bv")] + leaves["gk"] + leaves["bek"] + leaves["gv"] + \
        leaves["bev"] + [leaves[k_] for k_ in ("wfc", "bfc", "w1", "b1", "w2", "b2")]
    grads = torch.autograd.grad((x2 * D(dy).reshape(B * n, DM)).sum(), [X] + flat, retain_graph=False)
    dO = heads.grad.view(B, n, H, d).permute(0, 2, 1, 3)                       # B,H,n,d
    Graw = Qt.detach().transpose(-1, -2) @ dO                                  # B,H,d,d (unscaled, unmasked)
    o = 7 + 4 * H
    dvec = torch.cat([grads[4], grads[5], grads[6], *grads[7:o], grads[o + 1], grads[o + 3], grads[o + 5]])
    return dict(g1=z1.grad, dx1=x1.grad, dheads=heads.grad, dqkv=torch.cat([q.grad, k.grad, v.grad], 1), dx=grads[0],
                Graw=Graw, dvec=dvec, dwq=grads[1], dwk=grads[2], dwv=grads[3], dwfc=grads[o], dw1=grads[o + 2],
                dw2=grads[o + 4])


def run_bwd_stage(stage, P, x, pos, keep, packed, R, RB, dy, eps):
    """stage bit (1 ffn, 2 attn-out, 4 kv, 8 dx, 16 reduce) fed with reference inputs; 31 = everything chained"""
    lib = _lib.load()
    B, n, _ = x.shape
    p = pos.shape[-1]
    d = DK + p
    T = B * n
    tiles = (n + 127) // 128
    f32 = dict(dtype=torch.float32, device="cuda")
    chain = stage == 31
    qkv = R["qkv"].float().contiguous()
    rk, rv = R["rk"].float().contiguous(), R["rv"].float().contiguous()
    A = R["A"].float().contiguous()
    hid = R["hid"].float().contiguous()
    g2 = torch.zeros(T, DM, **f32)
    g1 = torch.zeros(T, DFF, **f32)
    dx1 = torch.zeros(T, DM, **f32) if (chain or stage == 1) else RB["dx1"].float().contiguous()
    gfc = torch.zeros(T, DM, **f32)
    dqkv = torch.zeros(T, 3 * DM, **f32) if (chain or stage in (2, 4)) else RB["dqkv"].float().contiguous()
    dx = torch.zeros(T, DM, **f32)
    dvec = torch.zeros(1408, **f32)
    nws = lib.gb200_encoder_bwd_workspace_bytes(B, n, H, DK, p) // 4
    ws = torch.zeros(nws, **f32)
    if stage == 4:
        ws[:B * tiles * H * d * d].view(B, tiles, H, d, d)[:, 0] = RB["Graw"].float()
    dyc = dy.reshape(T, DM).contiguous()
    rc = lib.gb200_encoder_layer_bwd(0, packed.data_ptr(), DM, H, p, DFF, dyc.data_ptr(), pos.data_ptr(), B, n, 1, 1.0 / n,
                                     keep.data_ptr(), 0.0, 0, 0.0, 11, 1.0, 0.0, 0.0, 33, qkv.data_ptr(), rk.data_ptr(),
                                     rv.data_ptr(), A.data_ptr(), hid.data_ptr(), g2.data_ptr(), g1.data_ptr(),
                                     dx1.data_ptr(), gfc.data_ptr(), dqkv.data_ptr(), dx.data_ptr(), dvec.data_ptr(),
                                     ws.data_ptr(), ws.numel() * 4, stage, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "encoder_layer_bwd")
    torch.cuda.synchronize()
    gp = ws[:B * tiles * H * d * d].view(B, tiles, H, d, d)
    part = ws[B * tiles * H * d * d:].view(B * tiles, 1408)
    return dict(g1=g1, dx1=dx1, dqkv=dqkv, dx=dx, dvec=dvec, gpart=gp, part=part)


def check_backward(P, x, pos, keep, packed, R, eps, stages=("1", "2", "4", "8", "31"), tol=1e-4):
    g = torch.Generator(device="cuda").manual_seed(77)
    dy = torch.randn(x.shape, device="cuda", generator=g)
    RB = reference_backward(P, x, pos, keep, eps, dy)
    ok = True
    V = dict(bqkv=(0, 384), gk=(384, 512), bk=(512, 640), gv=(640, 768), bv=(768, 896), bfc=(896, 1024), b1=(1024, 1280),
             b2=(1280, 1408))

    def seg(t, name):
        return t[..., V[name][0]:V[name][1]]
    if "1" in stages:
        o = run_bwd_stage(1, P, x, pos, keep, packed, R, RB, dy, eps)
        e = dict(g1=rel(o["g1"], RB["g1"]), dx1=rel(o["dx1"], RB["dx1"]), db1=rel(seg(o["part"].sum(0), "b1"), seg(RB["dvec"], "b1")),
                 db2=rel(seg(o["part"].sum(0), "b2"), seg(RB["dvec"], "b2")))
        print("bwd 1 (ffn)   :", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < tol
    if "2" in stages:
        o = run_bwd_stage(2, P, x, pos, keep, packed, R, RB, dy, eps)
        e = dict(dq=rel(o["dqkv"][:, :128], RB["dqkv"][:, :128]), G=rel(o["gpart"].sum(1), RB["Graw"]),
                 dbfc=rel(seg(o["part"].sum(0), "bfc"), seg(RB["dvec"], "bfc")))
        p = pos.shape[-1]
        Go, Gr = o["gpart"].sum(1).double(), RB["Graw"]
        e["G_ff"] = rel(Go[..., p:, p:], Gr[..., p:, p:])
        e["G_fp"] = rel(Go[..., p:, :p], Gr[..., p:, :p])
        e["G_pf"] = rel(Go[..., :p, p:], Gr[..., :p, p:])
        e["G_pp"] = rel(Go[..., :p, :p], Gr[..., :p, :p])
        print("bwd 2 (attn)  :", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < tol
    if "4" in stages:
        o = run_bwd_stage(4, P, x, pos, keep, packed, R, RB, dy, eps)
        e = dict(dk=rel(o["dqkv"][:, 128:256], RB["dqkv"][:, 128:256]), dv=rel(o["dqkv"][:, 256:], RB["dqkv"][:, 256:]))
        for nm in ("gk", "bk", "gv", "bv"):
            e["d" + nm] = rel(seg(o["part"].sum(0), nm), seg(RB["dvec"], nm))
        print("bwd 3 (kv+LN) :", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < tol
    if "8" in stages:
        o = run_bwd_stage(8, P, x, pos, keep, packed, R, RB, dy, eps)
        e = dict(dx=rel(o["dx"], RB["dx"]), dbqkv=rel(seg(o["part"].sum(0), "bqkv"), seg(RB["dvec"], "bqkv")))
        print("bwd 4 (dx)    :", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < tol
    if "31" in stages:
        o = run_bwd_stage(31, P, x, pos, keep, packed, R, RB, dy, eps)
        e = dict(g1=rel(o["g1"], RB["g1"]), dx1=rel(o["dx1"], RB["dx1"]), dqkv=rel(o["dqkv"], RB["dqkv"]), dx=rel(o["dx"], RB["dx"]))
        for nm in V:
            e["d" + nm] = rel(seg(o["dvec"], nm), seg(RB["dvec"], nm))
        print("bwd all       :", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < tol
    return ok


def unswizzle(tile_u8):
    """16 KB tile image -> (128, 64) bf16 tensor"""
    t = tile_u8.view(torch.int16).view(128, 8, 8)              # row, physical unit, 8 elems
    r = torch.arange(128, device=t.device)[:, None]
    u = torch.arange(8, device=t.device)[None, :]
    phys = u ^ (r & 7)
    out = torch.gather(t, 1, phys[:, :, None].expand(128, 8, 8))
    return out.reshape(128, 64).view(torch.bfloat16)


def params_struct(P, p):
    S = _lib.EncoderParams()
    for k in ("wq", "wk", "wv", "bq", "bk", "bv", "wfc", "bfc", "w1", "b1", "w2", "b2"):
        setattr(S, k, P[k].data_ptr())
    for h in range(H):
        S.gamma_k[h], S.beta_k[h] = P["gk"][h].data_ptr(), P["bek"][h].data_ptr()
        S.gamma_v[h], S.beta_v[h] = P["gv"][h].data_ptr(), P["bev"][h].data_ptr()
    S.d_model, S.n_head, S.pos_dim, S.d_ff = DM, H, p, DFF
    return S


def do_pack(P, p):
    lib = _lib.load()
    nb = lib.gb200_encoder_pack_bytes(DM, H, p, DFF)
    packed = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    S = params_struct(P, p)
    _lib.check(lib.gb200_encoder_pack(0, ctypes.byref(S), packed.data_ptr(), torch.cuda.current_stream().cuda_stream), "pack")
    torch.cuda.synchronize()
    return packed


def check_pack(P, p, packed):
    d = DK + p
    tiles = packed[:70 * 16384].view(70, 16384)

    def tile_pair(i):
        return unswizzle(tiles[2 * i]).float().double() + unswizzle(tiles[2 * i + 1]).float().double()

    errs = {}
    # forward: qkv
    for nb, w in enumerate((P["wq"], P["wk"], P["wv"])):
        for kc in range(2):
            errs[f"qkv{nb}{kc}"] = rel(tile_pair(nb * 2 + kc), w[:, kc * 64:(kc + 1) * 64])
    wfcp = torch.zeros(DM, H * HP, device="cuda")
    for h in range(H):
        wfcp[:, h * HP:h * HP + d] = P["wfc"][:, h * d:(h + 1) * d]
    for kc in range(3):
        errs[f"fc{kc}"] = rel(tile_pair(6 + kc), wfcp[:, kc * 64:(kc + 1) * 64])
    for nb in range(2):
        for kc in range(2):
            errs[f"w1{nb}{kc}"] = rel(tile_pair(9 + nb * 2 + kc), P["w1"][nb * 128:(nb + 1) * 128, kc * 64:(kc + 1) * 64])
    for kc in range(4):
        errs[f"w2{kc}"] = rel(tile_pair(13 + kc), P["w2"][:, kc * 64:(kc + 1) * 64])
    # backward
    w2t = P["w2"].t().contiguous()          # (256, 128)
    for nb in range(2):
        for kc in range(2):
            errs[f"w2t{nb}{kc}"] = rel(tile_pair(17 + nb * 2 + kc), w2t[nb * 128:(nb + 1) * 128, kc * 64:(kc + 1) * 64])
    w1t = P["w1"].t().contiguous()          # (128, 256)
    for kc in range(4):
        errs[f"w1t{kc}"] = rel(tile_pair(21 + kc), w1t[:, kc * 64:(kc + 1) * 64])
    wfct = torch.zeros(256, DM, device="cuda")
    wfct[:H * HP] = wfcp.t()
    for nb in range(2):
        for kc in range(2):
            errs[f"wfct{nb}{kc}"] = rel(tile_pair(25 + nb * 2 + kc) + 1e-30, wfct[nb * 128:(nb + 1) * 128, kc * 64:(kc + 1) * 64] + 1e-30)
    wqkvt = torch.cat([P["wq"], P["wk"], P["wv"]], 0).t().contiguous()      # (128, 384)
    for kc in range(6):
        errs[f"wqkvt{kc}"] = rel(tile_pair(29 + kc), wqkvt[:, kc * 64:(kc + 1) * 64])
    vec = packed[70 * 16384:].view(torch.float32)
    ref = torch.cat([P["bq"], P["bk"], P["bv"], *P["gk"], *P["bek"], *P["gv"], *P["bev"], P["bfc"], P["b1"], P["b2"]])
    errs["vec"] = rel(vec, ref)
    worst = max(errs.values())
    print("pack: worst tile rel err", worst, {k: f"{v:.1e}" for k, v in errs.items() if v > 1e-4})
    return worst < 1e-4


def run_stage(stage, P, x, pos, keep, packed, R, eps, drop=None):
    lib = _lib.load()
    B, n, _ = x.shape
    p = pos.shape[-1]
    d = DK + p
    T = B * n
    tiles = (n + 127) // 128
    f32 = dict(dtype=torch.float32, device="cuda")
    qkv = R["qkv"].float().contiguous() if stage != 1 else torch.full((T, 384), float("nan"), **f32)
    rk, rv = torch.zeros(T, H, **f32), torch.zeros(T, H, **f32)
    A = torch.zeros(B, H, d, d, **f32)
    heads = torch.zeros(T, H * d, **f32)
    x1 = R["x1"].float().contiguous() if stage == 4 else torch.zeros(T, DM, **f32)
    hid = torch.zeros(T, DFF, **f32)
    x2 = torch.zeros(T, DM, **f32)
    ws = torch.zeros(B, tiles, H, d, d, **f32)
    if stage == 2:
        ws[:, 0] = R["Araw"].float()
    dp = drop or dict(p1=0.0, pf=0.0, p2=0.0)
    rc = lib.gb200_encoder_layer_fwd(0, packed.data_ptr(), DM, H, p, DFF, x.data_ptr(), pos.data_ptr(), B, n, 1, eps,
                                     1.0 / n, keep.data_ptr(), 0.0, 0, dp["p1"], 11, 1.0, dp["pf"], 22, dp["p2"], 33,
                                     qkv.data_ptr(), rk.data_ptr(), rv.data_ptr(), A.data_ptr(), heads.data_ptr(),
                                     x1.data_ptr(), hid.data_ptr(), x2.data_ptr(), ws.data_ptr(), ws.numel() * 4, stage,
                                     torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "encoder_layer_fwd")
    torch.cuda.synchronize()
    return dict(qkv=qkv, rk=rk, rv=rv, A=A, heads=heads, x1=x1, hid=hid, x2=x2, ws=ws)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="all")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--p", type=int, default=2)
    args = ap.parse_args()
    eps = 1e-5
    P, x, pos, keep = make(args.B, args.n, args.p)
    R = reference(P, x, pos, keep, eps)
    packed = do_pack(P, args.p)
    ok = True
    if args.stage in ("pack", "all"):
        ok &= check_pack(P, args.p, packed)
    if args.stage in ("1", "all"):
        o = run_stage(1, P, x, pos, keep, packed, R, eps)
        e = dict(q=rel(o["qkv"][:, :128], R["qkv"][:, :128]), kh=rel(o["qkv"][:, 128:256], R["qkv"][:, 128:256]),
                 vh=rel(o["qkv"][:, 256:], R["qkv"][:, 256:]), rk=rel(o["rk"], R["rk"]), rv=rel(o["rv"], R["rv"]),
                 Araw=rel(o["ws"].sum(1), R["Araw"]))
        d = DK + args.p
        Ar, Ao = R["Araw"], o["ws"].sum(1).double()
        e["A_ff"] = rel(Ao[..., args.p:, args.p:], Ar[..., args.p:, args.p:])
        e["A_fp"] = rel(Ao[..., args.p:, :args.p], Ar[..., args.p:, :args.p])
        e["A_pf"] = rel(Ao[..., :args.p, args.p:], Ar[..., :args.p, args.p:])
        e["A_pp"] = rel(Ao[..., :args.p, :args.p], Ar[..., :args.p, :args.p])
        print("stage 1:", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < 1e-4
    if args.stage in ("2", "all"):
        o = run_stage(2, P, x, pos, keep, packed, R, eps)
        e = dict(A=rel(o["A"], R["A"]), heads=rel(o["heads"], R["heads"]), x1=rel(o["x1"], R["x1"]))
        print("stage 2:", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < 1e-4
    if args.stage in ("3", "all"):
        o = run_stage(4, P, x, pos, keep, packed, R, eps)
        e = dict(hid=rel(o["hid"], R["hid"]), x2=rel(o["x2"], R["x2"]))
        print("stage 3:", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < 1e-4
    if args.stage in ("layer", "all"):
        o = run_stage(7, P, x, pos, keep, packed, R, eps)
        e = {k: rel(o[k], R[k]) for k in ("qkv", "rk", "rv", "A", "heads", "x1", "hid", "x2")}
        print("layer  :", {k: f"{v:.2e}" for k, v in e.items()})
        ok &= max(e.values()) < 1e-4
    if args.stage.startswith("b"):
        sel = ("1", "2", "4", "8", "31") if args.stage == "b" else (args.stage[1:],)
        ok &= check_backward(P, x, pos, keep, packed, R, eps, sel)
    print("OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
