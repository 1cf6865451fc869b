"""GPU: each C-ABI kernel against a plain PyTorch fp64 evaluation of the same expression
(tolerances are fp32 round-off: the SIMT path does exact-fp32 FMAs)."""
import math

import pytest
import torch

from galerkin_transformer_b200 import _lib, functional as GF
from helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-6
TF32_TOL = 2e-3     # tcgen05 kind::tf32: 10-bit mantissa operands, fp32 accumulate


@pytest.fixture(autouse=True)
def _exact_fp32_by_default():
    """Kernel tests check the exact-fp32 path unless they opt into the tensor-core path."""
    GF.set_precision("fp32")
    yield
    GF.set_precision("x3")


def rn(*s, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(s))
    return torch.randn(*s, generator=g).to(DEV)


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (64, 64, 16), (130, 70, 33), (257, 129, 130),
                                   (1000, 384, 128), (3, 128, 1030)])
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_layouts(M, N, K, tA, tB):
    A = rn(K, M) if tA else rn(M, K)
    B = rn(N, K, seed=1) if tB else rn(K, N, seed=1)
    C = torch.empty(M, N, device=DEV)
    GF.gemm(A, B, C, M, N, K, lda=A.shape[1], ldb=B.shape[1], ldc=N, transA=tA, transB=tB)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    assert rel_l2(C, ref) < TOL


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 32, 64), (200, 64, 40), (1000, 384, 128), (14792, 128, 136),
                                   (300, 136, 128), (97, 24, 200), (4096, 256, 128)])
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tc_layouts(M, N, K, tA, tB):
    """tcgen05 TF32 GEMM, all four operand majors, ragged M/N/K (TMA zero fill)."""
    GF.set_precision("tf32")
    Ms, Ks, Ns = -(-M // 4) * 4, -(-K // 4) * 4, -(-N // 4) * 4      # storage pitches: multiples of 4 floats
    A = rn(K, Ms)[:, :M] if tA else rn(M, Ks)[:, :K]
    B = rn(N, Ks, seed=1)[:, :K] if tB else rn(K, Ns, seed=1)[:, :N]
    C = torch.zeros(M, N, device=DEV)
    lib = _lib.load()
    assert lib.gb200_gemm_tc_supported(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K)
    GF.Profiler.reset(); GF.Profiler.enabled = True
    GF.gemm(A, B, C, M, N, K, lda=A.stride(0), ldb=B.stride(0), ldc=N, transA=tA, transB=tB)
    GF.Profiler.enabled = False
    assert GF.Profiler.records[-1][0].startswith("gemm_tc_"), "tensor-core path was not taken"
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    assert rel_l2(C, ref) < TF32_TOL, rel_l2(C, ref)


@pytest.mark.parametrize("ksplit", [1, 4, 29])
def test_gemm_tc_splitk_epilogue(ksplit):
    GF.set_precision("tf32")
    M, N, K = 384, 128, 14792                         # the d(W_qkv) shape of the C3 encoder
    A, B, bias, R = rn(K, M), rn(K, N, seed=1), rn(N, seed=2), rn(M, N, seed=3)
    outs = []
    for _ in range(2):
        C = torch.empty(M, N, device=DEV)
        Z = torch.empty(M, N, device=DEV)
        GF.gemm(A, B, C, M, N, K, lda=M, ldb=N, ldc=N, transA=True, alpha=0.5, bias=bias, act=2, zout=Z, ldz=N,
                residual=R, ldr=N, rscale=-1.0, ksplit=ksplit)
        outs.append(C)
    z = 0.5 * A.double().t() @ B.double() + bias.double()
    assert rel_l2(Z, z) < TF32_TOL
    assert rel_l2(outs[0], R.double() - torch.nn.functional.silu(z)) < TF32_TOL
    assert torch.equal(outs[0], outs[1])             # fixed-order split-K reduction


@pytest.mark.parametrize("act", ["silu", "relu"])
def test_linear_autograd_tensor_cores(act):
    """TF32 forward <= 2e-3, gradients <= 5e-3.  For ReLU the reference gate is taken from the
    kernel's own output: a TF32-perturbed pre-activation flips the sign of ~4e-4 of the entries
    near zero, which alone moves gradients by sqrt(4e-4) ~ 2e-2 in any reduced-precision
    implementation (cuBLAS TF32 included) and says nothing about the kernel."""
    GF.set_precision("tf32")
    x = rn(8, 1849, 128).requires_grad_(True)
    W = (0.1 * rn(256, 128, seed=1)).requires_grad_(True)
    b = rn(256, seed=2).requires_grad_(True)
    y = GF.linear(x, W, b, act=act)
    cot = rn(8, 1849, 256, seed=4)
    grads = torch.autograd.grad((y * cot).sum(), [x, W, b])
    xd, Wd, bd = [t.detach().double().requires_grad_(True) for t in (x, W, b)]
    z = xd @ Wd.t() + bd
    yr = torch.nn.functional.silu(z) if act == "silu" else z * (y.detach() > 0).double()
    gr = torch.autograd.grad((yr * cot.double()).sum(), [xd, Wd, bd])
    assert rel_l2(y, yr) < TF32_TOL
    for g, r in zip(grads, gr):
        assert rel_l2(g, r) < 5e-3


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("act,p1,p2,shortcut,dims", [
    ("relu", 0.0, 0.0, True, (1849, 128, 256, 128)), ("relu", 0.2, 0.1, True, (1849, 128, 256, 128)),
    ("silu", 0.2, 0.1, True, (777, 64, 136, 64)), ("silu", 0.0, 0.0, False, (5000, 32, 128, 1)),
    ("relu", 0.3, 0.0, False, (5000, 32, 128, 3)), ("none", 0.2, 0.0, False, (300, 24, 40, 24))])
def test_mlp2_matches_two_linear_nodes(precision, act, p1, p2, shortcut, dims):
    """The fused Linear-act-dropout-Linear node (gated backward GEMM, shortcut gradient in the last GEMM's epilogue,
    weight gradients on side streams) against the same computation as two `linear` nodes with the same Philox
    keys: same forward bits, gradients to round-off."""
    GF.set_precision(precision)
    M, K, N1, N2 = dims
    x = rn(M, K).requires_grad_(True)
    W1 = (0.2 * rn(N1, K, seed=1)).requires_grad_(True)
    b1 = rn(N1, seed=2).requires_grad_(True)
    W2 = (0.2 * rn(N2, N1, seed=3)).requires_grad_(True)
    b2 = rn(N2, seed=4).requires_grad_(True)
    cot = rn(M, N2, seed=5)
    leaves = [x, W1, b1, W2, b2]
    torch.manual_seed(11); GF._seed_counter = 0
    h = GF.linear(x, W1, b1, act=act, drop_p=p1)
    ya = GF.linear(h, W2, b2, residual=x if shortcut else None, rscale=-1.0 if shortcut else 1.0, drop_p=p2)
    ga = torch.autograd.grad((ya * cot).sum(), leaves)
    torch.manual_seed(11); GF._seed_counter = 0
    yb = GF.mlp2(x, W1, b1, W2, b2, act=act, drop_p1=p1, drop_p2=p2, rscale=-1.0 if shortcut else 1.0,
                 shortcut=shortcut)
    gb = torch.autograd.grad((yb * cot).sum(), leaves)
    assert rel_l2(yb, ya) < 1e-6
    for a_, b_ in zip(gb, ga):
        assert rel_l2(b_, a_) < 2e-5, rel_l2(b_, a_)
    if p1 == 0.0 and p2 == 0.0 and precision == "fp32":      # and against plain fp64 math
        xd, W1d, b1d, W2d, b2d = [t.detach().double().requires_grad_(True) for t in leaves]
        z = xd @ W1d.t() + b1d
        hd = {"none": z, "relu": torch.relu(z), "silu": torch.nn.functional.silu(z)}[act]
        yd = hd @ W2d.t() + b2d
        yd = xd - yd if shortcut else yd
        gd = torch.autograd.grad((yd * cot.double()).sum(), [xd, W1d, b1d, W2d, b2d])
        assert rel_l2(yb, yd) < TOL
        for a_, b_ in zip(gb, gd):
            assert rel_l2(a_, b_) < 1e-5


@pytest.mark.parametrize("gate_act", ["relu", "silu"])
@pytest.mark.parametrize("M,N,K,tc", [(1000, 256, 128, True), (14792, 256, 128, True), (300, 40, 24, False),
                                      (5000, 128, 1, False), (130, 72, 33, True)])
def test_gated_gemm(gate_act, M, N, K, tc):
    """C = rscale * dropout((A.B) * act'(G)): tcgen05 float4 / scalar epilogues, SIMT and the rank-1 streaming kernel."""
    GF.set_precision("tf32" if tc else "fp32")
    Ks = -(-K // 4) * 4 if tc else K
    A = rn(M, Ks)[:, :K]
    B = rn(K, N, seed=1)
    G = rn(M, N, seed=2)
    C = torch.empty(M, N, device=DEV)
    p, seed = 0.25, 1234
    GF.gemm(A, B, C, M, N, K, lda=A.stride(0), ldb=N, ldc=N, drop_p=p, seed=seed, rscale=-0.5, gate=G, ldg=N,
            gate_act=GF.ACT[gate_act])
    mask = torch.ones(M, N, device=DEV)
    _lib.check(_lib.load().gb200_philox_scale(0, mask.data_ptr(), M * N, p, seed,
                                              torch.cuda.current_stream().cuda_stream), "philox")
    d = G.double()
    if gate_act == "relu":
        gd = (d > 0).double()
    else:
        s = torch.sigmoid(d)
        gd = s * (1 + d * (1 - s))
    ref = -0.5 * (A.double() @ B.double()) * gd * mask.double()
    assert rel_l2(C, ref) < (TF32_TOL if tc else 5e-6), rel_l2(C, ref)


@pytest.mark.parametrize("ksplit", [1, 3, 16])
def test_gemm_splitk_epilogue_is_deterministic(ksplit):
    M, N, K = 96, 80, 4000
    A, B, bias, R = rn(K, M), rn(K, N, seed=1), rn(N, seed=2), rn(M, N, seed=3)
    outs = []
    for _ in range(2):
        C = torch.empty(M, N, device=DEV)
        Z = torch.empty(M, N, device=DEV)
        GF.gemm(A, B, C, M, N, K, lda=M, ldb=N, ldc=N, transA=True, alpha=0.5, bias=bias, act=2, zout=Z,
                ldz=N, residual=R, ldr=N, rscale=-1.0, ksplit=ksplit)
        outs.append(C)
    z = 0.5 * A.double().t() @ B.double() + bias.double()
    ref = R.double() - torch.nn.functional.silu(z)
    assert rel_l2(outs[0], ref) < 5e-6 and rel_l2(Z, z) < 5e-6
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("M,N,K", [(20000, 1, 128), (20000, 2, 130), (5000, 7, 33), (20000, 128, 1), (20000, 32, 2),
                                   (1, 128, 20000), (4, 64, 20000), (32, 2, 20000), (3, 5, 9000), (30, 34, 9001)])
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_degenerate_shapes(M, N, K, tA, tB):
    """N = 1 output head, K = 2 grid columns, and their weight gradients (streaming kernels in gemm_simt.cu),
    with the full epilogue; every layout must agree whether or not it takes a streaming path."""
    A = rn(K, M) if tA else rn(M, K)
    B = rn(N, K, seed=1) if tB else rn(K, N, seed=1)
    bias, R = rn(N, seed=2), rn(M, N, seed=3)
    C0 = rn(M, N, seed=4)
    C = C0.clone()
    Z = torch.empty(M, N, device=DEV)
    GF.gemm(A, B, C, M, N, K, lda=A.shape[1], ldb=B.shape[1], ldc=N, transA=tA, transB=tB, alpha=0.5, bias=bias,
            act=2, zout=Z, ldz=N, residual=R, ldr=N, rscale=-1.0, accumulate=True)
    z = 0.5 * (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double()) + bias.double()
    ref = C0.double() + R.double() - torch.nn.functional.silu(z)
    assert rel_l2(Z, z) < 5e-6 and rel_l2(C, ref) < 5e-6
    C2 = C0.clone()
    GF.gemm(A, B, C2, M, N, K, lda=A.shape[1], ldb=B.shape[1], ldc=N, transA=tA, transB=tB, alpha=0.5, bias=bias,
            act=2, zout=Z, ldz=N, residual=R, ldr=N, rscale=-1.0, accumulate=True)
    assert torch.equal(C, C2)                                # deterministic reductions


def test_gemm_accumulate_and_offsets():
    M, N, K1, K2 = 50, 20, 12, 2
    x1, x2, W, b = rn(M, K1), rn(M, K2, seed=1), rn(N, K1 + K2, seed=2), rn(N, seed=3)
    y = GF.linear_cat(x1, x2, W, b)
    ref = torch.cat([x1, x2], -1).double() @ W.double().t() + b.double()
    assert rel_l2(y, ref) < TOL


def test_fused_dropout_statistics_and_backward_consistency():
    M, N, K, p = 512, 256, 64, 0.3
    x = rn(M, K).requires_grad_(True)
    W = rn(N, K, seed=1).requires_grad_(True)
    torch.manual_seed(3)
    y = GF.linear(x, W, None, act="relu", drop_p=p)
    dense = torch.relu(x.detach().double() @ W.detach().double().t())
    kept = (y != 0) & (dense > 0)
    frac = kept.sum().item() / (dense > 0).sum().item()
    assert abs(frac - (1 - p)) < 0.01                       # keep probability
    # the drop probability is quantised to thr / 65536 and the keep scale is its exact inverse complement (unbiased mask)
    keep_scale = 65536.0 / (65536.0 - round(p * 65536.0))
    assert rel_l2(y[kept], dense[kept] * keep_scale) < TOL  # inverted-dropout scale
    # backward regenerates the same mask: grad flows only through kept, positive entries
    y.sum().backward()
    mask = kept.double() * keep_scale
    assert rel_l2(x.grad, mask @ W.detach().double()) < 1e-5
    assert rel_l2(W.grad, mask.t() @ x.detach().double()) < 1e-5


@pytest.mark.parametrize("act", ["none", "relu", "silu"])
def test_linear_autograd(act):
    x = rn(3, 37, 24).requires_grad_(True)
    W = rn(40, 24, seed=1).requires_grad_(True)
    b = rn(40, seed=2).requires_grad_(True)
    R = rn(3, 37, 40, seed=3).requires_grad_(True)
    y = GF.linear(x, W, b, act=act, residual=R, rscale=-1.0)
    cot = rn(3, 37, 40, seed=4)
    grads = torch.autograd.grad((y * cot).sum(), [x, W, b, R])
    xd, Wd, bd, Rd = [t.detach().double().requires_grad_(True) for t in (x, W, b, R)]
    z = xd @ Wd.t() + bd
    a = {"none": z, "relu": torch.relu(z), "silu": torch.nn.functional.silu(z)}[act]
    yr = Rd - a
    gr = torch.autograd.grad((yr * cot.double()).sum(), [xd, Wd, bd, Rd])
    assert rel_l2(y, yr) < TOL
    for g, r in zip(grads, gr):
        assert rel_l2(g, r) < 1e-5


def test_layernorm_fwd_bwd():
    x = rn(1000, 48).requires_grad_(True)
    g = (1 + 0.3 * rn(48, seed=1)).requires_grad_(True)
    b = rn(48, seed=2).requires_grad_(True)
    y = GF.layer_norm(x, g, b, 1e-5)
    cot = rn(1000, 48, seed=3)
    grads = torch.autograd.grad((y * cot).sum(), [x, g, b])
    xd, gd, bd = [t.detach().double().requires_grad_(True) for t in (x, g, b)]
    yr = torch.nn.functional.layer_norm(xd, (48,), gd, bd, 1e-5)
    gr = torch.autograd.grad((yr * cot.double()).sum(), [xd, gd, bd])
    assert rel_l2(y, yr) < TOL
    for a, r in zip(grads, gr):
        assert rel_l2(a, r) < 1e-5


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("B,H,n,dk,p", [(2, 4, 49, 8, 2), (1, 1, 300, 48, 2), (3, 2, 1000, 16, 1),
                                        (2, 1, 130, 96, 1), (2, 4, 77, 32, 0), (8, 4, 1849, 32, 2),
                                        (2, 2, 513, 62, 2), (1, 3, 200, 22, 1)])
def test_attention_core_kernels(B, H, n, dk, p, tc):
    """tc=0: exact-fp32 SIMT kernels; tc=1: warp-level TF32 mma.sync kernels (d <= 64)."""
    lib = _lib.load()
    tol = TF32_TOL if tc else 5e-6
    dm, d, T = H * dk, dk + p, B * n
    qkv = rn(T, 3 * dm)
    pos = torch.rand(B, n, max(p, 1), device=DEV)[..., :p].contiguous() if p else None
    gam, bet = 1 + 0.2 * rn(H, dk, seed=1), 0.2 * rn(H, dk, seed=2)
    dev, st = 0, torch.cuda.current_stream().cuda_stream

    def aug(block, g=None, b=None):
        t = qkv[:, block * dm:(block + 1) * dm].double().view(B, n, H, dk)
        if g is not None:
            t = t * g.double() + b.double()
        t = t.permute(0, 2, 1, 3)
        if p:
            t = torch.cat([pos.double().unsqueeze(1).expand(-1, H, -1, -1), t], -1)
        return t                                             # (B,H,n,d)
    k, v, q = aug(1, gam, bet), aug(2), aug(0)
    mask = (torch.rand(B, H, d, d, device=DEV) > 0.5).to(torch.uint8)
    A = torch.empty(B, H, d, d, device=DEV)
    ops = [GF._hop(qkv, 3 * dm, i * dm, False, *((gam, bet) if i == 1 else (None, None))) for i in range(3)]
    nsplit = lib.gb200_attn_suggest_nsplit(B, H, n)
    wsb = lib.gb200_attn_xty_workspace_bytes(B, H, d, nsplit)
    ws = _lib.workspace(wsb, qkv)
    _lib.check(lib.gb200_attn_xty(dev, ops[1], ops[2], _lib.ptr(pos), B, H, n, dk, p, 1.0 / n, _lib.ptr(mask),
                                  0.0, 0, _lib.ptr(A), nsplit, _lib.ptr(ws), wsb, tc, st))
    Aref = (k.transpose(-1, -2) @ v) / n * (2.0 * mask.double())
    assert rel_l2(A, Aref) < tol
    out = torch.empty(B, n, H * d, device=DEV)
    _lib.check(lib.gb200_attn_xm(dev, ops[0], _lib.ptr(pos), _lib.ptr(A), 0, B, H, n, dk, p, _lib.ptr(out), H * d,
                                 0, 1, 1.0, tc, st))
    oref = (q @ A.double()).permute(0, 2, 1, 3).reshape(B, n, H * d)
    assert rel_l2(out, oref) < tol
    # transposed multiply, non-augmented output (gradient layout)
    dq = torch.zeros(T, 3 * dm, device=DEV)
    do = GF._hop(out, H * d, 0, True)
    _lib.check(lib.gb200_attn_xm(dev, do, _lib.ptr(pos), _lib.ptr(A), 1, B, H, n, dk, p, _lib.ptr(dq), 3 * dm, dm,
                                 0, 1.0, tc, st))
    ref = (oref.view(B, n, H, d).permute(0, 2, 1, 3) @ A.double().transpose(-1, -2))[..., p:]
    ref = ref.permute(0, 2, 1, 3).reshape(T, dm)
    assert rel_l2(dq[:, dm:2 * dm], ref) < tol
    assert dq[:, :dm].abs().max() == 0 and dq[:, 2 * dm:].abs().max() == 0


@pytest.mark.parametrize("B,Hin,Win,Hout,Wout,C", [(2, 141, 141, 77, 77, 128), (2, 77, 77, 43, 43, 128),
                                                    (2, 43, 43, 77, 77, 128), (1, 77, 77, 141, 141, 128),
                                                    (3, 29, 31, 10, 7, 6), (2, 10, 7, 29, 31, 5), (2, 9, 9, 1, 1, 4),
                                                    (2, 1, 1, 5, 6, 8), (2, 16, 16, 16, 16, 12)])
def test_interp_bilinear_matches_torch(B, Hin, Win, Hout, Wout, C):
    """F.interpolate(mode='bilinear', align_corners=True) forward and its autograd backward (libs/layers.py:493, 506,
    660, 668), channel-last streaming kernels; the C3 scaler sizes plus ragged / degenerate ones."""
    x = rn(B, Hin, Win, C).requires_grad_(True)
    y = GF.interp_bilinear(x, Hout, Wout)
    cot = rn(B, Hout, Wout, C, seed=1)
    (gx,) = torch.autograd.grad((y * cot).sum(), [x])
    xd = x.detach().double().requires_grad_(True)
    yr = torch.nn.functional.interpolate(xd.permute(0, 3, 1, 2), size=(Hout, Wout), mode="bilinear", align_corners=True)
    (gr,) = torch.autograd.grad((yr * cot.double().permute(0, 3, 1, 2)).sum(), [xd])
    # fp32 source-index arithmetic (scale * dst with indices up to 140) puts ~1e-5 of slack on the interpolation weight,
    # exactly as in ATen's own fp32 kernel; against exact fp64 weights that is a few 1e-6 in relative L2
    assert rel_l2(y, yr.permute(0, 2, 3, 1)) < 2e-5, rel_l2(y, yr.permute(0, 2, 3, 1))
    assert rel_l2(gx, gr) < 2e-5, rel_l2(gx, gr)
    # and the fp32 ATen kernel itself (same float source-index arithmetic): agreement to the last bits
    y32 = torch.nn.functional.interpolate(x.detach().permute(0, 3, 1, 2), size=(Hout, Wout), mode="bilinear",
                                          align_corners=True).permute(0, 2, 3, 1)
    assert (y - y32).abs().max() <= 2e-6 * max(1.0, float(y32.abs().max())), float((y - y32).abs().max())


@pytest.mark.parametrize("T,H,dk", [(777, 4, 32), (500, 4, 12), (1000, 1, 96), (333, 2, 16)])
def test_headnorm_fwd_bwd(T, H, dk):
    """(4,32) and (2,16) take the coalesced float4/shuffle kernels, the others the generic ones."""
    lib = _lib.load()
    eps = 1e-7
    dm = H * dk
    raw = rn(T, 3 * dm)
    gam = 1 + 0.2 * rn(H, dk, seed=1)
    buf = raw.clone()
    rstd = torch.empty(T, H, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.gb200_headnorm_fwd(0, _lib.ptr(buf), 3 * dm, dm, -1, T, H, dk, eps, _lib.ptr(rstd), None, st))
    x = raw[:, dm:2 * dm].double().view(T, H, dk).requires_grad_(True)
    mu, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
    xhat = (x - mu) / torch.sqrt(var + eps)
    assert rel_l2(buf[:, dm:2 * dm], xhat.reshape(T, dm)) < TOL
    assert torch.equal(buf[:, :dm], raw[:, :dm]) and torch.equal(buf[:, 2 * dm:], raw[:, 2 * dm:])
    dy = rn(T, 3 * dm, seed=5)
    y = xhat * gam.double()
    gx, = torch.autograd.grad((y * dy[:, dm:2 * dm].double().view(T, H, dk)).sum(), x)
    dbuf = dy.clone()
    dg, db = torch.empty(H, dk, device=DEV), torch.empty(H, dk, device=DEV)
    wsb = lib.gb200_headnorm_bwd_workspace_bytes(T, H, dk)
    ws = _lib.workspace(wsb, buf)
    _lib.check(lib.gb200_headnorm_bwd(0, _lib.ptr(dbuf), 3 * dm, dm, -1, _lib.ptr(buf), 3 * dm, dm, -1, _lib.ptr(rstd),
                                      None, _lib.ptr(gam), None, T, H, dk, _lib.ptr(dg), _lib.ptr(db), None, None, 0,
                                      _lib.ptr(ws), wsb, st))
    assert rel_l2(dbuf[:, dm:2 * dm], gx.reshape(T, dm)) < 1e-5
    dyk = dy[:, dm:2 * dm].double().view(T, H, dk)
    assert rel_l2(dg, (dyk * xhat.detach()).sum(0)) < 1e-5 and rel_l2(db, dyk.sum(0)) < 1e-5


def _sc_ref(x, wl, bl, fw0, fw1, m, act, two_d):
    """torch.fft evaluation of the spectral layer in fp64 (independent of the kernels' DFT form)."""
    xd = x.double()
    res = xd @ wl.double().t() + bl.double()
    if two_d:
        B, n = x.shape[0], x.shape[1]
        xf = torch.fft.rfft2(xd.permute(0, 3, 1, 2), s=(n, n), norm="ortho")
        of = xf.new_zeros(B, wl.shape[0], n, n // 2 + 1)
        of[:, :, :m, :m] = torch.einsum("bixy,ioxy->boxy", xf[:, :, :m, :m], torch.view_as_complex(fw0.double()))
        of[:, :, -m:, :m] = torch.einsum("bixy,ioxy->boxy", xf[:, :, -m:, :m], torch.view_as_complex(fw1.double()))
        y = torch.fft.irfft2(of, s=(n, n), norm="ortho").permute(0, 2, 3, 1)
    else:
        B, n = x.shape[0], x.shape[1]
        xf = torch.fft.rfft(xd.permute(0, 2, 1), n=n, norm="ortho")
        of = xf.new_zeros(B, wl.shape[0], n // 2 + 1)
        of[:, :, :m] = torch.einsum("bix,iox->box", xf[:, :, :m], torch.view_as_complex(fw0.double()))
        y = torch.fft.irfft(of, n=n, norm="ortho").permute(0, 2, 1)
    z = y + res
    return {"silu": torch.nn.functional.silu, "relu": torch.relu, "none": lambda t: t}[act](z)


@pytest.mark.parametrize("shape,Co,m,act", [((2, 15, 15, 6), 5, 4, "silu"), ((1, 32, 32, 20), 20, 12, "silu"),
                                            ((2, 141, 141, 8), 8, 12, "relu"), ((3, 16, 16, 4), 7, 8, "none"),
                                            ((2, 141, 141, 32), 32, 12, "silu"), ((1, 150, 150, 16), 24, 9, "silu"),
                                            ((5, 40, 40, 12), 8, 16, "none"),
                                            ((2, 64, 8), 6, 5, "silu"), ((2, 8192, 16), 12, 16, "silu"),
                                            ((3, 45, 4), 4, 7, "relu"), ((1, 40, 3), 3, 20, "none")])
@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_spectral_conv_forward_backward(shape, Co, m, act, prec):
    """fp32: exact-FMA DFT kernels (<= 5e-6 / 2e-5); tf32: the per-row twiddle products of the 2-D case run on
    TF32 tensor cores (<= 2e-3 / 5e-3)."""
    GF.set_precision(prec)
    ftol, gtol = (5e-6, 2e-5) if prec == "fp32" else (TF32_TOL, 5e-3)
    two_d = len(shape) == 4
    Ci = shape[-1]
    x = rn(*shape).requires_grad_(True)
    wl = (0.3 * rn(Co, Ci, seed=1)).requires_grad_(True)
    bl = (0.3 * rn(Co, seed=2)).requires_grad_(True)
    wshape = (Ci, Co, m, m, 2) if two_d else (Ci, Co, m, 2)
    fw0 = (0.3 * rn(*wshape, seed=3)).requires_grad_(True)
    fw1 = (0.3 * rn(*wshape, seed=4)).requires_grad_(True) if two_d else None
    y, _ = GF.spectral_conv(x, wl, bl, fw0, fw1, modes=m, act=act, two_d=two_d)
    params = [x, wl, bl, fw0] + ([fw1] if two_d else [])
    cot = rn(*y.shape, seed=9)
    grads = torch.autograd.grad((y * cot).sum(), params)
    pd = [t.detach().clone().requires_grad_(True) for t in params]
    yr = _sc_ref(pd[0], pd[1], pd[2], pd[3], pd[4] if two_d else None, m, act, two_d)
    gr = torch.autograd.grad((yr * cot.double()).sum(), pd)
    assert rel_l2(y, yr) < ftol
    if act == "relu" and prec == "tf32":
        # a reduced-precision pre-activation flips the ReLU gate of the few entries with |z| ~ 1e-3 |z|_rms; each flip
        # moves one gradient entry by O(1), so the gradient bound is looser here (and only here)
        gtol = 3e-2
    for g, r in zip(grads, gr):
        assert rel_l2(g, r) < gtol


@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(300, 128, 96), (1000, 32, 128), (257, 136, 40), (4096, 128, 32)])
def test_gemm_split_tf32_is_fp32_grade(M, N, K, tA, tB):
    """'x3' mode outside the fused kernels: tcgen05 GEMM with the two-term TF32 split (hi.hi + hi.lo + lo.hi), all four
    operand layouts, with the fused epilogue -- fp32-grade products (<= 5e-6), against 5e-4 for single-pass TF32."""
    GF.set_precision("x3")
    A = rn(K, M, seed=1) if tA else rn(M, K, seed=1)
    B = rn(N, K, seed=2) if tB else rn(K, N, seed=2)
    bias, R = rn(N, seed=3), rn(M, N, seed=4)
    C = torch.empty(M, N, device=DEV)
    GF.gemm(A, B, C, M, N, K, lda=A.shape[1], ldb=B.shape[1], ldc=N, transA=tA, transB=tB, bias=bias, act=2,
            residual=R, ldr=N, rscale=0.5)
    z = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double()) + bias.double()
    ref = R.double() + 0.5 * torch.nn.functional.silu(z)
    assert rel_l2(C, ref) < 5e-6, rel_l2(C, ref)
    Cw = torch.empty(M, N, device=DEV)
    GF.gemm(A, B, Cw, M, N, K, lda=A.shape[1], ldb=B.shape[1], ldc=N, transA=tA, transB=tB, wgrad=True)   # single-pass TF32
    plain = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    assert rel_l2(Cw, plain) < 2e-3, rel_l2(Cw, plain)      # (unaligned leading dimensions fall back to the exact kernel)


def test_fourier_quadratic_width_limit_is_a_clean_error():
    """reference ex1_burgers config (attention_type fourier, n_hidden 96, 1 head, pos_dim 1 -> d = 97): the always-on n x n
    dropout needs the quadratic kernels (d <= 64) -> NotImplementedError with a hint; 'off' runs the exact linear form."""
    import galerkin_transformer_b200 as G
    a = G.SimpleAttention(n_head=1, d_model=96, pos_dim=1, attention_type="fourier", norm=True).to(DEV)
    x, pos = rn(2, 64, 96), torch.rand(2, 64, 1, device=DEV)
    with pytest.raises(NotImplementedError, match="set_attn_dropout"):
        a(x, x, x, pos=pos)
    G.set_attn_dropout(a, "off")
    out, _ = a(x, x, x, pos=pos)
    assert out.shape == (2, 64, 96) and torch.isfinite(out).all()


def test_linear_relu_negative_rscale_and_bias_only_grads():
    """a ReLU gate must not be read from an output scaled by rscale <= 0; bias-only gradients need z for SiLU"""
    x, W, b = rn(40, 24), rn(16, 24, seed=1).requires_grad_(True), rn(16, seed=2).requires_grad_(True)
    y = GF.linear(x, W, b, act="relu", rscale=-1.0)
    y.sum().backward()
    ref_gate = ((x.double() @ W.detach().double().t() + b.detach().double()) > 0).double()
    assert rel_l2(b.grad, -ref_gate.sum(0)) < 1e-6
    b2 = rn(16, seed=3).requires_grad_(True)
    y2 = GF.linear(x, W.detach(), b2, act="silu")
    y2.sum().backward()
    z = x.double() @ W.detach().double().t() + b2.detach().double()
    sg = torch.sigmoid(z)
    assert rel_l2(b2.grad, (sg * (1 + z * (1 - sg))).sum(0)) < 1e-5
