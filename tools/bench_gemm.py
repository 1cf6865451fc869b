"""Micro-benchmark of the library GEMMs on the C3 encoder/decoder shapes (CUDA events, L2 flushed
between launches) against cuBLAS (torch.matmul, TF32 on) for orientation.  Not part of bench.py."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from galerkin_transformer_b200 import functional as GF

dev = "cuda"
T = 14792
shapes = [  # name, M, N, K, transA, transB
    ("qkv  nt", T, 384, 128, False, True), ("fc   nt", T, 128, 136, False, True),
    ("lr1  nt", T, 256, 128, False, True), ("lr2  nt", T, 128, 256, False, True),
    ("dx   nn", T, 128, 384, False, False), ("dh   nn", T, 256, 128, False, False),
    ("dWqkv tn", 384, 128, T, True, False), ("dW2  tn", 128, 256, T, True, False),
    ("reg0 nt", 159048, 128, 32, False, True), ("dreg nn", 159048, 32, 128, False, False),
]
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
only = sys.argv[1] if len(sys.argv) > 1 else None
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for name, M, N, K, tA, tB in shapes:
    if only and only not in name:
        continue
    A = torch.randn((K, M) if tA else (M, K), device=dev)
    B = torch.randn((N, K) if tB else (K, N), device=dev)
    C = torch.empty(M, N, device=dev)
    res = {}
    for mode in ("tf32", "fp32", "cublas"):
        times = []
        for it in range(iters + 3):
            flush.fill_(it)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if mode == "cublas":
                torch.backends.cuda.matmul.allow_tf32 = True
                a = A.t() if tA else A
                b = B.t() if tB else B
                s.record(); torch.matmul(a, b, out=C); e.record()
            else:
                GF.set_precision(mode)
                s.record()
                GF.gemm(A, B, C, M, N, K, lda=A.shape[1], ldb=B.shape[1], ldc=N, transA=tA, transB=tB)
                e.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(s.elapsed_time(e) * 1e3)
        times.sort()
        res[mode] = times[len(times) // 2]
    fl = 2.0 * M * N * K
    by = 4.0 * (M * K + N * K + M * N)
    print(f"{name:9s} M={M:6d} N={N:4d} K={K:6d}  tc {res['tf32']:7.1f} us ({fl/res['tf32']/1e6:6.1f} TF/s, {by/res['tf32']/1e3:6.0f} GB/s)"
          f"  simt {res['fp32']:7.1f} us  cublas {res['cublas']:7.1f} us")
