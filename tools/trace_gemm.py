"""Per-CTA phase timeline of the tcgen05 GEMM (gb200_gemm_tc_set_trace): where does a launch spend its time?
usage: python tools/trace_gemm.py [shape-substring]"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from galerkin_transformer_b200 import functional as GF, _lib

lib = _lib.load()
T = 14792
shapes = [("qkv nt", T, 384, 128, False, True), ("lr2 nt", T, 128, 256, False, True),
          ("dx nn", T, 128, 384, False, False), ("dWqkv tn", 384, 128, T, True, False)]
only = sys.argv[1] if len(sys.argv) > 1 else None
GF.set_precision("tf32")
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
names = ["entry", "setup", "tile0", "tileL", "round", "accum", "drain", "store"]
for name, M, N, K, tA, tB in shapes:
    if only and only not in name:
        continue
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.empty(M, N, device="cuda")
    run = lambda: GF.gemm(A, B, C, M, N, K, lda=A.shape[1], ldb=B.shape[1], ldc=N, transA=tA, transB=tB)
    for _ in range(3):
        run()
    trace = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
    flush.fill_(1.0)
    torch.cuda.synchronize()
    lib.gb200_gemm_tc_set_trace(trace.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record()
    torch.cuda.synchronize()
    lib.gb200_gemm_tc_set_trace(None)
    t = trace.cpu().numpy().reshape(-1, 8)
    t = t[t[:, 0] > 0].astype(np.float64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3                               # us since the first CTA started
    print(f"== {name} M={M} N={N} K={K}: {len(t)} CTAs, event time {s.elapsed_time(e) * 1e3:.1f} us, "
          f"first entry -> last store {rel[:, 7].max():.1f} us")
    print("   CTA start (us): " + " ".join(f"p{q}={np.percentile(rel[:, 0], q):.1f}" for q in (0, 25, 50, 75, 90, 100)))
    d = np.diff(t, axis=1) / 1e3
    for i in range(7):
        print(f"   {names[i]:>6s}->{names[i + 1]:<6s} median {np.median(d[:, i]):6.2f}  p90 {np.percentile(d[:, i], 90):6.2f}  max {d[:, i].max():6.2f} us")
    life = rel[:, 7] - rel[:, 0]
    print(f"   CTA lifetime median {np.median(life):.2f}  p90 {np.percentile(life, 90):.2f}  max {life.max():.2f} us")
    first = rel[:, 0] < 1.0
    print(f"   first wave: {first.sum()} CTAs; lifetime median {np.median(life[first]):.2f};  later CTAs lifetime median "
          f"{np.median(life[~first]) if (~first).any() else float('nan'):.2f}")
