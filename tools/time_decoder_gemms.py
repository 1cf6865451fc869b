"""Stand-alone timings of the tall-skinny decoder GEMMs of C3 (M = 8*141*141 tokens) in the active precision mode:
CUDA events around each launch, L2 flushed between launches.  Diagnostics only (the event pair includes the host-side launch
gap of the Python wrapper; these GEMMs are long enough, 35-60 us, for that not to matter much).

    python tools/time_decoder_gemms.py [--precision x3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import galerkin_transformer_b200 as G                                             # noqa: E402
from galerkin_transformer_b200 import functional as GF                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="x3")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    G.set_precision(args.precision)
    dev = torch.device("cuda", 0)
    M = 8 * 141 * 141
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [("fc fwd      K=130(ld 132) N=32 ", 132, 130, 32, False),
              ("mlp2 fwd    K=32  N=128 silu  ", 32, 32, 128, False),
              ("mlp2 dx     K=128 N=32        ", 128, 128, 32, True),
              ("fc dx       K=32  N=128       ", 32, 32, 128, True),
              ("square      K=128 N=128       ", 128, 128, 128, False)]
    for name, lda, K, N, wt in shapes:
        A = torch.randn(M, lda, device=dev)
        W = torch.randn(K, N, device=dev) if wt else torch.randn(N, K + (-K) % 4, device=dev)
        C = torch.empty(M, N, device=dev)
        ts = []
        for i in range(args.reps + 3):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if wt:
                GF.gemm(A, W, C, M, N, K, lda=lda, ldb=N, ldc=N, transB=False)
            else:
                GF.gemm(A, W, C, M, N, K, lda=lda, ldb=W.shape[1], ldc=N, transB=True)
            e.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        by = 4.0 * M * (K + N)
        print(f"{name} {ts[len(ts) // 2]:7.1f} us  min {ts[0]:7.1f}   {by / ts[len(ts) // 2] / 1e3:7.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
