"""Stand-alone timings of the SpectralConv2d stages at the C3 decoder shape (B=8, n=141, C=32, 12 modes): CUDA events around each
launch, L2 flushed between launches.  Diagnostics only -- and coarse: the event pair also brackets the host-side launch gap
of the Python wrapper (~15-30 us), which dominates for kernels shorter than that; use tools/prof_step.py (CUPTI kernel
durations) or ncu for per-kernel times.

    python tools/time_spectral.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import galerkin_transformer_b200 as G                                             # noqa: E402
from galerkin_transformer_b200 import functional as GF                            # noqa: E402


def timed(fn, flush, reps=15):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    G.set_precision("x3")
    dev = torch.device("cuda", 0)
    B, n, C, m = 8, 141, 32, 12
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    twY, twX = GF._twiddles(n, m, dev, True)
    x = torch.randn(B, n, n, C, device=dev)
    wm = torch.randn(C, C, device=dev)
    bl = torch.randn(C, device=dev)
    T1 = GF._ydft(x, B * n, n, C, m, twY, 1.0, False)
    Xf = GF._xdft(T1, B, n, m, C, twX, 1.0 / n, False)
    Z = GF._xdft(Xf, B, n, m, C, twX, 1.0, True)
    rows = [("ydft", lambda: GF._ydft(x, B * n, n, C, m, twY, 1.0, False), 4.0 * x.numel()),
            ("xdft", lambda: GF._xdft(T1, B, n, m, C, twX, 1.0 / n, False), 4.0 * T1.numel()),
            ("xidft", lambda: GF._xdft(Xf, B, n, m, C, twX, 1.0, True), 4.0 * Z.numel()),
            ("yidft+epilogue (z out)", lambda: GF._yidft_epi(Z, B * n, n, m, C, twY, 1.0 / n, True, x, C, wm, bl, 2, True),
             12.0 * x.numel()),
            ("yidft+epilogue (bwd)", lambda: GF._yidft_epi(Z, B * n, n, m, C, twY, 1.0, False, x, C, wm, None, 0, False),
             8.0 * x.numel())]
    for name, fn, by in rows:
        t = timed(fn, flush)
        print(f"{name:28s} {t:7.1f} us   {by / t / 1e3:7.0f} GB/s of its main tensor traffic")


if __name__ == "__main__":
    main()
