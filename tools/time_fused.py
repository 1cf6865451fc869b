"""Per-kernel device times of the fused encoder forward at C3 size (B 8, n 1849) next to the per-operator forward.

    python tools/time_fused.py [--iters 30]
Each launch is queued behind a spin kernel and bracketed by CUDA events; a 256 MiB write flushes L2 between iterations."""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import debug_fused as DF                                                         # noqa: E402
import galerkin_transformer_b200 as G                                            # noqa: E402
from galerkin_transformer_b200 import functional as GF                           # noqa: E402


def timed(fn, iters, flush):
    ts = []
    for i in range(iters):
        flush.fill_(float(i))
        torch.cuda._sleep(200000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--n", type=int, default=1849)
    args = ap.parse_args()
    B, n, p = args.B, args.n, 2
    eps = 1e-6
    P, x, pos, keep = DF.make(B, n, p)
    T, d = B * n, 32 + p
    R = dict(qkv=torch.randn(T, 384, device='cuda'), x1=torch.randn(T, 128, device='cuda'),
             Araw=torch.randn(B, 4, d, d, device='cuda'))      # finite stand-ins: only timing matters here
    packed = DF.do_pack(P, p)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for stage, name in ((1, "enc_qkv_kernel"), (2, "enc_attn_kernel"), (4, "enc_ffn_kernel"), (7, "layer (3 kernels)")):
        DF.run_stage(stage, P, x, pos, keep, packed, R, eps)
        lib = DF._lib.load()
        # re-issue the same launch without the allocation / sync of run_stage
        import ctypes
        T = B * n
        d = 32 + p
        tiles = (n + 127) // 128
        f32 = dict(dtype=torch.float32, device="cuda")
        qkv = R["qkv"].float().contiguous()
        rk, rv = torch.zeros(T, 4, **f32), torch.zeros(T, 4, **f32)
        A = torch.zeros(B, 4, d, d, **f32)
        heads = torch.zeros(T, 4 * d, **f32)
        x1 = R["x1"].float().contiguous()
        hid = torch.zeros(T, 256, **f32)
        x2 = torch.zeros(T, 128, **f32)
        ws = torch.zeros(B, tiles, 4, d, d, **f32)
        st = torch.cuda.current_stream().cuda_stream

        def launch():
            lib.gb200_encoder_layer_fwd(0, packed.data_ptr(), 128, 4, p, 256, x.data_ptr(), pos.data_ptr(), B, n, 1, eps,
                                        1.0 / n, keep.data_ptr(), 0.0, 0, 0.05, 11, 1.0, 0.05, 22, 0.05, 33,
                                        qkv.data_ptr(), rk.data_ptr(), rv.data_ptr(), A.data_ptr(), heads.data_ptr(),
                                        x1.data_ptr(), hid.data_ptr(), x2.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                        stage, st)
        for _ in range(3):
            launch()
        med, best = timed(launch, args.iters, flush)
        print(f"{name:24s} median {med:8.1f} us   best {best:8.1f} us")

    S = DF.params_struct(P, p)
    def pack():
        DF._lib.load().gb200_encoder_pack(0, ctypes.byref(S), packed.data_ptr(), torch.cuda.current_stream().cuda_stream)
    med, best = timed(pack, args.iters, flush)
    print(f"{'enc_pack_kernel':24s} median {med:8.1f} us   best {best:8.1f} us")

    # per-operator forward of the same layer (tf32 mode), inside a CUDA graph to exclude host launch gaps
    for mode in ("tf32", "x3"):
        G.set_precision(mode)
        m = G.SimpleTransformerEncoderLayer(d_model=128, n_head=4, pos_dim=2, dim_feedforward=256,
                                            attention_type="galerkin", layer_norm=False, attn_norm=True, norm_eps=1e-7,
                                            dropout=0.05, ffn_dropout=0.05).cuda()
        xx = x.clone()
        with torch.no_grad():
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                for _ in range(3):
                    m(xx, pos)
            torch.cuda.current_stream().wait_stream(s_)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y = m(xx, pos)
            med, best = timed(g.replay, args.iters, flush)
        print(f"{'module forward ' + mode:24s} median {med:8.1f} us   best {best:8.1f} us   (CUDA graph, no_grad)")


if __name__ == "__main__":
    main()
