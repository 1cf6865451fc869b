"""Pipeline timeline of the fused encoder kernels from their per-CTA clock stamps (gb200_encoder_set_trace).

    python tools/trace_fused.py [--B 8 --n 1849]
Prints, per kernel, the median / max over CTAs of every stamped milestone in microseconds after CTA entry
(clock64 at the measured SM clock) and the spread of CTA start times (globaltimer)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import debug_fused as DF                                                         # noqa: E402

NAMES = {
    1: {0: "entry", 1: "setup done (barriers, TMEM)", 2: "producer: x + 6 weight tiles issued", 3: "producer: all 12 issued",
        4: "mma: x split ready", 5: "mma: first weight tile landed", 6: "mma: projection MMAs issued",
        7: "mma: K~/V~ operands handed over", 8: "mma: token contraction issued", 9: "worker0: x tile landed",
        10: "worker0: split done", 11: "worker0: Q accum ready", 12: "worker0: Q chunks stored",
        13: "worker0: all projection MMAs complete", 14: "worker0: K,V chunks done, operands written",
        15: "worker0: contraction complete", 16: "worker0: partials written", 17: "exit",
        21: "worker0: K chunk 0 done", 22: "worker0: K chunk 1 done", 23: "worker0: V chunk 0 done", 24: "worker0: V chunk 1 done",
        20: "worker0: K,V loops done"},
    2: {0: "entry", 1: "setup done", 4: "mma: Q split + B_a ready", 7: "mma: heads operand handed over", 8: "mma: fc MMAs issued",
        9: "worker0: start", 10: "worker0: partials reduced", 11: "worker0: B_a written", 12: "worker0: Q tile landed",
        13: "worker0: Q split done", 14: "worker0: heads accum ready", 15: "worker0: heads stored + operand written",
        16: "worker0: fc accum ready", 17: "worker0: x1 stored", 18: "exit"},
    4: {0: "entry", 1: "setup done", 4: "mma: x1 split ready", 5: "mma: lr1 MMAs issued", 20: "mma: hidden chunk 0 handed",
        21: "mma: hidden chunk 1", 22: "mma: hidden chunk 2", 23: "mma: hidden chunk 3", 8: "mma: lr2 MMAs issued",
        9: "worker0: x1 landed", 10: "worker0: split done", 11: "worker(hf0): hidden block 0 ready",
        12: "worker(hf1): hidden block 1 ready", 13: "worker(hf0): hidden epilogue done", 14: "worker(hf1): hidden epilogue done",
        15: "worker0: y accum ready", 16: "worker0: x2 stored", 17: "exit"},
}


BNAMES = {
    1: {0: "entry", 9: "worker0: dy landed", 10: "worker0: mask+split done", 4: "mma: g2 ready", 5: "mma: g1_pre MMAs issued",
        11: "worker0: g1_pre block ready", 13: "worker0: g1 epilogue done", 8: "mma: dx1 MMAs issued", 15: "worker0: dx1 accum ready",
        16: "worker0: dx1 + partials stored", 17: "exit"},
    2: {0: "entry", 9: "worker0: dx1 landed", 10: "worker0: g_fc split done", 11: "worker0: Q split + pos block done",
        4: "mma: g_fc ready", 14: "worker0: dheads accum ready", 15: "worker0: A^T operand + dheads operand written",
        7: "mma: operands handed over", 8: "mma: dQ/G/Gp issued", 16: "worker0: all accum ready", 17: "worker0: stored", 18: "exit"},
    4: {0: "entry", 10: "worker0: G partials reduced", 11: "worker0: operands built + K,V split", 4: "mma: operands ready",
        8: "mma: issued", 14: "worker0: accum ready", 17: "worker0: LN bwd + stores done", 18: "exit"},
    8: {0: "entry", 11: "worker0: 3 blocks split", 8: "mma: all issued", 14: "worker0: accum ready", 17: "worker0: stored", 18: "exit"},
}


def trace_backward(args, P, x, pos, keep, packed, lib, grid):
    B, n, p = args.B, args.n, 2
    T, d = B * n, 32 + p
    R = dict(qkv=torch.randn(T, 384, device="cuda"), rk=torch.rand(T, 4, device="cuda"), rv=torch.rand(T, 4, device="cuda"),
             A=torch.randn(B, 4, d, d, device="cuda"), hid=torch.randn(T, 256, device="cuda"))
    RB = dict(dx1=torch.randn(T, 128, device="cuda"), dqkv=torch.randn(T, 384, device="cuda"), Graw=torch.randn(B, 4, d, d, device="cuda"))
    dy = torch.randn(B, n, 128, device="cuda")
    buf = torch.zeros(grid * 32, dtype=torch.int64, device="cuda")
    for stage, name in ((1, "enc_ffn_bwd_kernel"), (2, "enc_attn_bwd_kernel"), (4, "enc_kv_bwd_kernel"), (8, "enc_dx_kernel")):
        for _ in range(2):
            DF.run_bwd_stage(stage, P, x, pos, keep, packed, R, RB, dy, 1e-6)
        buf.zero_()
        torch.cuda.synchronize()
        lib.gb200_encoder_bwd_set_trace(buf.data_ptr())
        DF.run_bwd_stage(stage, P, x, pos, keep, packed, R, RB, dy, 1e-6)
        lib.gb200_encoder_bwd_set_trace(None)
        t = buf.view(grid, 32).cpu()
        print(f"== {name}: {grid} CTAs (inputs L2-warm)")
        rel = (t - t[:, :1]).double() / args.mhz
        for slot, label in sorted(BNAMES[stage].items(), key=lambda kv: rel[:, kv[0]].median().item()):
            col = rel[:, slot]
            if (t[:, slot] == 0).all():
                continue
            print(f"   {col.median().item():8.2f} us (max {col.max().item():8.2f})  [{slot:2d}] {label}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--n", type=int, default=1849)
    ap.add_argument("--mhz", type=float, default=1965.0)
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--warm", action="store_true", help="do not flush L2 before the traced launch")
    args = ap.parse_args()
    B, n, p = args.B, args.n, 2
    eps = 1e-6
    P, x, pos, keep = DF.make(B, n, p)
    T, d = B * n, 32 + p
    R = dict(qkv=torch.randn(T, 384, device='cuda'), x1=torch.randn(T, 128, device='cuda'),
             Araw=torch.randn(B, 4, d, d, device='cuda'))      # finite stand-ins: only timing matters here
    packed = DF.do_pack(P, p)
    lib = DF._lib.load()
    grid = B * ((n + 127) // 128)
    buf = torch.zeros(grid * 32, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    if args.bwd:
        return trace_backward(args, P, x, pos, keep, packed, lib, grid)
    for stage, name in ((1, "enc_qkv_kernel"), (2, "enc_attn_kernel"), (4, "enc_ffn_kernel")):
        for _ in range(2):
            DF.run_stage(stage, P, x, pos, keep, packed, R, eps)
        if not args.warm:
            flush.fill_(1.0)
        buf.zero_()
        torch.cuda.synchronize()
        lib.gb200_encoder_set_trace(buf.data_ptr())
        DF.run_stage(stage, P, x, pos, keep, packed, R, eps)
        lib.gb200_encoder_set_trace(None)
        t = buf.view(grid, 32).cpu()
        gt = t[:, 31]
        print(f"== {name}: {grid} CTAs; CTA start spread {(gt.max() - gt.min()).item() / 1e3:.2f} us")
        rel = (t - t[:, :1]).double() / args.mhz
        for slot, label in sorted(NAMES[stage].items(), key=lambda kv: rel[:, kv[0]].median().item()):
            col = rel[:, slot]
            if (t[:, slot] == 0).all():
                continue
            print(f"   {col.median().item():8.2f} us (max {col.max().item():8.2f})  [{slot:2d}] {label}")


if __name__ == "__main__":
    main()
