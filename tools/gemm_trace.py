"""Cycle-stamp timeline of CTA 0 of one GEMM launch (capdec_gemm_debug_trace): set-up, per-tile cadence of the MMA issuer
and of epilogue warp 4, and the tail.  Usage: python tools/gemm_trace.py [shape] [engine] [rows]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import capdec_b200 as cb  # noqa: E402
from tools.gemm_probe import SHAPES  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "qkv"
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    M, N, K, am, bm, acc = SHAPES[name]
    if len(sys.argv) > 3:
        M = int(sys.argv[3])
    pad = lambda n: (n + 127) // 128 * 128
    A = torch.randn((K, pad(M)) if am else (M, pad(K)), device="cuda")[:, : (M if am else K)]
    B = torch.randn((K, pad(N)) if bm else (N, pad(K)), device="cuda")[:, : (N if bm else K)]
    C = torch.zeros(M, pad(N), device="cuda")[:, :N]
    bias = None if acc else torch.zeros(N, device="cuda")
    lib = cb._lib.load()
    lib.capdec_gemm_debug_force_pair(mode)
    for _ in range(5):
        cb.ops.gemm(A, am, B, bm, C, M, N, K, bias=bias, accumulate=bool(acc))
    torch.cuda.synchronize()
    tr = torch.zeros(4 * 64, dtype=torch.int64, device="cuda")
    lib.capdec_gemm_debug_trace(tr.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cb.ops.gemm(A, am, B, bm, C, M, N, K, bias=bias, accumulate=bool(acc))
    e1.record()
    torch.cuda.synchronize()
    lib.capdec_gemm_debug_trace(None)
    us = e0.elapsed_time(e1) * 1e3
    t = tr.view(4, 64).cpu()
    t0 = int(t[0, 0])
    rel = lambda x: (int(x) - t0) if int(x) else None
    span = rel(t[0, 3])
    ghz = span / (us * 1e3) if span else float("nan")
    f = lambda x: "-" if x is None else f"{x / 1e3:7.2f}k"
    print(f"# {name} {M}x{N}x{K} engine {mode}: {us:.1f} us by CUDA events, CTA 0 lives {span} clk (~{ghz:.2f} GHz if it spans the launch)")
    print(f"set-up done {f(rel(t[0, 1]))} | stores drained {f(rel(t[0, 2]))} | closing sync {f(rel(t[0, 3]))}")
    print(f"warp1 entry {f(rel(t[0, 6]))} | barriers initialised {f(rel(t[0, 7]))} | producer enters role {f(rel(t[0, 4]))} | first tile decoded {f(rel(t[0, 5]))}")
    print("first k-blocks requested: " + " ".join(f(rel(t[0, 8 + i])) for i in range(8)))
    print("first k-blocks landed   : " + " ".join(f(rel(t[0, 16 + i])) for i in range(8)))
    print("| tile | producer starts | MMA: acc free | MMA: last k-block issued | epi: acc complete | epi: last TMEM read | epi: last chunk stored |")
    print("|---|---|---|---|---|---|---|")
    for i in range(21):
        if not int(t[2, 2 * i]):
            break
        print(f"| {i} | {f(rel(t[1, i]))} | {f(rel(t[2, 2 * i]))} | {f(rel(t[2, 2 * i + 1]))} | {f(rel(t[3, 3 * i]))} | {f(rel(t[3, 3 * i + 1]))} | {f(rel(t[3, 3 * i + 2]))} |")
    print("last tile, epilogue warp 4, per 32-column chunk: in registers | staged | fenced | store issued")
    for c in range(8):
        print(f"  chunk {c}: " + " | ".join(f(rel(t[3, 32 + 4 * c + i])) for i in range(4)))


if __name__ == "__main__":
    main()

