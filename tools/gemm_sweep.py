"""Time the trunk GEMM shapes of the C2 step at the dense (12 800 rows) and the packed (8 820 rows) extent, per engine, with
L2-cold operands (6 rotating operand sets), under whatever CAPDEC_GEMM_* bring-up switches the environment carries.
Prints one markdown row per (shape, engine).  Usage: python tools/gemm_sweep.py [rows ...] [--modes 1,3] [--only qkv,fc]"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import capdec_b200 as cb  # noqa: E402

# name: (N, K, a_major, b_major) with M = rows;  wgrad shapes: (M, N, a_major, b_major) with K = rows
FWD = {"qkv": (2304, 768, 0, 1), "attn_proj": (768, 768, 0, 1), "fc": (3072, 768, 0, 1), "fc_proj": (768, 3072, 0, 1),
       "qkv_dgrad": (768, 2304, 0, 0), "fc_dgrad": (768, 3072, 0, 0), "fcproj_dgrad": (3072, 768, 0, 0)}
WGRAD = {"qkv_wgrad": (768, 2304, 1, 1), "fc_wgrad": (768, 3072, 1, 1), "fcproj_wgrad": (3072, 768, 1, 1)}


def time_gemm(M, N, K, am, bm, acc, mode, iters=18, sets=6):
    pad = lambda n: (n + 127) // 128 * 128
    ops = []
    for _ in range(sets):
        A = torch.randn((K, pad(M)) if am else (M, pad(K)), device="cuda")[:, : (M if am else K)]
        B = torch.randn((K, pad(N)) if bm else (N, pad(K)), device="cuda")[:, : (N if bm else K)]
        C = torch.zeros(M, pad(N), device="cuda")[:, :N]
        ops.append((A, B, C))
    bias = None if acc else torch.zeros(N, device="cuda")
    cb._lib.load().capdec_gemm_debug_force_pair(mode)
    try:
        for i in range(sets):
            A, B, C = ops[i]
            cb.ops.gemm(A, am, B, bm, C, M, N, K, bias=bias, accumulate=bool(acc))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            A, B, C = ops[i % sets]
            cb.ops.gemm(A, am, B, bm, C, M, N, K, bias=bias, accumulate=bool(acc))
        e1.record()
        torch.cuda.synchronize()
    finally:
        cb._lib.load().capdec_gemm_debug_force_pair(-1)
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rows", nargs="*", type=int, default=[12800, 8820])
    ap.add_argument("--modes", default="1,3")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    modes = [int(x) for x in a.modes.split(",")]
    only = set(a.only.split(",")) if a.only else None
    env = {k: v for k, v in os.environ.items() if k.startswith("CAPDEC_GEMM")}
    print(f"# gemm sweep, env {env}")
    print("| shape | M x N x K | " + " | ".join(f"engine {m} us (TF/s)" for m in modes) + " |")
    print("|---|---|" + "---:|" * len(modes))
    for rows in a.rows:
        for name, (N, K, am, bm) in FWD.items():
            if only and name not in only:
                continue
            cells = []
            for m in modes:
                try:
                    us = time_gemm(rows, N, K, am, bm, 0, m)
                    cells.append(f"{us:.1f} ({2.0 * rows * N * K / us / 1e6:.0f})")
                except Exception as ex:   # an engine that is illegal for the shape
                    cells.append("n/a")
            print(f"| {name} | {rows}x{N}x{K} | " + " | ".join(cells) + " |", flush=True)
        for name, (M, N, am, bm) in WGRAD.items():
            if only and name not in only:
                continue
            cells = []
            for m in modes:
                try:
                    us = time_gemm(M, N, rows, am, bm, 1, m)
                    cells.append(f"{us:.1f} ({2.0 * M * N * rows / us / 1e6:.0f})")
                except Exception:
                    cells.append("n/a")
            print(f"| {name} | {M}x{N}x{rows} | " + " | ".join(cells) + " |", flush=True)


if __name__ == "__main__":
    main()
