"""3xTF32 (fp32-grade) GEMM probe on one B200: speed and error of the in-pipeline hi/lo split against the 1xTF32 mode,
the CUDA-core fp32 kernel and an fp64 product of the same operands (torch.float64 matmul = checker only).

    python tools/x3_probe.py            # C2 shapes: us / TFLOP/s per mode, max-abs and rel-L2 error vs fp64
    python tools/x3_probe.py ksweep     # error of one accumulation chain as a function of K (and of the split-K chain cap)
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import capdec_b200 as cb  # noqa: E402
from capdec_b200 import ops  # noqa: E402

SHAPES = {  # name: (M, N, K, a_major, b_major, accumulate)
    "qkv": (12800, 2304, 768, 0, 1, 0), "fc": (12800, 3072, 768, 0, 1, 0), "fc_proj": (12800, 768, 3072, 0, 1, 0),
    "attn_proj": (12800, 768, 768, 0, 1, 0), "qkv_dgrad": (12800, 768, 2304, 0, 0, 0),
    "qkv_wgrad": (768, 2304, 12800, 1, 1, 1), "fc_wgrad": (768, 3072, 12800, 1, 1, 1),
    "lm_head": (6144, 50257, 768, 0, 0, 0), "lm_dgrad": (6144, 768, 50257, 0, 1, 1), "lm_wgrad": (50257, 768, 6144, 1, 1, 1),
    "mlp_fc2": (256, 7680, 3840, 0, 0, 0),
}


def pad(n, q=128):
    return (n + q - 1) // q * q


def mat(rows, cols, scale=1.0):
    return (torch.randn(rows, pad(cols), device="cuda") * scale)[:, :cols]


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def ref64(A, am, B, bm):
    a = (A.t() if am else A).double()
    b = (B.t() if bm else B).double()
    return a @ b.t()


def errs(C, R):
    d = (C.double() - R)
    return d.abs().max().item() / R.abs().max().item(), (d.norm() / R.norm()).item(), (C.double().norm() / R.norm()).item() - 1.0


def shapes(only):
    torch.manual_seed(0)
    print("| shape | mode | us | TFLOP/s | max-abs/max | rel-L2 | norm bias |")
    print("|---|---|---|---|---|---|---|")
    for name in only or list(SHAPES):
        M, N, K, am, bm, acc = SHAPES[name]
        A = mat(K, M) if am else mat(M, K)
        B = mat(K, N) if bm else mat(N, K)
        R = ref64(A, am, B, bm) if M * N <= 400e6 else None
        for mode in ("tf32", "tf32x3", "fp32"):
            if mode == "fp32" and 2.0 * M * N * K > 3e11:
                continue
            C = torch.zeros(M, pad(N), device="cuda")[:, :N]
            fn = lambda: ops.gemm(A, am, B, bm, C, M, N, K, accumulate=bool(acc), precision=mode)
            C.zero_()
            fn()
            torch.cuda.synchronize()
            e = errs(C, R) if R is not None else (float("nan"),) * 3
            us = timeit(fn)
            print(f"| {name} {M}x{N}x{K} | {mode} | {us:.1f} | {2.0 * M * N * K / us / 1e6:.0f} | {e[0]:.2e} | {e[1]:.2e} | {e[2]:+.2e} |", flush=True)
        if R is not None:   # library fp32 (cuBLAS SGEMM, allow_tf32 off) for context only
            torch.backends.cuda.matmul.allow_tf32 = False
            a = (A.t() if am else A)
            b = (B.t() if bm else B)
            Cc = a @ b.t()
            e = errs(Cc, R)
            us = timeit(lambda: a @ b.t())
            print(f"| {name} | torch fp32 (cuBLAS, context) | {us:.1f} | {2.0 * M * N * K / us / 1e6:.0f} | {e[0]:.2e} | {e[1]:.2e} | {e[2]:+.2e} |", flush=True)


def ksweep():
    """One accumulation chain of length K: error growth of the TMEM accumulator (positive operands = worst case for a
    truncating adder, zero-mean operands = typical)."""
    torch.manual_seed(1)
    print("| K | operands | mode | split_k | max-abs/max | rel-L2 | norm bias |")
    print("|---|---|---|---|---|---|---|")
    M, N = 512, 768
    for K in (256, 768, 3072, 12800, 50304):
        for kind in ("randn", "uniform+"):
            gen = (lambda r, c: mat(r, c)) if kind == "randn" else (lambda r, c: torch.rand(r, pad(c), device="cuda")[:, :c] + 0.5)
            A, B = gen(M, K), gen(N, K)
            R = ref64(A, 0, B, 0)
            for mode, sk in (("tf32", 1), ("tf32x3", 1), ("tf32x3", max(1, K // 2048)), ("tf32x3", max(1, K // 512)), ("fp32", 1)):
                if mode != "tf32x3" and sk != 1:
                    continue
                if sk > 1 and K // 32 // sk < 4:
                    continue
                C = torch.zeros(M, N, device="cuda")
                ops.gemm(A, 0, B, 0, C, M, N, K, accumulate=True, split_k=sk, precision=mode)
                torch.cuda.synchronize()
                e = errs(C, R)
                print(f"| {K} | {kind} | {mode} | {sk} | {e[0]:.2e} | {e[1]:.2e} | {e[2]:+.2e} |", flush=True)


if __name__ == "__main__":
    os.environ.setdefault("CAPDEC_X3_CHAIN", "0")   # the probe controls split-K itself
    if len(sys.argv) > 1 and sys.argv[1] == "ksweep":
        ksweep()
    else:
        shapes(sys.argv[1:])
