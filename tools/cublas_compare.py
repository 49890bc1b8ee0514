"""Our tcgen05 TF32 GEMM vs the library's TF32 GEMM (torch.matmul with allow_tf32 = cuBLAS) on the trunk shapes of the C2
step, SAME process, same clocks, CUDA events, L2-cold rotation over distinct operand sets.  The library is context: it is
never on the product path.  `python tools/cublas_compare.py [shape ...]`; with `lib` as first argument only the library
kernels run (for `ncu -k regex:...` captures of what cuBLAS launches: tile shape, cluster size, stages)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import capdec_b200 as cb  # noqa: E402
from capdec_b200 import ops  # noqa: E402

SHAPES = {  # name: (M, N, K, a_major, b_major, accumulate)
    "qkv": (12800, 2304, 768, 0, 1, 0), "attn_proj": (12800, 768, 768, 0, 1, 0), "fc": (12800, 3072, 768, 0, 1, 0),
    "fc_proj": (12800, 768, 3072, 0, 1, 0), "qkv_dgrad": (12800, 768, 2304, 0, 0, 0), "fc_dgrad": (12800, 768, 3072, 0, 0, 0),
    "qkv_wgrad": (768, 2304, 12800, 1, 1, 1), "fc_wgrad": (768, 3072, 12800, 1, 1, 1), "fcproj_wgrad": (3072, 768, 12800, 1, 1, 1),
    "lm_head": (6144, 50257, 768, 0, 0, 0), "lm_wgrad": (50257, 768, 6144, 1, 1, 1),
}
NSETS = 6


def pad(n, q=128):
    return (n + q - 1) // q * q


def mat(rows, cols):
    return torch.randn(rows, pad(cols), device="cuda")[:, :cols]


def timeit(fns, iters=24, warm=6):
    for i in range(warm):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    args = sys.argv[1:]
    lib_only = bool(args) and args[0] == "lib"
    if lib_only:
        args = args[1:]
    torch.backends.cuda.matmul.allow_tf32 = True
    ops.set_precision("tf32")
    print("| shape | M x N x K | ours us | ours TF/s | cuBLAS-TF32 us | cuBLAS TF/s | ours / cuBLAS |")
    print("|---|---|---|---|---|---|---|")
    for name in args or list(SHAPES):
        M, N, K, am, bm, acc = SHAPES[name]
        sets = []
        for _ in range(NSETS if M * N < 1e8 else 2):
            A = mat(K, M) if am else mat(M, K)
            B = mat(K, N) if bm else mat(N, K)
            C = torch.zeros(M, pad(N), device="cuda")[:, :N]
            sets.append((A, B, C))
        ours = [(lambda A=A, B=B, C=C: ops.gemm(A, am, B, bm, C, M, N, K, accumulate=bool(acc))) for A, B, C in sets]
        libf = []
        for A, B, C in sets:
            a = A.t() if am else A
            b = B if bm else B.t()
            libf.append((lambda a=a, b=b, C=C: torch.matmul(a, b, out=C)) if not acc else (lambda a=a, b=b, C=C: C.addmm_(a, b)))
        t_lib = timeit(libf)
        if lib_only:
            print(f"| {name} | {M}x{N}x{K} | - | - | {t_lib:.1f} | {2.0 * M * N * K / t_lib / 1e6:.0f} | - |", flush=True)
            continue
        ops.gemm_autotune(1)          # measured plan for this problem (as the Trainer does), then the timed run
        ours[0]()
        ops.gemm_autotune(0)
        t_ours = timeit(ours)
        t_lib2 = timeit(libf)
        t_l = min(t_lib, t_lib2)
        print(f"| {name} | {M}x{N}x{K} | {t_ours:.1f} | {2.0 * M * N * K / t_ours / 1e6:.0f} | {t_l:.1f} | {2.0 * M * N * K / t_l / 1e6:.0f} | "
              f"{t_l / t_ours:.2f} |", flush=True)


if __name__ == "__main__":
    main()
