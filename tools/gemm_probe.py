"""Launch one GEMM shape of the C2 step repeatedly (for `ncu --set full -k regex:gemm_tf32 -s 5 -c 3 ...`)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import capdec_b200 as cb  # noqa: E402

SHAPES = {  # name: (M, N, K, a_major, b_major, accumulate)
    "qkv": (12800, 2304, 768, 0, 1, 0), "fc": (12800, 3072, 768, 0, 1, 0), "fc_proj": (12800, 768, 3072, 0, 1, 0),
    "qkv_wgrad": (768, 2304, 12800, 1, 1, 1), "lm_head": (10240, 50257, 768, 0, 0, 0),
    "attn_proj": (12800, 768, 768, 0, 1, 0), "qkv_dgrad": (12800, 768, 2304, 0, 0, 0), "fc_dgrad": (12800, 768, 3072, 0, 0, 0),
    "fc_wgrad": (768, 3072, 12800, 1, 1, 1), "lm_dgrad": (10240, 768, 50257, 0, 1, 0), "lm_wgrad": (50257, 768, 10240, 1, 1, 1),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "qkv"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    act = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    use_aux = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    M, N, K, am, bm, acc = SHAPES[name]
    pad = lambda n: (n + 127) // 128 * 128
    A = torch.randn((K, pad(M)) if am else (M, pad(K)), device="cuda")[:, : (M if am else K)]
    B = torch.randn((K, pad(N)) if bm else (N, pad(K)), device="cuda")[:, : (N if bm else K)]
    C = torch.zeros(M, pad(N), device="cuda")[:, :N]
    bias = None if acc else torch.zeros(N, device="cuda")
    aux = torch.zeros(M, pad(N), device="cuda")[:, :N] if use_aux else None
    kw = dict(bias=bias, accumulate=bool(acc), act=act, aux=aux)
    for _ in range(iters):
        cb.ops.gemm(A, am, B, bm, C, M, N, K, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        cb.ops.gemm(A, am, B, bm, C, M, N, K, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    import os
    print(f"{name} act={act} aux={use_aux} dbg={os.environ.get('CAPDEC_GEMM_DBG','0')} mode={os.environ.get('CAPDEC_GEMM_MODE', 'auto')}: {ms * 1e3:.1f} us, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
