"""Time individual kernels of the C2 step with CUDA events (one B200, no profiler):

    python tools/op_probe.py gemm      # every GEMM shape of the packed step x engine x tile width
    python tools/op_probe.py ln        # add_ln_bwd (set CAPDEC_LN_BWD_PIPE=0 for the register-staged kernel)
    python tools/op_probe.py attn      # attention forward / backward, dense and packed

Prints one line per configuration: microseconds per launch and TFLOP/s or GB/s of ALGORITHMIC work.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import capdec_b200 as cb  # noqa: E402
from capdec_b200 import ops, _lib  # noqa: E402

M_MAX, M_LIVE = 12800, 8820          # C2: 256 captions x 50 positions; live rows with lengths ~ U{8..40}
T_MAX, T_LIVE = 10240, 6144          # LM-head rows (targets) / non-ignored ones


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def pad(n, q=128):
    return (n + q - 1) // q * q


def mat(rows, cols):
    return torch.randn(rows, pad(cols), device="cuda")[:, :cols]


def gemm_sweep():
    lib = _lib.load()
    # name: (M, N, K, a_major, b_major, kind)   kind: plain | mul | acc_k (split-K wgrad with a k limit)
    shapes = {
        "qkv_fwd": (M_MAX, 2304, 768, 0, 1, "plain"), "aproj_fwd": (M_MAX, 768, 768, 0, 1, "plain"),
        "fc_fwd": (M_MAX, 3072, 768, 0, 1, "plain"), "fcproj_fwd": (M_MAX, 768, 3072, 0, 1, "plain"),
        "qkv_dgrad": (M_MAX, 768, 2304, 0, 0, "plain"), "aproj_dgrad": (M_MAX, 768, 768, 0, 0, "plain"),
        "fc_dgrad": (M_MAX, 768, 3072, 0, 0, "plain"), "fcproj_dgrad_mul": (M_MAX, 3072, 768, 0, 0, "mul"),
        "qkv_wgrad": (768, 2304, M_MAX, 1, 1, "acc_k"), "aproj_wgrad": (768, 768, M_MAX, 1, 1, "acc_k"),
        "fc_wgrad": (768, 3072, M_MAX, 1, 1, "acc_k"), "fcproj_wgrad": (3072, 768, M_MAX, 1, 1, "acc_k"),
        "lm_head": (T_MAX, 50257, 768, 0, 0, "plain_t"), "lm_dgrad": (T_MAX, 768, 50257, 0, 1, "plain_t"),
        "lm_wgrad": (50257, 768, T_MAX, 1, 1, "acc_kt"),
    }
    only = sys.argv[2:] or list(shapes)
    lim_rows = torch.tensor([M_LIVE], device="cuda", dtype=torch.int32)
    lim_t = torch.tensor([T_LIVE], device="cuda", dtype=torch.int32)
    for name in only:
        M, N, K, am, bm, kind = shapes[name]
        A = mat(K, M) if am else mat(M, K)
        B = mat(K, N) if bm else mat(N, K)
        C = torch.zeros(M, pad(N), device="cuda")[:, :N]
        u = mat(M, N) if kind == "mul" else None
        col = torch.zeros(N, device="cuda") if kind == "mul" else None
        live = {"plain": M_LIVE, "mul": M_LIVE, "plain_t": T_LIVE}.get(kind)
        klive = {"acc_k": M_LIVE, "acc_kt": T_LIVE}.get(kind)
        flops = 2.0 * N * ((live or M) * K if klive is None else M * klive)
        res = []
        for mode in (-1, 1, 3):
            for bn in (0, 128, 192, 256):
                if mode == -1 and bn != 0:
                    continue
                if mode != -1 and bn == 0:
                    continue
                lib.capdec_gemm_debug_force_pair(mode)
                ops.set_row_hint(live or 0)
                try:
                    if kind == "mul":
                        fn = lambda: ops.gemm_mul(A, am, B, bm, C, M, N, K, u, 1, col, block_n=bn, m_limit=lim_rows)
                    elif kind in ("acc_k", "acc_kt"):
                        kl = lim_rows if kind == "acc_k" else lim_t
                        fn = lambda: ops.gemm(A, am, B, bm, C, M, N, K, accumulate=True, block_n=bn, k_limit=kl)
                    else:
                        ml = lim_rows if kind == "plain" else lim_t
                        fn = lambda: ops.gemm(A, am, B, bm, C, M, N, K, block_n=bn, m_limit=ml)
                    us = timeit(fn, iters=10 if "lm_" in name else 20)
                    res.append((mode, bn, us))
                except _lib.CapdecError as e:
                    res.append((mode, bn, None))
        lib.capdec_gemm_debug_force_pair(-1)
        ops.set_row_hint(0)
        cells = ", ".join(f"m{m}/bn{b}: " + (f"{u_:.1f}us {flops / u_ / 1e6:.0f}TF" if u_ else "n/a") for m, b, u_ in res)
        print(f"{name:18s} {cells}", flush=True)


def ln_probe():
    d = 768
    for rows_live in (M_MAX, M_LIVE):
        dx, r, dh, dy = (torch.randn(M_MAX, d, device="cuda") for _ in range(4))
        st = torch.rand(M_MAX, 2, device="cuda") + 0.5
        gam = torch.randn(d, device="cuda")
        dg, dbt, dbr = (torch.zeros(d, device="cuda") for _ in range(3))
        seed = ops.make_seed(1, dx.device)
        rows = torch.tensor([rows_live, pad(rows_live, 32)], device="cuda", dtype=torch.int32)
        fn = lambda: ops.add_ln_bwd(dx, r, st, gam, dh, dh, dy, dg, dbt, p_drop=0.1, seed=seed, stream_id=3,
                                    dbias_branch=dbr, rows=rows[0:1])
        us = timeit(fn)
        mb = 5 * rows_live * d * 4 / 1e6
        print(f"add_ln_bwd rows={rows_live}: {us:.1f} us, {mb / us * 1e3 / 1e3:.2f} TB/s ({mb:.0f} MB algorithmic)", flush=True)
        y = torch.randn(M_MAX, d, device="cuda")
        h2, x2 = torch.empty_like(y), torch.empty_like(y)
        fn = lambda: ops.add_ln_fwd(dx, y, h2, x2, st, gam, gam, p_drop=0.1, seed=seed, stream_id=3, rows=rows[0:1])
        us = timeit(fn)
        mb = 4 * rows_live * d * 4 / 1e6
        print(f"add_ln_fwd rows={rows_live}: {us:.1f} us, {mb / us * 1e3 / 1e3:.2f} TB/s", flush=True)


def attn_probe():
    B, H, T, hd, d = 256, 12, 50, 64, 768
    M = B * T
    qkv = torch.randn(M, 3 * d, device="cuda")
    dqkv = torch.zeros(M, 3 * d, device="cuda")
    ctx, dctx = torch.zeros(M, d, device="cuda"), torch.randn(M, d, device="cuda")
    lse = torch.zeros(B * H * T, device="cuda")
    dbias = torch.zeros(3 * d, device="cuda")
    seed = ops.make_seed(1, qkv.device)
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    dq, dk, dv = dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:]
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(8, 41, (B,), generator=g) + 10
    for name, cu in (("dense", None), ("packed", torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).to(torch.int32).cuda())):
        rows = M if cu is None else int(cu[-1])
        f = lambda: ops.attention_fwd(q, k, v, ctx, lse, B, H, T, T, hd, T * 3 * d, 3 * d, T * 3 * d, 3 * d, T * d, d, 0.125, 1,
                                      p_drop=0.1, seed=seed, stream_id=16, cu_rows=cu)
        us = timeit(f)
        print(f"attention_fwd {name} rows={rows}: {us:.1f} us, {4 * rows * d * 4 / us / 1e6:.2f} TB/s", flush=True)
        b = lambda: ops.attention_bwd(q, k, v, ctx, dctx, lse, dq, dk, dv, B, H, T, T, hd, T * 3 * d, 3 * d, T * 3 * d, 3 * d,
                                      T * d, d, 0.125, 1, p_drop=0.1, seed=seed, stream_id=16, dbias_qkv=dbias, cu_rows=cu)
        us = timeit(b)
        print(f"attention_bwd {name} rows={rows}: {us:.1f} us, {8 * rows * d * 4 / us / 1e6:.2f} TB/s", flush=True)


def mul_probe():
    """The fused dgrad x act' (+ bias-gradient column sums) epilogue, piece by piece (set CAPDEC_GEMM_DBG=4 to drop stores)."""
    M, N, K = M_MAX, 3072, 768
    A, B, C, u = mat(M, K), mat(N, K), torch.zeros(M, pad(N), device="cuda")[:, :N], mat(M, N)
    col = torch.zeros(N, device="cuda")
    lim = torch.tensor([M_LIVE], device="cuda", dtype=torch.int32)
    ops.set_row_hint(M_LIVE)
    flops = 2.0 * N * M_LIVE * K
    for name, fn in {
        "plain dgrad (no epilogue input)": lambda: ops.gemm(A, 0, B, 0, C, M, N, K, m_limit=lim),
        "mul gelu' + colsum": lambda: ops.gemm_mul(A, 0, B, 0, C, M, N, K, u, 1, col, m_limit=lim),
        "mul stored-derivative + colsum": lambda: ops.gemm_mul(A, 0, B, 0, C, M, N, K, u, 4, col, m_limit=lim),
        "mul stored-derivative, no colsum": lambda: ops.gemm_mul(A, 0, B, 0, C, M, N, K, u, 4, None, m_limit=lim),
        "mul relu mask, no colsum": lambda: ops.gemm_mul(A, 0, B, 0, C, M, N, K, u, 3, None, m_limit=lim),
    }.items():
        us = timeit(fn)
        print(f"{name:36s} {us:7.1f} us  {flops / us / 1e6:5.0f} TF/s", flush=True)
    ops.set_row_hint(0)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    {"gemm": gemm_sweep, "ln": ln_probe, "attn": attn_probe, "mul": mul_probe}[what]()
