"""GPU bring-up driver for the tcgen05 GEMM (not a pytest file; run under gpurun).

    python tests/bringup_gemm.py master            # runs every group in watchdog'ed subprocesses
    python tests/bringup_gemm.py child <group> [start]

Writes gpurun_out/bringup_gemm.log.  Each group runs in its own process so that a trapped kernel (mbarrier
watchdog, illegal descriptor) cannot take the remaining groups down with it.
"""
from __future__ import annotations

import ctypes as C
import itertools
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"


def _bind():
    lib = C.CDLL(str(ROOT / "capdec_b200" / "libcapdec_b200.so"))
    p, i, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.capdec_gemm_tf32.argtypes = [p, i, i64, p, i, i64, p, i64, i, i, i, p, i, p, i, i, p, p, i, i, p]
    lib.capdec_gemm_fp32_simt.argtypes = [p, i, i64, p, i, i64, p, i64, i, i, i, p, i, p, i, p]
    lib.capdec_split_tf32.argtypes = [p, p, p, i64, p]
    lib.capdec_gemm_debug_mn_encoding.argtypes = [i, i, i, i]
    lib.capdec_gemm_debug_mn_encoding.restype = None
    lib.capdec_last_error.restype = C.c_char_p
    return lib


def trunc_tf32(x):
    import torch
    return (x.view(torch.int32) & -8192).view(torch.float32)


def rn_tf32(x):
    import torch
    i = x.view(torch.int32)
    return ((i + 0x1000) & -8192).view(torch.float32)


def run_gemm(lib, A, a_major, B, b_major, M, N, K, bias=None, act=0, aux=None, accumulate=0, C_out=None, precision=0,
             a_lo=None, b_lo=None, bn=0, split=0, simt=False):
    import torch
    ldc = (N + 3) // 4 * 4
    Cm = C_out if C_out is not None else torch.zeros(M, ldc, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ptr = lambda t: t.data_ptr() if t is not None else None
    if simt:
        rc = lib.capdec_gemm_fp32_simt(A.data_ptr(), a_major, A.stride(0), B.data_ptr(), b_major, B.stride(0),
                                       Cm.data_ptr(), Cm.stride(0), M, N, K, ptr(bias), act, ptr(aux), accumulate, st)
    else:
        rc = lib.capdec_gemm_tf32(A.data_ptr(), a_major, A.stride(0), B.data_ptr(), b_major, B.stride(0), Cm.data_ptr(),
                                  Cm.stride(0), M, N, K, ptr(bias), act, ptr(aux), accumulate, precision, ptr(a_lo),
                                  ptr(b_lo), bn, split, st)
    if rc != 0:
        raise RuntimeError(f"rc={rc}: {lib.capdec_last_error().decode()}")
    torch.cuda.synchronize()
    return Cm[:, :N]


def make_operands(M, N, K, a_major, b_major, seed=0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = lambda n: (n + 3) // 4 * 4
    if a_major == 0:
        A = torch.randn(M, pad(K), device="cuda", generator=g)[:, :K]
        Al = A  # logical [M,K]
    else:
        A = torch.randn(K, pad(M), device="cuda", generator=g)[:, :M]
        Al = A.t()
    if b_major == 0:
        B = torch.randn(N, pad(K), device="cuda", generator=g)[:, :K]
        Bl = B
    else:
        B = torch.randn(K, pad(N), device="cuda", generator=g)[:, :N]
        Bl = B.t()
    return A, Al, B, Bl


def check_case(lib, M, N, K, a_major, b_major, bn=0, tag=""):
    import torch
    A, Al, B, Bl = make_operands(M, N, K, a_major, b_major)
    out = run_gemm(lib, A, a_major, B, b_major, M, N, K, bn=bn)
    ref_t = (trunc_tf32(Al.contiguous()).double() @ trunc_tf32(Bl.contiguous()).double().t())
    ref_r = (rn_tf32(Al.contiguous()).double() @ rn_tf32(Bl.contiguous()).double().t())
    ref_x = Al.double() @ Bl.double().t()
    scale = ref_x.abs().max().item()
    e_t = (out.double() - ref_t).abs().max().item() / scale
    e_r = (out.double() - ref_r).abs().max().item() / scale
    e_x = (out.double() - ref_x).abs().max().item() / scale
    ok = e_t < 2e-5 or e_r < 2e-5
    rec = dict(tag=tag, M=M, N=N, K=K, a_major=a_major, b_major=b_major, bn=bn, err_trunc=e_t, err_rn=e_r,
               err_exact=e_x, ok=ok)
    print("CASE " + json.dumps(rec), flush=True)
    return ok


def group_basic(lib, a_major, b_major):
    shapes = [(128, 256, 32), (128, 256, 64), (128, 64, 32), (128, 128, 256), (256, 512, 768), (200, 300, 100),
              (1600, 2304, 768), (384, 50257, 64)]
    allok = True
    for (M, N, K) in shapes:
        for bn in ([0] if (M, N, K) != (256, 512, 768) else [64, 128, 256]):
            allok &= check_case(lib, M, N, K, a_major, b_major, bn=bn, tag=f"basic{a_major}{b_major}")
    return allok


MN_CANDIDATES = [(lay, lbo, sbo, swz) for lay, swz in ((1, 4), (2, 3), (1, 3), (2, 4))
                 for (lbo, sbo) in ((4096, 512), (512, 4096), (4096, 1024), (1024, 4096), (4096, 256), (256, 4096),
                                    (4096, 128), (128, 4096))]


def group_sweep(lib, a_major, b_major, start):
    for idx in range(start, len(MN_CANDIDATES)):
        lay, lbo, sbo, swz = MN_CANDIDATES[idx]
        lib.capdec_gemm_debug_mn_encoding(lay, lbo, sbo, swz)
        print(f"SWEEP_BEGIN {idx}", flush=True)
        try:
            ok1 = check_case(lib, 128, 256, 32, a_major, b_major, tag=f"sweep[{idx}]{lay},{lbo},{sbo},{swz}")
            ok2 = check_case(lib, 256, 512, 96, a_major, b_major, tag=f"sweep[{idx}]{lay},{lbo},{sbo},{swz}") if ok1 else False
        except Exception as e:  # noqa
            print(f"SWEEP_ERR {idx} {e}", flush=True)
            import torch
            torch.cuda.synchronize()  # raises (and ends this child) if the context is dead
            ok1 = ok2 = False
        print(f"SWEEP_END {idx} ok={ok1 and ok2}", flush=True)
    lib.capdec_gemm_debug_mn_encoding(-1, -1, -1, -1)


def group_epilogue(lib):
    import torch
    M, N, K = 384, 640, 256
    A, Al, B, Bl = make_operands(M, N, K, 0, 0)
    bias = torch.randn(N, device="cuda")
    base = Al.double() @ Bl.double().t() + bias.double()
    for act, f in ((0, lambda x: x), (1, lambda x: torch.nn.functional.gelu(x, approximate="tanh")),
                   (2, torch.tanh), (3, torch.relu)):
        aux = torch.zeros(M, N, device="cuda")
        out = run_gemm(lib, A, 0, B, 0, M, N, K, bias=bias, act=act, aux=aux)
        e = (out.double() - f(base)).abs().max().item()
        ea = (aux.double() - base).abs().max().item()
        print("CASE " + json.dumps(dict(tag="epilogue", act=act, err=e, err_aux=ea, ok=e < 0.1 and ea < 0.1)), flush=True)
    # accumulate + split-K (wgrad-like: K large)
    M, N, K = 768, 512, 4096
    A, Al, B, Bl = make_operands(M, N, K, 1, 1)
    C0 = torch.randn(M, N, device="cuda")
    for split in (1, 4, 0):
        Cc = C0.clone()
        out = run_gemm(lib, A, 1, B, 1, M, N, K, accumulate=1, C_out=Cc, split=split)
        ref = C0.double() + Al.double() @ Bl.double().t()
        e = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        print("CASE " + json.dumps(dict(tag="splitk", split=split, err=e, ok=e < 3e-3)), flush=True)
    # 3xTF32
    M, N, K = 512, 768, 768
    for (am, bm) in ((0, 0), (0, 1)):
        A, Al, B, Bl = make_operands(M, N, K, am, bm)
        Ac, Bc = A.contiguous(), B.contiguous()
        ah, al, bh, bl = (torch.empty_like(Ac), torch.empty_like(Ac), torch.empty_like(Bc), torch.empty_like(Bc))
        st = torch.cuda.current_stream().cuda_stream
        lib.capdec_split_tf32(Ac.data_ptr(), ah.data_ptr(), al.data_ptr(), Ac.numel(), st)
        lib.capdec_split_tf32(Bc.data_ptr(), bh.data_ptr(), bl.data_ptr(), Bc.numel(), st)
        ref = Al.double() @ Bl.double().t()
        o1 = run_gemm(lib, Ac, am, Bc, bm, M, N, K)
        o3 = run_gemm(lib, ah, am, bh, bm, M, N, K, precision=1, a_lo=al, b_lo=bl)
        o3b = run_gemm(lib, Ac, am, Bc, bm, M, N, K, precision=1, a_lo=al, b_lo=bl)  # hi = raw operand (hw truncation)
        of = run_gemm(lib, Ac, am, Bc, bm, M, N, K, simt=True)
        t32 = (Al @ Bl.t())
        sc = ref.abs().max().item()
        rec = dict(tag="3xtf32", a_major=am, b_major=bm, err_1x=(o1.double() - ref).abs().max().item() / sc,
                   err_3x=(o3.double() - ref).abs().max().item() / sc,
                   err_3x_rawhi=(o3b.double() - ref).abs().max().item() / sc,
                   err_simt=(of.double() - ref).abs().max().item() / sc,
                   err_torch_fp32=(t32.double() - ref).abs().max().item() / sc)
        rec["ok"] = rec["err_3x"] < 5e-6
        print("CASE " + json.dumps(rec), flush=True)


def group_perf(lib):
    import torch
    shapes = [  # (name, M, N, K, a_major, b_major, accumulate)
        ("qkv_fwd", 12800, 2304, 768, 0, 1, 0), ("qkv_fwd_kk", 12800, 2304, 768, 0, 0, 0),
        ("attn_proj", 12800, 768, 768, 0, 1, 0), ("fc", 12800, 3072, 768, 0, 1, 0), ("fc_proj", 12800, 768, 3072, 0, 1, 0),
        ("qkv_dgrad", 12800, 768, 2304, 0, 0, 0), ("qkv_wgrad", 768, 2304, 12800, 1, 1, 1),
        ("fc_wgrad", 768, 3072, 12800, 1, 1, 1), ("lm_head", 10240, 50257, 768, 0, 0, 0),
        ("lm_dgrad", 10240, 768, 50257, 0, 1, 0), ("lm_wgrad", 50257, 768, 10240, 1, 1, 1),
        ("mlp_fc2", 256, 7680, 3840, 0, 0, 0),
    ]
    for (name, M, N, K, am, bm, acc) in shapes:
        A, Al, B, Bl = make_operands(M, N, K, am, bm)
        ldc = (N + 3) // 4 * 4
        Cm = torch.zeros(M, ldc, device="cuda")
        for bn in (256, 128, 64, 0):
            try:
                for _ in range(3):
                    run_gemm(lib, A, am, B, bm, M, N, K, accumulate=acc, C_out=Cm, bn=bn)
                st = torch.cuda.current_stream().cuda_stream
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 10
                e0.record()
                for _ in range(iters):
                    lib.capdec_gemm_tf32(A.data_ptr(), am, A.stride(0), B.data_ptr(), bm, B.stride(0), Cm.data_ptr(),
                                         Cm.stride(0), M, N, K, None, 0, None, acc, 0, None, None, bn, 0, st)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                tf = 2.0 * M * N * K / ms / 1e9
                print("PERF " + json.dumps(dict(name=name, M=M, N=N, K=K, a_major=am, b_major=bm, bn=bn, ms=ms, tflops=tf)),
                      flush=True)
            except Exception as e:  # noqa
                print(f"PERF_ERR {name} bn={bn}: {e}", flush=True)
        # cuBLAS TF32 for context (library baseline, not our path)
        torch.backends.cuda.matmul.allow_tf32 = True
        Ac, Bc = Al.contiguous(), Bl.contiguous()
        for _ in range(3):
            Ac @ Bc.t()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            Ac @ Bc.t()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("PERF " + json.dumps(dict(name=name + "_cublas_tf32", ms=ms, tflops=2.0 * M * N * K / ms / 1e9)), flush=True)
        torch.backends.cuda.matmul.allow_tf32 = False


def child(group, start):
    import torch
    assert torch.cuda.is_available()
    lib = _bind()
    if group.startswith("basic"):
        ok = group_basic(lib, int(group[5]), int(group[6]))
        print(f"GROUP {group} ok={ok}", flush=True)
    elif group.startswith("sweep"):
        group_sweep(lib, int(group[5]), int(group[6]), start)
    elif group == "epilogue":
        group_epilogue(lib)
    elif group == "perf":
        group_perf(lib)


def master(groups=None):
    OUT.mkdir(exist_ok=True)
    log = open(OUT / "bringup_gemm.log", "a")
    groups = groups or ["basic00", "basic01", "basic10", "basic11", "epilogue", "perf"]
    results = {}

    def run(group, start=0, timeout=240):
        cmd = [sys.executable, __file__, "child", group, str(start)]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
            out, rc = r.stdout + "\n" + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout.decode() if e.stdout else "") + "\nTIMEOUT"
            rc = -9
        log.write(f"===== {group} start={start} rc={rc} ({time.time() - t0:.1f}s)\n{out}\n")
        log.flush()
        print(f"===== {group} start={start} rc={rc}\n{out[-6000:]}", flush=True)
        return rc, out

    for g in groups:
        rc, out = run(g)
        ok = rc == 0 and "ok=False" not in out and '"ok": false' not in out
        results[g] = ok
        if g.startswith("basic") and not ok and g != "basic00":
            # sweep the MN-major descriptor encodings
            sg = "sweep" + g[5:]
            start = 0
            while start < len(MN_CANDIDATES):
                rc, out = run(sg, start)
                last = [int(l.split()[1]) for l in out.splitlines() if l.startswith("SWEEP_BEGIN")]
                if rc == 0 or not last:
                    break
                start = last[-1] + 1
    print("SUMMARY " + json.dumps(results), flush=True)
    log.write("SUMMARY " + json.dumps(results) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "master":
        master(sys.argv[2:] or None)
    else:
        child(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
