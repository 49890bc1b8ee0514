"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU, and exports exactly the
entry points include/capdec_b200.h declares; the ctypes table binds every one of them with the declared arity; and the
product path refuses to run without CUDA (no CPU fallback).  No compute calls are made here."""
import re
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "capdec_b200.h"


def declared():
    """name -> number of parameters, parsed from the header's prototypes."""
    src = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(capdec_\w+)\s*\(([^;{]*?)\)\s*;", src):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_header_library_and_ctypes_table_agree():
    from capdec_b200 import _lib, build
    so = build.build()
    decl = declared()
    assert len(decl) >= 40, "header parse found too few prototypes"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(so)], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if " T " in ln and "capdec_" in ln}
    missing = sorted(set(decl) - exported)
    assert not missing, f"declared in the header but not exported by the library: {missing}"
    undeclared = sorted(n for n in exported if n.startswith("capdec_") and n not in decl)
    assert not undeclared, f"exported extern \"C\" symbols missing from the header: {undeclared}"
    lib = _lib.load()                       # dlopen works without a GPU (no CUDA call happens at load)
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in decl, f"ctypes table binds {name}, which the header does not declare"
        assert len(argtypes) == decl[name], f"{name}: ctypes table has {len(argtypes)} args, header {decl[name]}"
        assert getattr(lib, name) is not None
    unbound = sorted(set(decl) - set(_lib.SIGNATURES))
    assert not unbound, f"declared but not bound by capdec_b200/_lib.py: {unbound}"
    assert lib.capdec_version() >= 100


def test_plain_c_host_links_against_the_abi(tmp_path):
    """The boundary is a C ABI, not a Python extension: a C99 translation unit that includes only include/capdec_b200.h
    compiles with gcc, links against libcapdec_b200.so, takes the address of EVERY declared entry point and calls the
    two that need no GPU (the stub a non-Python host - cgo / JNI / a C++ trainer - would start from)."""
    from capdec_b200 import build
    so = build.build()
    names = sorted(declared())
    refs = "\n".join(f"  table[{i}] = (void*)&{n};" for i, n in enumerate(names))
    src = tmp_path / "host.c"
    src.write_text(f"""#include <stdio.h>
#include "capdec_b200.h"
int main(void) {{
  void* table[{len(names)}];
{refs}
  for (int i = 0; i < {len(names)}; ++i) if (!table[i]) return 2;
  const char* err = capdec_last_error();
  printf("%d %d %s.\\n", capdec_version(), (int)capdec_launch_count(), err ? err : "(null)");
  /* bad arguments are rejected before any CUDA call: error code + message, no crash */
  int rc = capdec_gemm_tf32(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  printf("%d %s\\n", rc, capdec_last_error());
  return rc < 0 ? 0 : 3;
}}
""")
    exe = tmp_path / "host"
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(HEADER.parent), str(src), "-o", str(exe),
                         "-L", str(so.parent), "-l:" + so.name, "-Wl,-rpath," + str(so.parent)],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    first, second = run.stdout.strip().splitlines()
    assert first.split()[0] == "100"
    assert second.startswith("-1 ") and "null operand" in second


def test_every_entry_point_rejects_null_arguments_with_a_message():
    """Error behaviour of the boundary: bad arguments never reach a launch or crash the host - every
    compute entry point answers all-null / all-zero arguments with CAPDEC_ERR_INVALID (-1) and a message in
    capdec_last_error().  Argument checks run before any CUDA call, so this needs no GPU."""
    import ctypes as C
    from capdec_b200 import _lib
    lib = _lib.load()
    not_compute = {"capdec_last_error", "capdec_version", "capdec_launch_count", "capdec_gemm_debug_mn_encoding",
                   "capdec_gemm_debug_force_pair", "capdec_gemm_debug_trace", "capdec_gemm_set_row_hint", "capdec_gemm_autotune", "capdec_gemm_plan_query",
                   "capdec_gemm_set_schedule"}
    before = lib.capdec_launch_count()
    checked = 0
    for name, argtypes in _lib.SIGNATURES.items():
        if name in not_compute:
            continue
        args = [None if t is C.c_void_p else (0.0 if t is C.c_float else 0) for t in argtypes]
        rc = getattr(lib, name)(*args)
        msg = lib.capdec_last_error()
        assert rc == -1, (name, rc)
        assert msg and len(msg) > 4, name
        checked += 1
    assert checked == len(_lib.SIGNATURES) - len(not_compute) >= 35
    assert lib.capdec_launch_count() == before          # nothing was launched


def test_no_cpu_fallback():
    """CPU tensors raise instead of silently computing somewhere else."""
    from capdec_b200 import ops
    from capdec_b200._lib import CapdecError
    a = torch.zeros(4, 4)
    with pytest.raises(CapdecError):
        ops.gemm(a, 0, a, 0, a, 4, 4, 4)


def test_gemm_planner_returns_legal_plans_on_cpu():
    """Host-side planner (engine / tile width / split-K) over the shapes of the C1-C4 steps and some ragged ones: every
    plan must be one the kernel accepts (pairs need >= 128 columns, 192-wide tiles never with the B-sharing quad on an
    MN-major B, split-K only for accumulating GEMMs and never finer than 4 k-blocks)."""
    from capdec_b200 import _lib
    lib = _lib.load()
    shapes = [(12800, 2304, 768), (12800, 768, 768), (12800, 3072, 768), (12800, 768, 3072), (768, 2304, 12800),
              (3072, 768, 12800), (10240, 50257, 768), (50257, 768, 10240), (10240, 768, 50257), (256, 7680, 3840),
              (32, 3840, 512), (1, 50257, 768), (5120, 2304, 768), (200, 300, 100), (20480, 1536, 768), (130, 120, 40)]
    seen_modes = set()
    for (M, N, K) in shapes:
        for b_major in (0, 1):
            for acc in (0, 1):
                for hinted in (0, 1):
                    if hinted:
                        lib.capdec_gemm_set_row_hint(max(1, int(M * 0.69)))
                    code = lib.capdec_gemm_plan_query(M, N, K, b_major, acc, 0, 0, hinted)
                    lib.capdec_gemm_set_row_hint(0)
                    assert code > 0, (M, N, K, code)
                    mode, bn, splits = code & 0xFF, (code >> 8) & 0xFFF, code >> 20
                    seen_modes.add(mode)
                    assert mode in (0, 1, 2, 3) and bn in (64, 128, 192, 256) and splits >= 1, (M, N, K, mode, bn, splits)
                    assert mode == 0 or (bn >= 128 and M > 128 and N >= 128), (M, N, K, mode, bn)
                    assert not (bn == 192 and mode == 3 and b_major), (M, N, K)
                    assert splits == 1 or (acc and (K + 31) // 32 // splits >= 4), (M, N, K, splits)
    assert {0, 1, 3} <= seen_modes
    # engine heuristic (profiles/r2_gemm_timeline.md): CTA pairs at dense extents, the B-sharing quad for row-limited launches
    assert lib.capdec_gemm_plan_query(12800, 2304, 768, 1, 0, 0, 0, 0) & 0xFF == 1
    lib.capdec_gemm_set_row_hint(8820)
    assert lib.capdec_gemm_plan_query(12800, 2304, 768, 1, 0, 0, 0, 1) & 0xFF == 3
    lib.capdec_gemm_set_row_hint(0)
    # explicit requests are honoured
    code = lib.capdec_gemm_plan_query(12800, 768, 3072, 0, 1, 192, 3, 0)
    assert (code >> 8) & 0xFFF == 192 and code >> 20 == 3
    assert lib.capdec_gemm_plan_query(0, 5, 5, 0, 0, 0, 0, 0) < 0


def test_reference_max_seq_len_rule():
    """train.py:102-103: min(int(mean + 10 * std), max) with torch's unbiased std."""
    from capdec_b200.data import reference_max_seq_len
    from oracle import capdec_oracle as O
    lens = torch.tensor([5, 9, 12, 40, 7, 7, 33])
    caps = [torch.zeros(int(k), dtype=torch.int64) for k in lens]
    assert reference_max_seq_len(lens) == O.dataset_max_seq_len(caps) == 40
    tight = torch.tensor([10] * 500 + [11] * 500 + [300])   # one outlier: the rule cuts it off
    assert reference_max_seq_len(tight) == min(int(tight.float().mean() + tight.float().std() * 10), 300) < 300
