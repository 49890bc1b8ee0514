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


def test_no_cpu_fallback():
    """CPU tensors raise instead of silently computing somewhere else."""
    from capdec_b200 import ops
    from capdec_b200._lib import CapdecError
    a = torch.zeros(4, 4)
    with pytest.raises(CapdecError):
        ops.gemm(a, 0, a, 0, a, 4, 4, 4)
