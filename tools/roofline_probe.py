"""bench.py's roofline probe alone (fused-QKV GEMM 12800 x 2304 x 768 + bias, 12 rotating operand sets, CUDA events)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import capdec_b200 as cb  # noqa: E402

for _ in range(3):
    tf, ms = bench.qkv_gemm_roofline(cb, torch)
    print(f"{tf:.1f} TF/s, {ms * 1e3:.1f} us, frac {tf / (bench.peaks()['bf16'] / 2):.3f}", flush=True)
