#!/usr/bin/env python
"""Run the reference's own, unmodified `train.py` on the capdec_b200 kernels.

    CAPDEC_REFERENCE_DIR=/path/to/CapDec python launchers/run_train_b200.py --data ... --noise_variance 0.016 \
        --mapping_type mlp --only_prefix ...            # train.py's own flags, untouched (train.py:397-416)

`train.py` resolves ClipCaptionModel / ClipCaptionPrefix / noise_injection / MappingType / AdamW /
get_linear_schedule_with_warmup as module globals at call time (train.py:326-328,347,446-454), so rebinding them on the
imported module is the whole integration: no reference file is edited.  What then runs per batch is exactly
train.py:345-354 — `model(tokens, prefix, mask)` enters `Engine.logits_autograd`, `loss.backward()` our hand-written
backward through one autograd.Function, `optimizer.step()` the fused HF-semantics AdamW kernel; checkpoints keep the
reference's key layout (train.py:359-371).

`--fast` (or CAPDEC_FAST=1) additionally rebinds the reference's `train()` loop (train.py:307-393) to
`capdec_b200.fit.train`: same flags, files and epoch bookkeeping, but each batch is one `Trainer.step_from` on a
device-resident dataset (CUDA-graph replay, no [B,T,V] logits, no per-step host sync); under
`torchrun --nproc-per-node N` it trains data-parallel.
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _provide_adamw(cb):
    """`transformers.AdamW` left the library in 4.5x; the reference imports it by name.  transformers' lazy top-level module
    REPLACES itself in sys.modules on the first real import, so force that first and then set the name on the live module."""
    from transformers import GPT2LMHeadModel, GPT2Tokenizer  # noqa: F401  (what the reference imports next to AdamW)
    live = sys.modules["transformers"]
    if not hasattr(live, "AdamW"):
        live.AdamW = cb.AdamW


def bind(ref_dir: str, fast: bool = False):
    import capdec_b200 as cb
    _provide_adamw(cb)                             # train.py:6 imports transformers.AdamW
    sys.path.insert(0, ref_dir)
    import train                                   # the unmodified reference module
    train.ClipCaptionModel = cb.ClipCaptionModel   # train.py:447-454 construct these by name
    train.ClipCaptionPrefix = cb.ClipCaptionPrefix
    train.noise_injection = cb.noise_injection     # train.py:347
    train.MappingType = cb.MappingType             # train.py:446
    train.AdamW = cb.AdamW                         # train.py:326
    train.get_linear_schedule_with_warmup = cb.get_linear_schedule_with_warmup   # train.py:328
    if fast:   # main() calls the module-global `train(dataset, model, args, ...)` (train.py:466): same signature, fast path
        train.train = cb.fit.train
    return train


def main():
    ref_dir = os.environ.get("CAPDEC_REFERENCE_DIR", "")
    if not ref_dir or not (Path(ref_dir) / "train.py").exists():
        sys.exit("set CAPDEC_REFERENCE_DIR to a checkout of DavidHuji/CapDec (the directory that holds train.py)")
    fast = os.environ.get("CAPDEC_FAST", "0") == "1"
    if "--fast" in sys.argv:                       # our only extra flag: removed before the reference's argparse runs
        sys.argv.remove("--fast")
        fast = True
    if fast and int(os.environ.get("WORLD_SIZE", "1")) > 1:   # torchrun: data parallel over the box's GPUs
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    train = bind(ref_dir, fast=fast)
    return train.main()


if __name__ == "__main__":
    sys.exit(main())
