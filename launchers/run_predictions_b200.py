#!/usr/bin/env python
"""Run the reference's own `predictions_runner.py` with the capdec_b200 model classes and KV-cached beam search.

    CAPDEC_REFERENCE_DIR=/path/to/CapDec python launchers/run_predictions_b200.py <predictions_runner.py flags>

predictions_runner.py:7 does `from gpt2_prefix import ClipCaptionModel, MappingType`; providing a `gpt2_prefix` module
with our classes before it is imported is the binding (constructor keyword `prefix_dim`, the three mapping types of
gpt2_prefix.py:15-18 and the checkpoint key layout are the reference's).  `gpt2_prefix_eval.generate_beam`
(gpt2_prefix_eval.py:50-115) is rebound to the batched, KV-cached decoder with identical beam semantics
(tests/test_decode_gpu.py pins its token ids against the reference function's own output); `generate2`
(gpt2_prefix_eval.py:118-198, greedy decoding) is rebound to the same decoder with one beam.
"""
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def bind(ref_dir: str):
    import capdec_b200 as cb
    shim = types.ModuleType("gpt2_prefix")
    shim.ClipCaptionModel = cb.ClipCaptionModel
    shim.ClipCaptionPrefix = cb.ClipCaptionPrefix
    shim.MappingType = cb.MappingType
    shim.MLP = cb.MLP
    sys.modules["gpt2_prefix"] = shim              # predictions_runner.py:7
    sys.path.insert(0, ref_dir)
    import gpt2_prefix_eval
    gpt2_prefix_eval.generate_beam = cb.generate_beam            # gpt2_prefix_eval.py:50-115
    import predictions_runner
    if hasattr(predictions_runner, "generate_beam"):
        predictions_runner.generate_beam = cb.generate_beam      # predictions_runner.py:232 calls it by name
    # the non-beam branch (generate2, predictions_runner.py:234 / gpt2_prefix_eval.py:118-198): greedy decoding on the same
    # KV-cached decoder (ids pinned on the reference function, tests/test_decode_gpu.py::test_generate2_*);
    # CAPDEC_FAST_GREEDY=0 keeps the reference's own generate2 running through `model.gpt(inputs_embeds=...)`
    if os.environ.get("CAPDEC_FAST_GREEDY", "1") != "0":
        gpt2_prefix_eval.generate2 = cb.generate2
        if hasattr(predictions_runner, "generate2"):
            predictions_runner.generate2 = cb.generate2
    return predictions_runner


def main():
    ref_dir = os.environ.get("CAPDEC_REFERENCE_DIR", "")
    if not ref_dir or not (Path(ref_dir) / "predictions_runner.py").exists():
        sys.exit("set CAPDEC_REFERENCE_DIR to a checkout of DavidHuji/CapDec (the directory that holds predictions_runner.py)")
    pr = bind(ref_dir)
    return pr.main()


if __name__ == "__main__":
    sys.exit(main())
