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


def _provide_adamw(cb):
    """`transformers.AdamW` left the library in 4.5x; the reference imports it by name.  transformers' lazy top-level module
    REPLACES itself in sys.modules on the first real import, so force that first and then set the name on the live module."""
    from transformers import GPT2LMHeadModel, GPT2Tokenizer  # noqa: F401  (what the reference imports next to AdamW)
    live = sys.modules["transformers"]
    if not hasattr(live, "AdamW"):
        live.AdamW = cb.AdamW


def _reference_dataset_class(ref_dir: str):
    """gpt2_prefix_eval.py:7 imports `ClipCocoDataset` from gpt2_prefix next to the model class.  It is the reference's own
    data loading (out of scope here), so hand back the reference's class when its module imports (it needs `clip` etc.)."""
    import importlib.util
    try:
        spec = importlib.util.spec_from_file_location("_capdec_reference_gpt2_prefix", str(Path(ref_dir) / "gpt2_prefix.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.ClipCocoDataset
    except Exception as e:                              # missing optional dependency of the reference: fail at USE, not import
        err = repr(e)

        class ClipCocoDataset:                          # noqa: D401 - placeholder with the reference's name
            def __init__(self, *a, **k):
                raise ImportError(f"the reference's gpt2_prefix.ClipCocoDataset could not be imported: {err}")
        return ClipCocoDataset


def bind(ref_dir: str):
    import capdec_b200 as cb
    _provide_adamw(cb)                                 # gpt2_prefix_eval.py:2 imports transformers.AdamW
    shim = types.ModuleType("gpt2_prefix")
    shim.ClipCaptionModel = cb.ClipCaptionModel
    shim.ClipCaptionPrefix = cb.ClipCaptionPrefix
    shim.MappingType = cb.MappingType
    shim.MLP = cb.MLP
    sys.path.insert(0, ref_dir)
    shim.ClipCocoDataset = _reference_dataset_class(ref_dir)     # gpt2_prefix_eval.py:7
    sys.modules["gpt2_prefix"] = shim              # predictions_runner.py:7
    # --modality_bridger (predictions_runner.py:182-184): `from others.supervised_embedding_bridger import
    # get_map_to_text_space_using_modality_bridger` resolves to the 8-GEMM stack on our kernels (same weights file)
    others = sys.modules.setdefault("others", types.ModuleType("others"))
    if not hasattr(others, "__path__"):
        others.__path__ = [str(Path(ref_dir) / "others")]
    bridger = types.ModuleType("others.supervised_embedding_bridger")
    bridger.get_map_to_text_space_using_modality_bridger = cb.get_map_to_text_space_using_modality_bridger
    bridger.MLP = cb.ModalityBridger
    sys.modules["others.supervised_embedding_bridger"] = bridger
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
