"""CPU check of the drop-in launchers: importing the reference's unmodified train.py through
launchers/run_train_b200.py rebinds the names train.py resolves at call time to the capdec_b200 classes, and the
reference's own argparse surface (the --noise_variance / --mapping_type / --only_prefix flags) is untouched.
Needs the reference checkout, which exists only in the build container: skipped elsewhere."""
import importlib.util
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

pytestmark = pytest.mark.skipif(not (REF / "train.py").exists(), reason="reference checkout not mounted")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, ROOT / "launchers" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_train_launcher_rebinds_the_reference_globals():
    import capdec_b200 as cb
    train = _load("run_train_b200").bind(str(REF))
    assert train.ClipCaptionModel is cb.ClipCaptionModel and train.ClipCaptionPrefix is cb.ClipCaptionPrefix
    assert train.noise_injection is cb.noise_injection and train.AdamW is cb.AdamW
    assert train.get_linear_schedule_with_warmup is cb.get_linear_schedule_with_warmup
    # train.py:446 maps the CLI strings onto MappingType members by attribute name
    assert train.MappingType.MLP.value == "mlp" and train.MappingType.Transformer.value == "transformer"
    assert train.main.__code__.co_filename.startswith(str(REF))            # the reference's own main(), unmodified


def test_train_launcher_keeps_the_reference_cli():
    import os
    env = dict(os.environ, CAPDEC_REFERENCE_DIR=str(REF))
    r = subprocess.run([sys.executable, str(ROOT / "launchers" / "run_train_b200.py"), "--help"], capture_output=True,
                       text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for flag in ("--noise_variance", "--mapping_type", "--only_prefix", "--prefix_length_clip", "--uniform_noise",
                 "--dont_norm", "--add_modality_offset", "--val_pt", "--pretrain_weights"):
        assert flag in r.stdout, flag


def test_fast_flag_rebinds_the_training_loop():
    """`--fast`: the reference's main() keeps parsing its own flags and building its own dataset / model, but the
    module-global `train` it calls (train.py:466) is capdec_b200.fit.train, which has the reference's signature."""
    import inspect
    import capdec_b200 as cb
    train = _load("run_train_b200").bind(str(REF), fast=True)
    assert train.train is cb.fit.train
    ours = list(inspect.signature(cb.fit.train).parameters)[:6]
    assert ours == ["dataset", "model", "args", "warmup_steps", "output_dir", "output_prefix"]
    sig = inspect.signature(cb.fit.train)
    assert sig.parameters["warmup_steps"].default == 5000 and sig.parameters["output_dir"].default == "."


def test_predictions_launcher_binds_model_and_decoders():
    """launchers/run_predictions_b200.py: predictions_runner.py imports its model from `gpt2_prefix` and its decoders from
    `gpt2_prefix_eval` (predictions_runner.py:7,13).  After bind() the reference module resolves all of them to
    capdec_b200 (run in a subprocess: the reference's optional imaging / CLIP dependencies are stubbed there)."""
    code = r"""
import sys, types, importlib.util
for name in ("clip", "pycocotools", "pycocotools.coco", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["pycocotools.coco"].COCO = object
spec = importlib.util.spec_from_file_location("run_predictions_b200", sys.argv[1])
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
pr = m.bind(sys.argv[2])
import capdec_b200 as cb, gpt2_prefix_eval as ev
assert pr.ClipCaptionModel is cb.ClipCaptionModel and pr.MappingType is cb.MappingType
assert pr.generate_beam is cb.generate_beam and ev.generate_beam is cb.generate_beam
assert pr.generate2 is cb.generate2 and ev.generate2 is cb.generate2
assert all(hasattr(pr.MappingType, n) for n in ("MLP", "TransformerEncoder", "TransformerDecoder"))    # predictions_runner.py:457-458
assert pr.make_preds.__code__.co_filename.startswith(sys.argv[2])        # the reference's own loop, unmodified
assert sys.modules["gpt2_prefix"].ClipCocoDataset.__name__ == "ClipCocoDataset"
from others.supervised_embedding_bridger import get_map_to_text_space_using_modality_bridger as g    # predictions_runner.py:183
assert g is cb.get_map_to_text_space_using_modality_bridger
print("bound")
"""
    r = subprocess.run([sys.executable, "-c", code, str(ROOT / "launchers" / "run_predictions_b200.py"), str(REF)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "bound" in r.stdout, r.stderr[-3000:]


def test_checkpoint_loads_strictly_into_the_reference_class(tmp_path):
    """A checkpoint written from our class (`torch.save(model.state_dict())`, as train.py:359-371 and capdec_b200.fit.train
    do) loads with the reference's strict `model.load_state_dict(torch.load(...))` (train.py:457, predictions_runner.py:461)
    into the reference's OWN ClipCaptionModel built on the installed transformers, and the weights arrive bit-exactly.
    Run in a subprocess (the reference module needs the AdamW / from_pretrained / device shims of SURVEY §8c)."""
    code = r"""
import sys, torch, transformers
sys.path.insert(0, sys.argv[3])
import capdec_b200 as cb
ours = cb.ClipCaptionModel(10, prefix_size=512, mapping_type=cb.MappingType.MLP, gpt_config=cb.GPT2Config())
torch.save(ours.state_dict(), sys.argv[2])
transformers.GPT2LMHeadModel.from_pretrained = staticmethod(lambda name, *a, **k: transformers.GPT2LMHeadModel(transformers.GPT2Config()))  # shim 2
sys.path.insert(0, sys.argv[1])
from transformers import GPT2Tokenizer, get_linear_schedule_with_warmup     # resolve the lazy attributes first
sys.modules["transformers"].AdamW = torch.optim.AdamW        # shim 1 (the import at train.py:6): set right before the import
import train
ref = train.ClipCaptionModel(10, prefix_size=512, mapping_type=train.MappingType.MLP)
ref.load_state_dict(torch.load(sys.argv[2], map_location="cpu"))            # strict, like the reference
rsd = ref.state_dict()
for k, v in ours.state_dict().items():
    assert torch.equal(rsd[k], v), k
print("strict-ok", len(rsd))
"""
    r = subprocess.run([sys.executable, "-c", code, str(REF), str(tmp_path / "ck.pt"), str(ROOT)], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "strict-ok" in r.stdout, r.stderr[-3000:]
