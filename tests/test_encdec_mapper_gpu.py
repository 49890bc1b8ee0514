"""MappingType.TransformerDecoder (transformer_mapper.TransformerEncoderDecoder, gpt2_prefix.py:167-168) on the GPU:
`model.clip_project(prefix)` against the CPU oracle and the fixture recorded from the reference's own module.
Tolerances: fp32 GEMM mode rel-L2 <= 2e-5; 1xTF32 (default) rel-L2 <= 1e-2 (six pre-LN layers deep; measured 5.4e-3)."""
import json
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import capdec_oracle as O  # noqa: E402  (checker only)

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("case", [0, 1, 2])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 1e-2)])
def test_encdec_mapper_forward(case, precision, tol):
    import capdec_b200 as cb
    c = json.loads((GOLD / "encdec_mapper.json").read_text())["cases"][case]
    sd = O.make_encdec_state_dict(seed=c["sd_seed"], prefix_length=c["P"], clip_length=c["C"], prefix_size=c["D"],
                                  num_layers=c["num_layers"])
    cfg = cb.GPT2Config(n_layer=1)
    model = cb.ClipCaptionModel(c["P"], clip_length=c["C"], prefix_dim=c["D"], num_layers=c["num_layers"],
                                mapping_type="transformer_decoder", gpt_config=cfg)
    assert isinstance(model.clip_project, cb.TransformerEncoderDecoder)
    mapper_keys = {k for k in model.state_dict() if k.startswith("clip_project.")}
    assert mapper_keys == set(sd)                                      # the reference module's key layout
    missing = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if k.startswith("clip_project.")]
    model = model.to("cuda").eval()
    x = torch.randn(c["B"], c["D"], generator=torch.Generator().manual_seed(c["x_seed"]))
    x = x / x.norm(2, -1, keepdim=True)
    cb.ops.set_precision(precision)
    try:
        out = model.clip_project(x.cuda())
    finally:
        cb.ops.set_precision("tf32")
    assert tuple(out.shape) == (c["B"], c["P"], 768)
    ref = O.encdec_mapper(sd, x, c["C"])
    err = (out.cpu().double() - ref.double()).norm() / ref.double().norm()
    assert err <= tol, float(err)
    got = out.cpu().flatten()[torch.tensor(c["idx"])].double()
    assert (got - torch.tensor(c["val"], dtype=torch.float64)).abs().max() <= 20 * tol * c["absmax"]


def test_encdec_mapper_is_inference_only():
    import capdec_b200 as cb
    from capdec_b200._lib import CapdecError
    model = cb.ClipCaptionModel(4, clip_length=4, prefix_dim=512, num_layers=1, mapping_type=cb.MappingType.TransformerDecoder,
                                gpt_config=cb.GPT2Config(n_layer=1)).to("cuda").train()
    tok = torch.randint(1, 100, (2, 8), device="cuda")
    pfx = torch.randn(2, 512, device="cuda")
    with pytest.raises(CapdecError):
        model.engine().loss_and_grads(tok, pfx)
