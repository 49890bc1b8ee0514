"""`--modality_bridger` (predictions_runner.py:182-184,225-227; others/supervised_embedding_bridger.py:87-108): the oracle
restatement is pinned bit-exactly on the reference's own MLP class (oracle/pin_against_reference.py::pin_modality_bridger,
tests/golden/bridger.json); the CUDA stack (8 tcgen05 GEMMs with fused bias + ReLU) is checked against the oracle."""
import json
from pathlib import Path

import pytest
import torch

from oracle import capdec_oracle as O  # checker only

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "bridger.json").read_text())


def _inputs():
    g = torch.Generator().manual_seed(GOLD["cases"][0]["x_seed"])
    x = torch.randn(5, 640, generator=g)
    return x / x.norm(2, -1, keepdim=True)


def test_oracle_bridger_matches_the_reference_golden():
    rec = GOLD["cases"][0]
    y = O.modality_bridger(O.make_bridger_state_dict(seed=9), _inputs())
    assert y[0, :16].double().tolist() == pytest.approx(rec["y_row0"], abs=1e-6)
    assert y.norm(2, -1).double().tolist() == pytest.approx(rec["y_norms"], rel=1e-6)


def test_bridger_parameter_layout_is_the_reference_checkpoint_layout():
    import capdec_b200 as cb
    m = cb.ModalityBridger()
    sd = O.make_bridger_state_dict(seed=9)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    m.load_state_dict(sd)                                            # strict
    assert torch.equal(m.layers[0].weight.detach(), sd["layers.0.weight"])
    fresh = cb.ModalityBridger()                                     # :94-97: identity weights at construction
    assert torch.equal(fresh.layers[3].weight.detach(), torch.eye(640))
    with pytest.raises(cb._lib.CapdecError):
        m(torch.zeros(1, 640))                                       # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("fp32", 2e-6), ("tf32x3", 1e-4), ("tf32", 2e-2)])
def test_bridger_cuda_matches_oracle(mode, tol):
    import capdec_b200 as cb
    sd = O.make_bridger_state_dict(seed=9)
    x = _inputs()
    ref = O.modality_bridger(sd, x)
    m = cb.ModalityBridger()
    m.load_state_dict(sd)
    m = m.to("cuda").eval()
    cb.ops.set_precision(mode)
    try:
        y = m(x.cuda()).cpu()
        y1 = m(x[:1].cuda()).cpu()          # one image at a time, as predictions_runner.py calls it
    finally:
        cb.ops.set_precision("tf32")
    rel = ((y.double() - ref.double()).norm() / ref.double().norm()).item()
    assert rel <= tol, rel
    assert ((y1.double() - ref[:1].double()).norm() / ref[:1].double().norm()).item() <= tol
