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


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-4), ("tf32x3", 5e-3), ("tf32", 5e-2)])
def test_encdec_mapper_trains_like_the_reference(precision, tol):
    """gpt2_prefix.py:219-243 trains a MappingType.TransformerDecoder model through autograd: loss and EVERY gradient
    (encoder, cross- and stream-attention decoder layers, prefix_const, linear, GPT-2) of one step against the oracle, whose
    TransformerEncoderDecoder forward is pinned bit-exactly on the reference module (tests/golden/encdec_mapper.json).
    The mapper gradients of a randomly initialised ReLU transformer are ill-conditioned (tests/test_scale_parity_gpu.py),
    hence the per-mode bounds of that class of tensors."""
    import capdec_b200 as cb
    P, C, D, nl = 6, 5, 512, 2
    sd = O.make_state_dict(seed=21, mapping_type="mlp", prefix_length=P, prefix_size=D, n_layer=2)
    sd = {k: v for k, v in sd.items() if not k.startswith("clip_project.")}
    sd.update(O.make_encdec_state_dict(seed=22, prefix_length=P, clip_length=C, prefix_size=D, num_layers=nl))
    tokens, prefix, _ = O.make_batch(seed=23, B=3, L=12, prefix_size=D)
    o_loss, _, o_grads = O.loss_and_grads(sd, tokens, prefix, O.make_mask(tokens, P), P, C)
    model = cb.ClipCaptionModel(P, clip_length=C, prefix_dim=D, num_layers=nl, mapping_type="transformer_decoder",
                                gpt_config=cb.GPT2Config(n_layer=2, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0))
    model.load_state_dict(sd)
    model = model.to("cuda").train()
    cb.ops.set_precision(precision)
    try:
        eng = model.engine()
        eng.zero_grads()
        tail = eng.loss_and_grads(tokens.cuda(), prefix.cuda(), mean_reduce=True)
        torch.cuda.synchronize()
        loss = tail[1].item() / tail[0].item()
        assert abs(loss - float(o_loss)) <= (2e-4 if precision == "tf32" else 3e-6) * float(o_loss)
        g = eng.grad_views()
        assert set(o_grads) <= set(g)
        worst = max(((g[k].cpu().double() - og.double()).norm() / og.double().norm().clamp_min(1e-30)).item()
                    for k, og in o_grads.items())
        assert worst <= tol, worst
        # a second step reuses the saved-activation buffers; inference (eval) still runs on the scratch path
        eng.zero_grads()
        eng.loss_and_grads(tokens.cuda(), prefix.cuda(), mean_reduce=True)
        model.eval()
        out = model.clip_project(prefix.cuda())
        ref = O.encdec_mapper(sd, prefix, C)
        assert ((out.cpu().double() - ref.double()).norm() / ref.double().norm()).item() <= (1e-2 if precision == "tf32" else 5e-5)
    finally:
        cb.ops.set_precision("tf32")
