"""CPU tests: the oracle restatement (oracle/capdec_oracle.py) against the golden vectors recorded from the
reference's own classes (oracle/pin_against_reference.py -> tests/golden/*.json).  No GPU, no /root/reference."""
import json
import math
from pathlib import Path

import pytest
import torch

from oracle import capdec_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
CASES = ["mlp_full_b4", "mlp_prefix_only_b4", "transformer_full_b2", "mlp_full_d640_b3", "transformer_prefix_only_b2",
         "transformer_short_clip_b3"]


def load_case(name):
    rec = json.loads((GOLD / f"{name}.json").read_text())
    c = rec["config"]
    sd = O.make_state_dict(seed=c["sd_seed"], mapping_type=c["mapping_type"], prefix_length=c["P"], clip_length=c["C"],
                           prefix_size=c["D"], num_layers=c["num_layers"])
    tokens, prefix, _ = O.make_batch(seed=c["batch_seed"], B=c["B"], L=40, prefix_size=c["D"], full_length=c["full_length"])
    torch.manual_seed(c["noise_torch_seed"])
    draw = torch.randn(prefix.shape)
    pfx = O.noise_injection(prefix, c["noise_variance"], noise=draw)
    return rec, c, sd, tokens, pfx


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_golden(name):
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    rec, c, sd, tokens, pfx = load_case(name)
    assert pfx[0, :8].double().tolist() == pytest.approx(rec["noised_prefix_row0"], rel=1e-6)
    mask = O.make_mask(tokens, c["P"])
    trainable = (lambda k: k.startswith("clip_project")) if c["only_prefix"] else None
    loss, logits, grads = O.loss_and_grads(sd, tokens, pfx, mask, c["P"], c["C"], trainable)
    assert float(loss) == pytest.approx(rec["loss"], rel=2e-6)
    assert list(logits.shape) == rec["logits_shape"]
    got = logits.flatten()[torch.tensor(rec["logits_idx"])].double()
    # sampled logits may fall on padded positions where the reference's additive mask arithmetic is restated exactly
    assert (got - torch.tensor(rec["logits_val"], dtype=torch.float64)).abs().max() < 2e-4 * max(1.0, rec["logits_absmax"])
    assert set(grads) == set(rec["grads"])
    for k, g in rec["grads"].items():
        t = grads[k]
        assert float(t.double().norm()) == pytest.approx(g["norm"], rel=5e-4, abs=1e-9), k
        vals = t.flatten()[torch.tensor(g["idx"])].double()
        assert (vals - torch.tensor(g["val"], dtype=torch.float64)).abs().max() <= 1e-3 * g["norm"] + 1e-9, k


def test_mask_does_not_change_the_loss():
    """SURVEY §8c probe (i): right padding + causal attention => the padding mask cannot reach a consumed logit."""
    rec, c, sd, tokens, pfx = load_case("mlp_full_b4")
    assert rec["loss_without_mask"] == pytest.approx(rec["loss"], rel=1e-6)
    lg = O.clipcap_forward(sd, tokens, pfx, None, c["P"], c["C"])
    assert float(O.caption_loss(lg, tokens, c["P"])) == pytest.approx(rec["loss"], rel=2e-6)


def test_noise_injection_golden():
    rec = json.loads((GOLD / "noise_injection.json").read_text())
    _, prefix, _ = O.make_batch(seed=rec["batch_seed"], B=6, prefix_size=640)
    x = prefix * 3.0
    off = torch.randn(1, 640, generator=torch.Generator().manual_seed(rec["offset_seed"])) * 0.05
    kws = {"plain": {}, "offset": {"modality_offset": off}, "dont_norm": {"dont_norm": True}, "zero_var": {"variance": 0.0}}
    for tag, kw in kws.items():
        var = kw.pop("variance", 0.016)
        torch.manual_seed(rec["seed"])
        y = O.noise_injection(x, var, noise=torch.randn(x.shape), **kw)
        g = rec["cases"][tag]
        assert y[0, :6].double().tolist() == pytest.approx(g["row0"], rel=1e-6)
        assert float(y.double().sum()) == pytest.approx(g["sum"], rel=1e-6)
    # variance 0 returns the input itself, unnormalised (train.py:28-29)
    assert O.noise_injection(x, 0.0) is x


def test_hf_adamw_and_schedule_restatement():
    """HF-4.24 AdamW: eps added before bias correction; lr == 0 on the very first step of the warm-up schedule."""
    p = torch.tensor([1.0, -2.0]); g = torch.tensor([0.5, 0.25])
    m = torch.zeros(2); v = torch.zeros(2)
    O.hf_adamw_step(p, g, m, v, step=1, lr=0.1)
    denom = (0.001 * g * g).sqrt() + 1e-6
    expect = torch.tensor([1.0, -2.0]) - 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / denom
    assert torch.allclose(p, expect, rtol=1e-6)
    assert O.linear_warmup_lr(2e-5, 0, 5000, 10000) == 0.0
    assert O.linear_warmup_lr(2e-5, 2500, 5000, 10000) == pytest.approx(1e-5)
    assert O.linear_warmup_lr(2e-5, 7500, 5000, 10000) == pytest.approx(1e-5)
    w = torch.nn.Parameter(torch.ones(3))
    opt = O.HFAdamW([w], lr=0.01)
    w.grad = torch.ones(3)
    opt.step()
    assert torch.allclose(w.detach(), torch.full((3,), 1 - 0.01 * math.sqrt(0.001) / 0.1 * 0.1 / (math.sqrt(0.001) + 1e-6)), rtol=1e-5)


def test_hf_adamw_equals_torch_adam_with_rescaled_eps():
    """Independent pin of the one restated piece whose original class is gone (transformers.AdamW, train.py:326):
    HF-4.24 puts eps INSIDE the bias-corrected step, p -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps), while
    torch.optim.Adam computes p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps').  The two are the same update
    when eps' = eps / sqrt(1-b2^t), so torch's own Adam, fed that eps every step, must reproduce the oracle."""
    torch.manual_seed(11)
    w0 = torch.randn(257)
    w_hf, w_pt = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(w0.clone())
    hf = O.HFAdamW([w_hf], lr=3e-3)
    pt = torch.optim.Adam([w_pt], lr=3e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0)
    for t in range(1, 41):
        g = torch.randn(257) * (10.0 ** torch.randint(-6, 1, (257,)).float())    # gradient scales across 7 decades
        w_hf.grad, w_pt.grad = g.clone(), g.clone()
        pt.param_groups[0]["eps"] = 1e-6 / math.sqrt(1.0 - 0.999 ** t)
        hf.step(); pt.step()
        assert torch.allclose(w_hf.detach(), w_pt.detach(), rtol=2e-6, atol=1e-8), t
    # and the plain torch AdamW defaults (eps 1e-8 after the correction) are NOT the same optimizer: tiny gradients differ
    w_a, w_b = torch.nn.Parameter(torch.ones(4)), torch.nn.Parameter(torch.ones(4))
    a, b = O.HFAdamW([w_a], lr=1e-2), torch.optim.AdamW([w_b], lr=1e-2, weight_decay=0.0)
    w_a.grad, w_b.grad = torch.full((4,), 1e-7), torch.full((4,), 1e-7)
    a.step(); b.step()
    assert (w_a - w_b).abs().max() > 1e-3


def test_schedule_restatement_matches_the_installed_transformers():
    """get_linear_schedule_with_warmup still exists in the installed transformers: the lr the oracle uses for optimizer
    step k equals the lr the real scheduler has set after k scheduler.step() calls (train.py:328-330,352-353)."""
    transformers = pytest.importorskip("transformers")
    w = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([w], lr=2e-5)
    sch = transformers.get_linear_schedule_with_warmup(opt, num_warmup_steps=7, num_training_steps=30)
    for k in range(34):
        assert opt.param_groups[0]["lr"] == pytest.approx(O.linear_warmup_lr(2e-5, k, 7, 30), rel=1e-12, abs=1e-18), k
        opt.step(); sch.step()


def test_oracle_beam_search_reproduces_the_reference_ids():
    """tests/golden/beam.json holds the ids the reference's own generate_beam produced (pin_generate_beam)."""
    rec = json.loads((GOLD / "beam.json").read_text())
    c = rec["config"]
    sd = O.make_state_dict(seed=c["sd_seed"], mapping_type="mlp", prefix_length=c["P"], prefix_size=c["D"],
                           weight_std=c["weight_std"])
    for case in rec["cases"][3:5]:  # one sharp-temperature run and its early-stop variant (seconds each on CPU)
        _, prefix, _ = O.make_batch(seed=case["batch_seed"], B=1, prefix_size=c["D"])
        ids, scores, lens = O.generate_beam(sd, O.mlp_mapper(sd, prefix).view(1, c["P"], -1), c["beam_size"],
                                            c["entry_length"], case["temperature"], case["stop_token_index"])
        assert ids == case["ids"] and lens == case["seq_lengths"]
        assert max(abs(a - b) for a, b in zip(scores, case["scores"])) < 1e-5


def test_oracle_greedy_decode_reproduces_the_reference_ids():
    """tests/golden/greedy.json holds the ids the reference's own generate2 produced (pin_generate2)."""
    rec = json.loads((GOLD / "greedy.json").read_text())
    c = rec["config"]
    sd = O.make_state_dict(seed=c["sd_seed"], mapping_type="mlp", prefix_length=c["P"], prefix_size=c["D"],
                           weight_std=c["weight_std"])
    for case in rec["cases"][:6]:
        _, prefix, _ = O.make_batch(seed=case["batch_seed"], B=1, prefix_size=c["D"])
        with torch.no_grad():
            embed = O.mlp_mapper(sd, prefix).view(1, c["P"], -1)
            ids = O.generate2(sd, embed, entry_length=case["entry_length"], temperature=case["temperature"],
                              stop_token_index=case["stop_token_index"])
            # greedy == the best (only) beam of a 1-beam search cut at the stop token: the identity the CUDA path relies on
            beams, _, _ = O.generate_beam(sd, embed, beam_size=1, entry_length=case["entry_length"],
                                          temperature=case["temperature"], stop_token_index=case["stop_token_index"])
        assert ids == case["ids"]
        assert beams[0] == case["ids"] or 764 in beams[0]


def test_oracle_dataset_item_reproduces_the_reference_dataset():
    """train.py:52-72 restated (oracle.dataset_item) vs what the reference's own ClipCocoDataset + DataLoader returned
    (tests/golden/datafeed.json): token ids exact, mask exact, prefix to fp32 round-off."""
    rec = json.loads((GOLD / "datafeed.json").read_text())
    for c in rec["cases"]:
        caps, cap2emb, table = O.make_caption_table(seed=c["table_seed"], n=c["n"], n_emb=c["n_emb"], half=c["half"])
        L, P = c["max_seq_len"], c["prefix_length"]
        toks, masks, pfx = [], [], []
        for it in c["idx"]:
            t, m, p = O.dataset_item(caps, cap2emb, table, it, L, P, c["normalize_prefix"])
            toks.append(t), masks.append(m), pfx.append(p)
        tokens, mask, prefix = torch.stack(toks), torch.stack(masks), torch.stack(pfx)
        assert tokens.tolist() == c["tokens"]
        assert mask.sum(1).tolist() == c["mask_sum_rows"]
        assert torch.equal(mask[:, P:], torch.stack([torch.arange(L) < min(len(caps[i]), L) for i in c["idx"]]).float())
        assert str(prefix.dtype).replace("torch.", "") == c["prefix_dtype"]
        assert prefix.float().norm(2, -1).double().tolist() == pytest.approx(c["prefix_row_norms"], rel=1e-6)
        assert (prefix[:, :6].double() - torch.tensor(c["prefix_head"])).abs().max() < 1e-6
        assert float(prefix.double().sum()) == pytest.approx(c["prefix_sum"], rel=1e-5, abs=1e-5)


def test_oracle_encdec_mapper_reproduces_the_reference_module():
    """oracle.encdec_mapper vs values recorded from transformer_mapper.TransformerEncoderDecoder itself."""
    rec = json.loads((GOLD / "encdec_mapper.json").read_text())
    for c in rec["cases"]:
        sd = O.make_encdec_state_dict(seed=c["sd_seed"], prefix_length=c["P"], clip_length=c["C"], prefix_size=c["D"],
                                      num_layers=c["num_layers"])
        x = torch.randn(c["B"], c["D"], generator=torch.Generator().manual_seed(c["x_seed"]))
        x = x / x.norm(2, -1, keepdim=True)
        out = O.encdec_mapper(sd, x, c["C"])
        assert tuple(out.shape) == (c["B"], c["P"], 768)
        got = out.flatten()[torch.tensor(c["idx"])].double()
        assert (got - torch.tensor(c["val"], dtype=torch.float64)).abs().max() <= 1e-5 * c["absmax"]
        assert float(out.double().norm()) == pytest.approx(c["norm"], rel=1e-6)
