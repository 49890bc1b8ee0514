"""ClipCaptionModel starts GPT-2 from the pretrained 'gpt2' checkpoint like the reference (train.py:266,
gpt2_prefix.py:162: `GPT2LMHeadModel.from_pretrained('gpt2')`), and says so loudly when it cannot.  CPU only: a fake
`from_pretrained` stands in for the hub (no network here)."""
import types

import pytest
import torch


def _fake_hf_state_dict(cb, seed=5, with_mask_buffers=True):
    """A GPT-2-small state_dict in the HF layout with recognisable values (and the 4.24-era mask buffers)."""
    g = torch.Generator().manual_seed(seed)
    ref = cb.model.GPT2LMHead(cb.GPT2Config())
    sd = {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in ref.state_dict().items()}
    sd["lm_head.weight"] = sd["transformer.wte.weight"]
    if with_mask_buffers:
        for i in range(12):
            sd[f"transformer.h.{i}.attn.bias"] = torch.ones(1, 1, 8, 8, dtype=torch.uint8)
            sd[f"transformer.h.{i}.attn.masked_bias"] = torch.tensor(-1e4)
    return sd


def test_default_constructor_loads_pretrained_gpt2(monkeypatch):
    import transformers
    import capdec_b200 as cb
    monkeypatch.setenv("CAPDEC_GPT2_PRETRAINED", "1")
    sd = _fake_hf_state_dict(cb)
    calls = []

    def fake_from_pretrained(name, *a, **k):
        calls.append(name)
        return types.SimpleNamespace(state_dict=lambda: sd)

    monkeypatch.setattr(transformers.GPT2LMHeadModel, "from_pretrained", staticmethod(fake_from_pretrained))
    for cls in (cb.ClipCaptionModel, cb.ClipCaptionPrefix):
        m = cls(10, prefix_size=512, mapping_type=cb.MappingType.MLP)      # the reference's call, train.py:447-454
        assert calls[-1] == "gpt2" and m.gpt_init.startswith("pretrained")
        own = m.state_dict()
        for k, v in sd.items():
            if k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"):
                assert "gpt." + k not in own
                continue
            assert torch.equal(own["gpt." + k], v), k
        assert own["gpt.lm_head.weight"].data_ptr() == own["gpt.transformer.wte.weight"].data_ptr()   # still tied
    # an explicit architecture (tests, synthetic benchmark) is a request for random init: the hub is not consulted
    n = len(calls)
    m = cb.ClipCaptionModel(10, prefix_size=512, gpt_config=cb.GPT2Config())
    assert len(calls) == n and m.gpt_init.startswith("random")
    m = cb.ClipCaptionModel(10, prefix_size=512, pretrained=False)
    assert len(calls) == n


def test_unavailable_checkpoint_falls_back_loudly(monkeypatch, capsys):
    import transformers
    import capdec_b200 as cb
    monkeypatch.setenv("CAPDEC_GPT2_PRETRAINED", "1")

    def offline(name, *a, **k):
        raise OSError("We couldn't connect to 'https://huggingface.co'")

    monkeypatch.setattr(transformers.GPT2LMHeadModel, "from_pretrained", staticmethod(offline))
    m = cb.ClipCaptionModel(10, prefix_size=512)
    err = capsys.readouterr().err
    assert "RANDOM weights" in err and "from_pretrained('gpt2') failed" in err
    assert "unavailable" in m.gpt_init
    w = m.gpt.transformer.h[0].attn.c_attn.weight
    assert 0.015 < w.std().item() < 0.025           # HF-style N(0, 0.02) init, not zeros / garbage
