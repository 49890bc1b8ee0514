"""CPU checks of the drop-in class surface and checkpoint layout — construction and state_dict handling need
no GPU (only running the model does, and that raises on a CPU model: there is no fallback).

The oracle's `make_state_dict` is the layout witness: oracle/pin_against_reference.py loads it into the REFERENCE's own
classes with `load_state_dict(strict=True)` (train.ClipCaptionModel / gpt2_prefix.ClipCaptionModel), so a capdec_b200 model
whose state_dict has exactly those names and shapes reads and writes the reference's checkpoints (train.py:359-371,
predictions_runner.py:461)."""
import pytest
import torch

from oracle import capdec_oracle as O  # layout witness / checker only


def _shapes(sd):
    return {k: tuple(v.shape) for k, v in sd.items()}


@pytest.mark.parametrize("kw", [
    dict(mapping_type="mlp", prefix_length=10, prefix_size=512),
    dict(mapping_type="mlp", prefix_length=10, prefix_size=640),
    dict(mapping_type="transformer", prefix_length=12, clip_length=6, prefix_size=512, num_layers=2),
    dict(mapping_type="transformer", prefix_length=40, clip_length=40, prefix_size=640, num_layers=8),
])
def test_state_dict_names_and_shapes_equal_the_reference_layout(kw):
    import capdec_b200 as cb
    sd = O.make_state_dict(seed=2, **kw)
    ctor = dict(prefix_size=kw["prefix_size"], mapping_type=kw["mapping_type"])
    if "clip_length" in kw:
        ctor.update(clip_length=kw["clip_length"], num_layers=kw["num_layers"])
    for cls in (cb.ClipCaptionModel, cb.ClipCaptionPrefix):
        model = cls(kw["prefix_length"], **ctor)
        ours = model.state_dict()
        assert _shapes(ours) == _shapes(sd)
        assert "gpt.lm_head.weight" in ours
        assert ours["gpt.lm_head.weight"].data_ptr() == ours["gpt.transformer.wte.weight"].data_ptr()   # tied (HF:...:646-651)
        model.load_state_dict(sd)                                   # strict
        back = model.state_dict()
        assert all(torch.equal(back[k], sd[k]) for k in sd)
        # checkpoints written under transformers 4.24 carry the causal-mask buffers: accepted on load, emitted on request
        old = model.state_dict_hf424()
        assert set(old) - set(sd) and all(k.endswith((".attn.bias", ".attn.masked_bias")) for k in set(old) - set(sd))
        model.load_state_dict(old)


def test_state_dict_follows_the_installed_transformers_mask_buffers(monkeypatch):
    """transformers 4.24 (the reference's pin) keeps `attn.bias` / `attn.masked_bias` in GPT-2's state_dict and the
    reference loads strictly: `state_dict()` emits them exactly when the environment's own GPT2LMHeadModel has them."""
    import capdec_b200 as cb
    m = cb.ClipCaptionModel(10, prefix_size=512)
    monkeypatch.setenv("CAPDEC_CKPT_MASK_BUFFERS", "0")
    plain = m.state_dict()
    assert len(plain) == 153
    monkeypatch.setenv("CAPDEC_CKPT_MASK_BUFFERS", "1")
    old = m.state_dict()
    assert len(old) == 153 + 24
    assert old["gpt.transformer.h.11.attn.bias"].shape == (1, 1, 1024, 1024) and old["gpt.transformer.h.11.attn.bias"].dtype == torch.uint8
    assert old["gpt.transformer.h.0.attn.bias"][0, 0, 5, :8].tolist() == [1, 1, 1, 1, 1, 1, 0, 0]      # causal (lower-triangular)
    assert float(old["gpt.transformer.h.0.attn.masked_bias"]) == -1e4
    m.load_state_dict(old)                                        # and accepted back (strict)
    monkeypatch.delenv("CAPDEC_CKPT_MASK_BUFFERS")
    import transformers
    tiny = transformers.GPT2LMHeadModel(transformers.GPT2Config(n_layer=1, n_embd=8, n_head=1, vocab_size=8, n_positions=8))
    assert cb.model.hf_expects_mask_buffers() == ("transformer.h.0.attn.bias" in tiny.state_dict())


def test_prefix_only_model_trains_the_mapper_alone():
    """ClipCaptionPrefix (train.py:276-284): parameters() yields clip_project only; train() leaves GPT-2 in eval mode."""
    import capdec_b200 as cb
    m = cb.ClipCaptionPrefix(10, prefix_size=512, mapping_type=cb.MappingType.MLP)
    n_map = sum(p.numel() for p in m.clip_project.parameters())
    assert sum(p.numel() for p in m.parameters()) == n_map == 31_468_800
    m.train()
    assert m.training and m.clip_project.training and not m.gpt.training
    full = cb.ClipCaptionModel(10, prefix_size=512)
    assert sum(p.numel() for p in full.parameters()) == 31_468_800 + 124_439_808
    full.train()
    assert full.gpt.training
    assert not m.gpt_trainable() and full.gpt_trainable()


def test_constructor_flavours_and_mapping_type_spellings():
    """train.py:262-263 (prefix_size, MappingType.MLP/Transformer) and gpt2_prefix.py:157-158 (prefix_dim,
    MappingType.TransformerEncoder/TransformerDecoder), plus the CLI strings train.py:446 / predictions_runner.py:457 map."""
    import capdec_b200 as cb
    a = cb.ClipCaptionModel(10, prefix_size=512, mapping_type=cb.MappingType.MLP)
    b = cb.ClipCaptionModel(10, prefix_dim=512, mapping_type="mlp")
    assert _shapes(a.state_dict()) == _shapes(b.state_dict()) and a.prefix_length == b.prefix_length == 10
    assert a.gpt_embedding_size == 768
    t1 = cb.ClipCaptionModel(8, clip_length=4, prefix_size=512, num_layers=1, mapping_type=cb.MappingType.Transformer)
    t2 = cb.ClipCaptionModel(8, clip_length=4, prefix_dim=512, num_layers=1, mapping_type=cb.MappingType.TransformerEncoder)
    t3 = cb.ClipCaptionModel(8, clip_length=4, prefix_dim=512, num_layers=1, mapping_type="transformer_encoder")
    assert _shapes(t1.state_dict()) == _shapes(t2.state_dict()) == _shapes(t3.state_dict())
    assert "clip_project.prefix_const" in t1.state_dict() and t1.state_dict()["clip_project.prefix_const"].shape == (8, 768)
    d = cb.ClipCaptionModel(8, clip_length=4, prefix_dim=512, num_layers=2, mapping_type=cb.MappingType.TransformerDecoder)
    ref = O.make_encdec_state_dict(seed=1, prefix_length=8, clip_length=4, prefix_size=512, num_layers=2)
    mapper = {k: tuple(v.shape) for k, v in d.state_dict().items() if k.startswith("clip_project.")}
    assert mapper == {k: tuple(v.shape) for k, v in ref.items() if k.startswith("clip_project.")}
    with pytest.raises(Exception):
        cb.ClipCaptionModel(10, mapping_type="no_such_mapper")


def test_cpu_model_refuses_to_run():
    import capdec_b200 as cb
    m = cb.ClipCaptionModel(10, prefix_size=512)
    with pytest.raises(cb._lib.CapdecError):
        m(torch.zeros(1, 40, dtype=torch.int64), torch.zeros(1, 512), None)
    with pytest.raises(cb._lib.CapdecError):
        cb.noise_injection(torch.zeros(2, 512), 0.016)
