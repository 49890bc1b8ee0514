"""Pin the oracle against the REAL reference and write the golden fixtures — runs only where /root/reference exists.

    python oracle/pin_against_reference.py            # writes tests/golden/*.json

What runs: the reference's own classes imported from /root/reference/train.py (ClipCaptionModel,
ClipCaptionPrefix, noise_injection, MappingType) on top of the installed transformers' GPT2LMHeadModel, with the
three shims of SURVEY §8c (AdamW symbol, from_pretrained -> GPT2LMHeadModel(GPT2Config()), device -> cpu).
Weights/inputs come from oracle.capdec_oracle.make_state_dict / make_batch (seeded), are loaded into the
reference model with load_state_dict(strict=True), and the reference executes train.py:348-351 verbatim.
Recorded per case: loss, 96 sampled logits, per-parameter gradient norms and 8 sampled gradient entries each.
The same quantities from the oracle restatement are asserted equal (fp32 tolerance) before anything is written.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import torch
import torch.nn.functional as nnf

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import capdec_oracle as O  # noqa: E402

REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"


def import_reference(pdrop: float = 0.0):
    import transformers
    from transformers import GPT2Config, GPT2LMHeadModel

    orig = GPT2LMHeadModel.from_pretrained

    def fake_from_pretrained(name, *a, **k):  # shim 2 (train.py:266)
        cfg = GPT2Config(resid_pdrop=pdrop, embd_pdrop=pdrop, attn_pdrop=pdrop)
        cfg._attn_implementation = "eager"
        return GPT2LMHeadModel(cfg)

    GPT2LMHeadModel.from_pretrained = staticmethod(fake_from_pretrained)
    sys.path.insert(0, str(REF))
    sys.modules.pop("train", None)
    from transformers import GPT2Tokenizer, get_linear_schedule_with_warmup  # noqa: F401  (resolve lazies first)
    sys.modules["transformers"].AdamW = O.HFAdamW  # shim 1 (train.py:6) — must be set right before the import
    import train as ref_train  # noqa

    ref_train.device = torch.device("cpu")  # shim 3 (train.py:15,36)
    return ref_train  # from_pretrained stays patched: ClipCaptionModel.__init__ calls it (train.py:266)


def sample_idx(n: int, k: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (min(k, n),), generator=g)


def summarise(loss, logits, grads, seed=99):
    rec = {"loss": float(loss), "logits_shape": list(logits.shape)}
    li = sample_idx(logits.numel(), 96, seed)
    rec["logits_idx"] = li.tolist()
    rec["logits_val"] = logits.flatten()[li].double().tolist()
    rec["logits_absmax"] = float(logits.abs().max())
    g = {}
    for k in sorted(grads):
        t = grads[k]
        gi = sample_idx(t.numel(), 8, seed + 1)
        g[k] = {"norm": float(t.double().norm()), "idx": gi.tolist(), "val": t.flatten()[gi].double().tolist()}
    rec["grads"] = g
    return rec


CASES = {
    # name: (mapping_type, only_prefix, B, P, C, D, num_layers, full_length)
    "mlp_full_b4": ("mlp", False, 4, 10, 10, 512, 8, False),
    "mlp_prefix_only_b4": ("mlp", True, 4, 10, 10, 512, 8, False),
    "transformer_full_b2": ("transformer", False, 2, 40, 40, 512, 8, False),
    "mlp_full_d640_b3": ("mlp", False, 3, 10, 10, 640, 8, True),
    # flag combinations beyond the BASELINE configs: --only_prefix with the transformer mapper, prefix_length != clip_length
    "transformer_prefix_only_b2": ("transformer", True, 2, 12, 6, 640, 2, False),
    "transformer_short_clip_b3": ("transformer", False, 3, 8, 20, 512, 3, True),
}


def run_case(ref_train, name, cfg):
    mtype, only_prefix, B, P, C, D, nl, full_len = cfg
    sd = O.make_state_dict(seed=1, mapping_type=mtype, prefix_length=P, clip_length=C, prefix_size=D, num_layers=nl)
    tokens, prefix, noise = O.make_batch(seed=2, B=B, L=40, prefix_size=D, full_length=full_len)
    mask = O.make_mask(tokens, P)
    var = 0.016
    MT = ref_train.MappingType.MLP if mtype == "mlp" else ref_train.MappingType.Transformer
    cls = ref_train.ClipCaptionPrefix if only_prefix else ref_train.ClipCaptionModel
    torch.manual_seed(0)
    model = cls(P, clip_length=C, prefix_size=D, num_layers=nl, mapping_type=MT)
    missing = model.load_state_dict(sd, strict=True)
    model.train()  # ClipCaptionPrefix keeps gpt in eval (train.py:281-284); full model has pdrop = 0 via the shim
    # ---- the reference's own step, train.py:345-351 (noise draw made explicit through the global torch RNG) ----
    model.zero_grad()
    torch.manual_seed(1234)
    pfx = ref_train.noise_injection(prefix, var)
    torch.manual_seed(1234)
    g_draw = torch.randn(prefix.shape)
    pfx_oracle = O.noise_injection(prefix, var, noise=g_draw)
    assert torch.equal(pfx, pfx_oracle), "noise_injection restatement differs"
    outputs = model(tokens, pfx, mask)
    logits = outputs.logits[:, P - 1: -1]
    loss = nnf.cross_entropy(logits.reshape(-1, logits.shape[-1]), tokens.flatten(), ignore_index=0)
    loss.backward()
    ref_grads = {}
    for k, p in model.named_parameters() if not only_prefix else super(cls, model).named_parameters():
        if p.grad is not None and (not only_prefix or k.startswith("clip_project")):
            ref_grads[k] = p.grad.detach()
    # mask has no effect on the consumed logits / loss (SURVEY §8c probe) — re-verify and record
    with torch.no_grad():
        lg_nomask = model(tokens, pfx, None).logits[:, P - 1: -1]
        loss_nomask = nnf.cross_entropy(lg_nomask.reshape(-1, lg_nomask.shape[-1]), tokens.flatten(), ignore_index=0)
    # ---- oracle restatement on the same inputs ----
    trainable = (lambda k: k.startswith("clip_project")) if only_prefix else None
    o_loss, o_logits, o_grads = O.loss_and_grads(sd, tokens, pfx, mask, P, C, trainable)
    full_logits = outputs.logits.detach()
    assert abs(float(o_loss) - float(loss)) < 2e-6 * abs(float(loss)), (float(o_loss), float(loss))
    valid = torch.cat((torch.ones(B, P, dtype=torch.bool), tokens > 0), dim=1)  # compare logits at unpadded positions
    d_lg = (o_logits - full_logits)[valid].abs().max().item()
    assert d_lg < 1e-4 * max(1.0, full_logits.abs().max().item()), d_lg
    assert set(o_grads) == set(ref_grads), (sorted(set(o_grads) ^ set(ref_grads)))
    worst = 0.0
    for k in ref_grads:
        rel = (o_grads[k] - ref_grads[k]).norm().item() / max(ref_grads[k].norm().item(), 1e-12)
        worst = max(worst, rel)
        assert rel < 2e-4, (k, rel)
    rec = summarise(loss.detach(), full_logits, ref_grads)
    rec.update({"case": name, "config": dict(mapping_type=mtype, only_prefix=only_prefix, B=B, P=P, C=C, D=D,
                                             num_layers=nl, full_length=full_len, noise_variance=var,
                                             sd_seed=1, batch_seed=2, noise_torch_seed=1234),
                "loss_without_mask": float(loss_nomask), "noised_prefix_row0": pfx[0, :8].double().tolist(),
                "oracle_vs_reference": {"loss_abs": abs(float(o_loss) - float(loss)), "logits_maxabs_valid": d_lg,
                                        "worst_grad_rel_l2": worst},
                "versions": {"torch": torch.__version__, "transformers": __import__("transformers").__version__,
                             "reference_commit": "4451bfd"}})
    return rec


def main():
    assert REF.exists(), "the reference checkout is only mounted in the build container"
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    ref_train = import_reference(pdrop=0.0)
    for name, cfg in CASES.items():
        rec = run_case(ref_train, name, cfg)
        (GOLD / f"{name}.json").write_text(json.dumps(rec, indent=1))
        print(name, "loss", rec["loss"], "oracle-vs-ref", rec["oracle_vs_reference"])
    # noise_injection variants (train.py:27-39) incl. modality offset and dont_norm, seeded through torch's global RNG
    tokens, prefix, _ = O.make_batch(seed=5, B=6, prefix_size=640)
    x = prefix * 3.0
    off = torch.randn(1, 640, generator=torch.Generator().manual_seed(8)) * 0.05
    out = {}
    for tag, kw in {"plain": {}, "offset": {"modality_offset": off}, "dont_norm": {"dont_norm": True},
                    "zero_var": {"variance": 0.0}}.items():
        var = kw.pop("variance", 0.016)
        torch.manual_seed(77)
        y = ref_train.noise_injection(x, var, **kw)
        torch.manual_seed(77)
        yo = O.noise_injection(x, var, noise=torch.randn(x.shape), **kw)
        assert torch.equal(y, yo), tag
        out[tag] = {"row0": y[0, :6].double().tolist(), "norm0": float(y[0].norm()), "sum": float(y.double().sum())}
    (GOLD / "noise_injection.json").write_text(json.dumps({"seed": 77, "batch_seed": 5, "offset_seed": 8, "cases": out}, indent=1))
    print("noise ok")


class _FakeTokenizer:
    """No GPT-2 vocab files offline: ids are the comparable part. encode('.') -> [13] as for the real tokenizer."""

    def encode(self, text):
        if text == ".":
            return [13]
        return [int(text)] if text.isdigit() else [ord(ch) % 50257 for ch in text]

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)


def pin_generate_beam():
    """gpt2_prefix_eval.generate_beam (the reference's own function) vs the oracle restatement -> tests/golden/beam.json"""
    for name in ("clip", "pycocotools", "pycocotools.coco", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    import transformers
    sys.modules["transformers"].AdamW = O.HFAdamW
    import gpt2_prefix_eval as ev  # noqa
    import gpt2_prefix as gp  # noqa
    P, D = 10, 640
    sd = O.make_state_dict(seed=1, mapping_type="mlp", prefix_length=P, prefix_size=D, weight_std=0.08)  # sharper logits
    torch.manual_seed(0)
    model = gp.ClipCaptionModel(P, prefix_dim=D, mapping_type=gp.MappingType.MLP)
    model.gpt.config._attn_implementation = "eager"
    model.load_state_dict(sd, strict=True)
    model.eval()
    recs = []

    def run(case, stop, temperature):
        _, prefix, _ = O.make_batch(seed=20 + case, B=1, prefix_size=D)
        with torch.no_grad():
            embed = model.clip_project(prefix).reshape(1, P, -1)
            texts = ev.generate_beam(model, _FakeTokenizer(), embed=embed, entry_length=12, temperature=temperature,
                                     stop_token="." if stop == 13 else str(stop))
        ref_ids = [[int(t) for t in txt.split()] for txt in texts]
        ids, scores, lens = O.generate_beam(sd, O.mlp_mapper(sd, prefix).view(1, P, -1), entry_length=12,
                                            temperature=temperature, stop_token_index=stop)
        assert ids == ref_ids, (ids, ref_ids)
        recs.append({"batch_seed": 20 + case, "stop_token_index": stop, "temperature": temperature, "ids": ref_ids,
                     "scores": scores, "seq_lengths": lens})
        return ref_ids

    for case in range(3):
        ids = run(case, 13, 1.0)
        # early stops: make tokens that the beams actually emit the stop token (exercises :90-91, :106-108)
        run(case, ids[-1][1], 0.7)
        sharp = run(case, 13, 0.05)      # peaked distribution: the best beam's log-probs are ~0 ...
        run(case, sharp[0][2], 0.05)     # ... so when its 3rd token is the stop token it stays in the beam, stopped
        run(case, sharp[0][0], 0.05)
    (GOLD / "beam.json").write_text(json.dumps({"config": dict(P=P, D=D, sd_seed=1, weight_std=0.08, entry_length=12,
                                                                 beam_size=5), "cases": recs}, indent=1))
    print("generate_beam ok", [r["ids"][0][:6] for r in recs])


def pin_generate_beam_prompt():
    """The PROMPT path of gpt2_prefix_eval.generate_beam (:65-68, :82-88, :111-112: the prompt ids stay in `tokens` and each
    beam is cut to `seq_length` ids, a count of generated tokens only) -> tests/golden/beam_prompt.json (decoded texts)."""
    for name in ("clip", "pycocotools", "pycocotools.coco", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["transformers"].AdamW = O.HFAdamW
    import gpt2_prefix_eval as ev  # noqa
    import gpt2_prefix as gp  # noqa
    P, D = 10, 640
    sd = O.make_state_dict(seed=1, mapping_type="mlp", prefix_length=P, prefix_size=D, weight_std=0.08)
    torch.manual_seed(0)
    model = gp.ClipCaptionModel(P, prefix_dim=D, mapping_type=gp.MappingType.MLP)
    model.gpt.config._attn_implementation = "eager"
    model.load_state_dict(sd, strict=True)
    model.eval()
    recs = []
    for prompt, temperature in (("capdec", 1.0), ("a b", 0.7), ("x", 1.0)):       # _FakeTokenizer: one id per character
        with torch.no_grad():
            texts = ev.generate_beam(model, _FakeTokenizer(), prompt=prompt, entry_length=12, temperature=temperature)
        recs.append({"prompt": prompt, "temperature": temperature, "texts": texts})
    (GOLD / "beam_prompt.json").write_text(json.dumps({"config": dict(P=P, D=D, sd_seed=1, weight_std=0.08, entry_length=12,
                                                                        beam_size=5), "cases": recs}, indent=1))
    print("generate_beam (prompt) ok", [r["texts"][0] for r in recs])


def pin_modality_bridger():
    """others/supervised_embedding_bridger.MLP (the reference's own class; `wandb` stubbed) vs the oracle restatement, on
    seeded weights and on the reference's trained weights file -> tests/golden/bridger.json"""
    import importlib.util
    sys.modules.setdefault("wandb", types.ModuleType("wandb"))
    spec = importlib.util.spec_from_file_location("_ref_bridger", str(REF / "others" / "supervised_embedding_bridger.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    recs = []
    g = torch.Generator().manual_seed(77)
    x = torch.randn(5, 640, generator=g)
    x = x / x.norm(2, -1, keepdim=True)
    for name, sd in (("seeded", O.make_bridger_state_dict(seed=9)),
                     ("trained", torch.load(REF / "others" / "weights_modality_mapper.pt", map_location="cpu"))):
        ref = mod.MLP(640, 640, 640, 8)
        ref.load_state_dict(sd, strict=True)
        ref.eval()
        with torch.no_grad():
            y_ref = ref(x)
        y = O.modality_bridger(sd, x)
        assert torch.equal(y, y_ref), (name, (y - y_ref).abs().max())
        if name == "seeded":
            recs.append({"weights": "make_bridger_state_dict(seed=9)", "x_seed": 77, "y_row0": y_ref[0, :16].double().tolist(),
                         "y_norms": y_ref.norm(2, -1).double().tolist(), "y_sum": float(y_ref.double().sum())})
    (GOLD / "bridger.json").write_text(json.dumps({"cases": recs}, indent=1))
    print("modality bridger ok (oracle == reference MLP bit for bit on seeded and on the trained weights)")


def pin_generate2():
    """gpt2_prefix_eval.generate2 (the reference's own function: greedy decode behind a top-p mask) vs the oracle
    restatement -> tests/golden/greedy.json"""
    for name in ("clip", "pycocotools", "pycocotools.coco", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["transformers"].AdamW = O.HFAdamW
    import gpt2_prefix_eval as ev  # noqa
    import gpt2_prefix as gp  # noqa
    P, D = 10, 640
    sd = O.make_state_dict(seed=1, mapping_type="mlp", prefix_length=P, prefix_size=D, weight_std=0.08)
    torch.manual_seed(0)
    model = gp.ClipCaptionModel(P, prefix_dim=D, mapping_type=gp.MappingType.MLP)
    model.gpt.config._attn_implementation = "eager"
    model.load_state_dict(sd, strict=True)
    model.eval()
    recs = []

    def run(case, stop, temperature, entry_length=12):
        _, prefix, _ = O.make_batch(seed=20 + case, B=1, prefix_size=D)
        with torch.no_grad():
            embed = model.clip_project(prefix).reshape(1, P, -1)
            text = ev.generate2(model, _FakeTokenizer(), embed=embed, entry_length=entry_length, temperature=temperature,
                                stop_token="." if stop == 13 else str(stop))
            ref_ids = [int(t) for t in text.split()]
            ids = O.generate2(sd, O.mlp_mapper(sd, prefix).view(1, P, -1), entry_length=entry_length,
                              temperature=temperature, stop_token_index=stop)
        assert ids == ref_ids, (ids, ref_ids)
        recs.append({"batch_seed": 20 + case, "stop_token_index": stop, "temperature": temperature,
                     "entry_length": entry_length, "ids": ref_ids})
        return ref_ids

    for case in range(3):
        ids = run(case, 13, 1.0)
        run(case, ids[3], 1.0)           # the 4th greedy token as stop token: early exit, stop token included (:181-187)
        run(case, ids[1], 0.5)
        run(case, 13, 0.05, entry_length=7)
    (GOLD / "greedy.json").write_text(json.dumps({"config": dict(P=P, D=D, sd_seed=1, weight_std=0.08), "cases": recs}, indent=1))
    print("generate2 ok", [r["ids"][:6] for r in recs])


def pin_dataset():
    """ClipCocoDataset.__getitem__ + the default collate (the reference's own class, constructed without its file /
    tokenizer loading) vs the oracle restatement -> tests/golden/datafeed.json."""
    ref_train = import_reference(pdrop=0.0)
    from torch.utils.data import DataLoader
    recs = []
    for case, (half, norm, P) in enumerate([(False, True, 10), (True, True, 10), (False, False, 40)]):
        caps, cap2emb, table = O.make_caption_table(seed=40 + case, n=24, n_emb=9, prefix_size=512, half=half)
        ds = ref_train.ClipCocoDataset.__new__(ref_train.ClipCocoDataset)   # skip __init__: no pickle / tokenizer offline
        ds.captions_tokens = [c.clone() for c in caps]
        ds.caption2embedding = list(cap2emb)
        ds.prefixes = table
        ds.prefix_length = P
        ds.normalize_prefix = norm
        all_len = torch.tensor([len(c) for c in caps]).float()
        ds.max_seq_len = min(int(all_len.mean() + all_len.std() * 10), int(all_len.max())) - (7 if case == 2 else 0)
        L = ds.max_seq_len                                                  # case 2 forces truncation (train.py:56-58)
        assert case == 2 or L == O.dataset_max_seq_len(caps)
        idx = [3, 0, 23, 11, 7, 19]
        batch = next(iter(DataLoader(torch.utils.data.Subset(ds, idx), batch_size=len(idx), shuffle=False)))
        tokens, mask, prefix = batch
        for j, it in enumerate(idx):
            t, m, pf = O.dataset_item(caps, cap2emb, table, it, L, P, norm)
            assert torch.equal(t, tokens[j]) and torch.equal(m, mask[j]) and torch.equal(pf, prefix[j]), (case, it)
        # reference quirk (train.py:55-61): pad_tokens caches the padded tensor and then zeroes its padding IN PLACE, so a
        # second visit of the same item sees no negative ids and returns mask == 1 everywhere (tokens unchanged).  The
        # mask cannot change a consumed logit, so the restatement keeps the first-visit semantics.
        t2, m2, _ = ds[idx[0]]
        assert torch.equal(t2, tokens[0]) and bool((m2 == 1).all())
        recs.append({"table_seed": 40 + case, "n": 24, "n_emb": 9, "half": half, "normalize_prefix": norm,
                     "prefix_length": P, "max_seq_len": L, "idx": idx, "tokens": tokens.tolist(),
                     "mask_sum_rows": mask.sum(1).tolist(), "prefix_dtype": str(prefix.dtype).replace("torch.", ""),
                     "prefix_row_norms": prefix.float().norm(2, -1).double().tolist(),
                     "prefix_head": prefix[:, :6].double().tolist(), "prefix_sum": float(prefix.double().sum())})
    (GOLD / "datafeed.json").write_text(json.dumps({"cases": recs}, indent=1))
    print("dataset ok", [r["max_seq_len"] for r in recs])


def pin_encdec_mapper():
    """transformer_mapper.TransformerEncoderDecoder (the reference's own module) vs oracle.encdec_mapper
    -> tests/golden/encdec_mapper.json."""
    sys.path.insert(0, str(REF))
    import transformer_mapper as tm  # noqa
    recs = []
    for case, (P, C, D, nl, B) in enumerate([(10, 10, 512, 2, 3), (40, 40, 640, 1, 2), (4, 7, 512, 3, 1)]):
        sd = O.make_encdec_state_dict(seed=60 + case, prefix_length=P, clip_length=C, prefix_size=D, num_layers=nl)
        torch.manual_seed(0)
        net = tm.TransformerEncoderDecoder(D, 768, P, C, nl)
        net.load_state_dict({k[len("clip_project."):]: v for k, v in sd.items()}, strict=True)
        x = torch.randn(B, D, generator=torch.Generator().manual_seed(70 + case))
        x = x / x.norm(2, -1, keepdim=True)
        with torch.no_grad():
            ref = net(x)
        out = O.encdec_mapper(sd, x, C)
        assert ref.shape == out.shape == (B, P, 768)
        assert (ref - out).abs().max() <= 1e-5 * ref.abs().max(), float((ref - out).abs().max())
        idx = sample_idx(ref.numel(), 64, 5)
        recs.append({"sd_seed": 60 + case, "x_seed": 70 + case, "P": P, "C": C, "D": D, "num_layers": nl, "B": B,
                     "idx": idx.tolist(), "val": ref.flatten()[idx].double().tolist(), "absmax": float(ref.abs().max()),
                     "norm": float(ref.double().norm()), "oracle_maxabs_diff": float((ref - out).abs().max())})
    (GOLD / "encdec_mapper.json").write_text(json.dumps({"cases": recs}, indent=1))
    print("encdec mapper ok", [r["oracle_maxabs_diff"] for r in recs])


if __name__ == "__main__":
    main()
    pin_generate_beam()
    pin_generate_beam_prompt()
    pin_modality_bridger()
    pin_generate2()
    pin_dataset()
    pin_encdec_mapper()
