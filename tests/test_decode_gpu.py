"""Beam-search decode path (GPU): kernels vs torch restatements, and the whole KV-cached search vs tests/golden/beam.json
(ids produced by the reference's own generate_beam, gpt2_prefix_eval.py:50-115, through oracle/pin_against_reference.py).

Tolerances: token ids exact in fp32 mode (CUDA-core GEMMs); scores abs 2e-4 (fp32) — the candidates' averaged log-probs
are separated by >= 1e-3 in the golden cases.  tf32 mode: beam search amplifies 1xTF32 logit noise whenever the 5th/6th candidates are
close (a different pruning decision leads to a different final beam), so only the agreement rate is asserted; best-beam ids must agree on at least half of the cases.
"""
import json
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import capdec_oracle as O  # noqa: E402  (checker only)

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "beam.json").read_text())


def _state(n_img, beam, Tmax, max_sel, dev="cuda"):
    R = n_img * beam
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
    return SimpleNamespace(step=z(1, dt=torch.int32), ticket=z(1, dt=torch.int32), scores=z(R), seq_len=z(R),
                           stopped=z(R, dt=torch.int32), src=z(2, R, Tmax, dt=torch.int32),
                           hist_tok=z(max_sel, R, dt=torch.int32), hist_parent=z(max_sel, R, dt=torch.int32),
                           img_done=z(n_img, dt=torch.int32), cand_val=z(R, 8), cand_idx=z(R, 8, dt=torch.int32),
                           row_lse=z(R))


@pytest.mark.parametrize("V,temp,k", [(50257, 1.0, 5), (50257, 0.7, 8), (1000, 0.05, 1), (33, 1.0, 5)])
def test_row_topk_matches_torch(V, temp, k):
    from capdec_b200 import ops
    torch.manual_seed(V + k)
    rows, ld = 7, (V + 127) // 128 * 128
    buf = torch.randn(rows, ld, device="cuda") * 3
    x = buf[:, :V]
    val, idx, lse = torch.zeros(rows, 8, device="cuda"), torch.zeros(rows, 8, device="cuda", dtype=torch.int32), torch.zeros(rows, device="cuda")
    ops.row_topk(x, V, temp, k, val, idx, lse)
    xs = x.double() / temp
    rv, ri = xs.topk(k, -1)
    assert torch.equal(idx[:, :k].long(), ri)
    assert torch.allclose(val[:, :k].double(), rv, atol=1e-5)
    assert torch.allclose(lse.double(), xs.logsumexp(-1), atol=2e-5)


def test_decode_attention_matches_dense_over_lineage():
    from capdec_b200 import ops
    torch.manual_seed(3)
    n_img, beam, P, Tmax, H, hd = 2, 5, 4, 40, 12, 64
    R, d = n_img * beam, H * hd
    c = 9                      # selections done -> this token sits at position pos = P + c - 1
    pos = P + c - 1
    st = _state(n_img, beam, Tmax, 16)
    st.step.fill_(c)
    kc, vc = torch.randn(R, Tmax, d, device="cuda"), torch.randn(R, Tmax, d, device="cuda")
    src = torch.stack([torch.randint(0, beam, (R, Tmax)) + (torch.arange(R) // beam * beam)[:, None] for _ in range(2)])
    st.src.copy_(src.to(torch.int32))
    qkv = torch.randn(R, 3 * d, device="cuda")
    ctx = torch.zeros(R, d, device="cuda")
    kc0, vc0 = kc.clone(), vc.clone()
    ops.decode_attention(qkv, kc, vc, st, ctx, H, hd, P, Tmax, hd ** -0.5)
    torch.cuda.synchronize()
    # the token's own K/V were appended at (row, pos); nothing else in the cache moved
    assert torch.equal(kc[:, pos], qkv[:, d:2 * d]) and torch.equal(vc[:, pos], qkv[:, 2 * d:])
    kc0[:, pos], vc0[:, pos] = kc[:, pos], vc[:, pos]
    assert torch.equal(kc, kc0) and torch.equal(vc, vc0)
    table = src[c & 1].cuda()
    t_idx = torch.arange(pos, device="cuda")
    for b in range(R):
        K = torch.cat([kc0[table[b, :pos], t_idx], qkv[b:b + 1, d:2 * d]]).double().view(pos + 1, H, hd)
        Vv = torch.cat([vc0[table[b, :pos], t_idx], qkv[b:b + 1, 2 * d:]]).double().view(pos + 1, H, hd)
        q = qkv[b, :d].double().view(H, hd)
        w = torch.einsum("hd,thd->ht", q, K) * hd ** -0.5
        ref = torch.einsum("ht,thd->hd", w.softmax(-1), Vv).reshape(-1)
        assert torch.allclose(ctx[b].double(), ref, atol=2e-5), (b, (ctx[b].double() - ref).abs().max())


def _ref_select(logp, scores, seq_len, stopped, first, beam, stop):
    """gpt2_prefix_eval.py:80-106 for one image on log-prob rows (torch, same dtype/ops as the reference)."""
    if first:
        sc, nt = logp[:1].topk(beam, -1)
        return sc.squeeze(0), nt.squeeze(0), torch.zeros(beam, dtype=torch.long), seq_len.clone(), nt.squeeze(0).eq(stop)
    logp = logp.clone()
    logp[stopped] = -float("inf")
    logp[stopped, 0] = 0
    ssum = scores[:, None] + logp
    seq_len = seq_len.clone()
    seq_len[~stopped] += 1
    avg = ssum / seq_len[:, None]
    avg, nt = avg.view(-1).topk(beam, -1)
    parent = nt // ssum.shape[1]
    seq_len = seq_len[parent]
    nt = nt % ssum.shape[1]
    return avg * seq_len, nt, parent, seq_len, stopped[parent] | nt.eq(stop)


def test_row_topk_plus_beam_select_follow_the_reference_recurrence():
    from capdec_b200 import ops
    torch.manual_seed(11)
    n_img, beam, P, Tmax, V, stop, n_steps = 3, 5, 2, 24, 61, 7, 14
    R = n_img * beam
    st = _state(n_img, beam, Tmax, n_steps)
    ops.beam_init(st, n_img, beam, P, Tmax)
    ref = [dict(scores=None, seq_len=torch.ones(beam), stopped=torch.zeros(beam, dtype=torch.bool)) for _ in range(n_img)]
    lineage = [[[] for _ in range(beam)] for _ in range(n_img)]  # per beam: list of physical rows per decode position
    for c in range(n_steps):
        rows = n_img if c == 0 else R
        logits = (torch.randn(rows, V) * 4).cuda()
        ops.row_topk(logits, V, 1.0, beam, st.cand_val, st.cand_idx, st.row_lse)
        ops.beam_select(st, n_img, beam, P, Tmax, V, stop)
        torch.cuda.synchronize()
        assert int(st.step.item()) == c + 1
        logp = logits.cpu().log_softmax(-1)
        for i in range(n_img):
            r = ref[i]
            lp = logp[i:i + 1] if c == 0 else logp[i * beam:(i + 1) * beam]
            sc, nt, parent, sl, stp = _ref_select(lp, r["scores"], r["seq_len"], r["stopped"], c == 0, beam, stop)
            g = slice(i * beam, (i + 1) * beam)
            assert torch.equal(st.hist_tok[c, g].cpu().long(), nt), (c, i)
            assert torch.equal(st.hist_parent[c, g].cpu().long(), parent), (c, i)
            assert torch.allclose(st.scores[g].cpu(), sc, atol=1e-4)
            assert torch.equal(st.seq_len[g].cpu(), sl)
            assert torch.equal(st.stopped[g].cpu().bool(), stp)
            assert int(st.img_done[i].item()) == int(stp.all())
            r.update(scores=sc, seq_len=sl, stopped=stp)
            # lineage table: prefix -> row i*beam; decode position j -> the physical row that wrote it
            lineage[i] = [lineage[i][int(p)] + [i * beam + int(p)] for p in parent] if c > 0 else [[] for _ in range(beam)]
            tab = st.src[(c + 1) & 1, g].cpu()
            for b in range(beam):
                assert tab[b, :P].tolist() == [i * beam] * P
                assert tab[b, P:P + c].tolist() == lineage[i][b], (c, i, b)
                assert int(tab[b, P + c]) == i * beam + b
    assert any(bool(r["stopped"].any()) for r in ref)  # the stop logic was actually exercised


def _beam_model(mode):
    import capdec_b200 as cb
    c = GOLD["config"]
    cb.ops.set_precision(mode)
    sd = O.make_state_dict(seed=c["sd_seed"], mapping_type="mlp", prefix_length=c["P"], prefix_size=c["D"],
                           weight_std=c["weight_std"])
    model = cb.ClipCaptionModel(c["P"], prefix_size=c["D"], mapping_type=cb.MappingType.MLP)
    model.load_state_dict(sd)
    return model.to("cuda").eval(), c


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
@pytest.mark.parametrize("graph", [False, True])
def test_generate_beam_matches_reference_goldens(graph, mode):
    """Token ids IDENTICAL to the reference's generate_beam (gpt2_prefix_eval.py:50-115) on all pinned cases, in the
    CUDA-core fp32 mode and in the tensor-core fp32-grade mode (3xTF32 tcgen05 GEMMs)."""
    import capdec_b200 as cb
    model, c = _beam_model(mode)
    try:
        for case in GOLD["cases"]:
            _, prefix, _ = O.make_batch(seed=case["batch_seed"], B=1, prefix_size=c["D"])
            embed = model.clip_project(prefix.cuda()).view(1, c["P"], -1)
            (ids, scores, lens), = cb.generate_beam_ids(model, embed, c["beam_size"], c["entry_length"], case["temperature"],
                                                        case["stop_token_index"], use_cuda_graph=graph)
            assert ids == case["ids"], (case["stop_token_index"], case["temperature"])
            assert lens == case["seq_lengths"]
            # scores: fp32 noise x 1/temperature; the 3xTF32 logits are within 5e-5 rel-L2 of the reference's (measured 4.7e-4 on the
            # temperature-0.05 cases); the token ids above are identical in both modes
            assert max(abs(a - b) for a, b in zip(scores, case["scores"])) < (2e-4 if mode == "fp32" else 2e-3)
    finally:
        cb.ops.set_precision("tf32")


def test_generate_beam_batched_images_equal_one_at_a_time_and_api_mirror():
    import capdec_b200 as cb
    model, c = _beam_model("fp32")
    try:
        cases = [k for k in GOLD["cases"] if k["temperature"] == 1.0 and k["stop_token_index"] == 13]
        prefixes = torch.cat([O.make_batch(seed=k["batch_seed"], B=1, prefix_size=c["D"])[1] for k in cases]).cuda()
        embed = model.clip_project(prefixes).view(len(cases), c["P"], -1)
        out = cb.generate_beam_ids(model, embed, c["beam_size"], c["entry_length"], 1.0, 13)
        for (ids, _, _), k in zip(out, cases):
            assert ids == k["ids"]

        class Tok:  # ids are the comparable part offline (no GPT-2 vocab files): '.' -> 13 like the real tokenizer
            def encode(self, text):
                return [13] if text == "." else [int(text)]

            def decode(self, ids):
                return " ".join(str(i) for i in ids)

        texts = cb.generate_beam(model, Tok(), embed=embed[:1], entry_length=c["entry_length"])
        assert texts == [" ".join(str(i) for i in ids) for ids in cases[0]["ids"]]
        # the many-images form of the same call (make_preds decodes one image per call, predictions_runner.py:213-233)
        many = cb.generate_beam_batch(model, Tok(), embed, beam_size=c["beam_size"], entry_length=c["entry_length"])
        assert many == [[" ".join(str(i) for i in ids) for ids in k["ids"]] for k in cases]
    finally:
        cb.ops.set_precision("tf32")


def test_generate_beam_tf32_best_beam_score_close():
    import capdec_b200 as cb
    model, c = _beam_model("tf32")
    agree, worst = 0, 0.0
    for case in GOLD["cases"]:
        _, prefix, _ = O.make_batch(seed=case["batch_seed"], B=1, prefix_size=c["D"])
        embed = model.clip_project(prefix.cuda()).view(1, c["P"], -1)
        (ids, scores, _), = cb.generate_beam_ids(model, embed, c["beam_size"], c["entry_length"], case["temperature"],
                                                 case["stop_token_index"])
        diff = abs(scores[0] - case["scores"][0])
        print(f"tf32 beam: T={case['temperature']} stop={case['stop_token_index']} best-score diff {diff:.4f} "
              f"ids equal {ids[0] == case['ids'][0]}")
        if case["temperature"] >= 0.7:  # the 0.05 cases amplify the 1xTF32 logit error 20x: reported via `agree` only
            worst = max(worst, diff)
        agree += ids[0] == case["ids"][0]
    print(f"tf32 best-beam id agreement: {agree}/{len(GOLD['cases'])}, worst score diff at T>=0.7: {worst:.4f}")
    assert agree >= len(GOLD["cases"]) // 2


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_generate2_greedy_matches_reference_goldens(mode):
    """gpt2_prefix_eval.generate2 (:118-198) = greedy decoding; ids pinned on the reference's own function
    (tests/golden/greedy.json, oracle/pin_against_reference.py::pin_generate2); identical in the CUDA-core fp32 mode and in
    the tensor-core fp32-grade mode."""
    import json
    import capdec_b200 as cb
    rec = json.loads((Path(__file__).resolve().parent / "golden" / "greedy.json").read_text())
    model, c = _beam_model(mode)
    try:
        class Tok:
            def __init__(self, stop):
                self.stop = stop

            def encode(self, text):
                return [self.stop] if text == "." else [int(text)]

            def decode(self, ids):
                return " ".join(str(int(i)) for i in ids)

        for case in rec["cases"]:
            _, prefix, _ = O.make_batch(seed=case["batch_seed"], B=1, prefix_size=c["D"])
            embed = model.clip_project(prefix.cuda()).view(1, c["P"], -1)
            ids, = cb.generate_greedy_ids(model, embed, case["entry_length"], case["temperature"], case["stop_token_index"])
            assert ids == case["ids"], case
            text = cb.generate2(model, Tok(case["stop_token_index"]), embed=embed, entry_length=case["entry_length"],
                                temperature=case["temperature"])
            assert text == " ".join(str(i) for i in case["ids"])
        # many images in one call
        cases = [k for k in rec["cases"] if k["stop_token_index"] == 13 and k["temperature"] == 1.0]
        prefixes = torch.cat([O.make_batch(seed=k["batch_seed"], B=1, prefix_size=c["D"])[1] for k in cases]).cuda()
        embed = model.clip_project(prefixes).view(len(cases), c["P"], -1)
        out = cb.generate_greedy_ids(model, embed, cases[0]["entry_length"], 1.0, 13)
        assert out == [k["ids"] for k in cases]
    finally:
        cb.ops.set_precision("tf32")


def test_generate_beam_prompt_path_matches_reference_goldens():
    """`generate_beam(prompt=...)` (gpt2_prefix_eval.py:65-68): the prompt ids stay in front and every beam is cut to its
    generated-token count, exactly as the reference's own function returns it (tests/golden/beam_prompt.json,
    oracle/pin_against_reference.py::pin_generate_beam_prompt)."""
    import json
    import capdec_b200 as cb
    rec = json.loads((Path(__file__).resolve().parent / "golden" / "beam_prompt.json").read_text())
    model, c = _beam_model("fp32")
    try:
        class Tok:      # the pin script's tokenizer: one id per character, '.' -> 13
            def encode(self, text):
                if text == ".":
                    return [13]
                return [int(text)] if text.isdigit() else [ord(ch) % 50257 for ch in text]

            def decode(self, ids):
                return " ".join(str(int(i)) for i in ids)

        for case in rec["cases"]:
            texts = cb.generate_beam(model, Tok(), prompt=case["prompt"], entry_length=rec["config"]["entry_length"],
                                     temperature=case["temperature"])
            assert texts == case["texts"], case["prompt"]
    finally:
        cb.ops.set_precision("tf32")
