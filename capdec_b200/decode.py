"""KV-cached, batched beam search over the hand-written kernels.

Mirror of the reference's `generate_beam` (gpt2_prefix_eval.py:50-115; the same function is repeated in
predictions_runner.py and gpt2_prefix_e2e.py).  The reference keeps `generated` = all embeddings so far and re-runs the
full GPT-2 forward over it for every new token; here the prefix is run once (prefill), K/V of every position stay in HBM,
and each new token costs one single-row pass per beam.  Beam re-ordering (`generated[next_tokens_source]`, :99) re-points a
lineage table instead of moving the cache.  `n_img` images are decoded together (the reference does one at a time); every
image follows exactly the reference's per-image recurrence.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import CapdecError, launch_count as _launch_count

_CHECK_EVERY = 4  # host polls the "all beams stopped" flags every few steps (gpt2_prefix_eval.py:107-108 early exit)


class BeamDecoder:
    """Decode state + buffers for (n_img, beam, P, entry_length) on one engine; reusable across calls."""

    def __init__(self, engine, n_img: int, beam: int, P: int, entry_length: int, use_cuda_graph: bool = True):
        eng = self.eng = engine
        if beam < 1 or beam > 8:
            raise CapdecError("beam_size must be in 1..8")
        self.n_img, self.beam, self.P, self.entry_length = n_img, beam, P, entry_length
        self.R = R = n_img * beam
        self.Tmax = Tmax = P + entry_length
        if Tmax > 128:
            raise CapdecError(f"prefix ({P}) + entry_length ({entry_length}) exceeds the 128-position decode cache")
        dev, d, F = eng.dev, eng.d, eng.F
        e = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        st = self.st = SimpleNamespace()
        st.step = e(1, dt=torch.int32)
        st.ticket = e(1, dt=torch.int32)
        st.scores, st.seq_len = e(R), e(R)
        st.stopped = e(R, dt=torch.int32)
        st.src = e(2, R, Tmax, dt=torch.int32)
        st.hist_tok = e(entry_length, R, dt=torch.int32)
        st.hist_parent = e(entry_length, R, dt=torch.int32)
        st.img_done = e(n_img, dt=torch.int32)
        st.cand_val = e(R, 8)
        st.cand_idx = e(R, 8, dt=torch.int32)
        st.row_lse = e(R)
        self.kc = [e(R, Tmax, d) for _ in range(eng.nl)]
        self.vc = [e(R, Tmax, d) for _ in range(eng.nl)]
        # single-position activations
        self.x, self.x1, self.y, self.h1, self.x2, self.ctx, self.xf = (e(R, d) for _ in range(7))
        self.h = [e(R, d), e(R, d)]
        self.st1 = e(R, 2)
        self.qkv = e(R, 3 * d)
        self.g = e(R, F)
        self.logits = e(R, eng.Vp)
        self.xlast = e(n_img, d)
        self.logits0 = e(n_img, eng.Vp)
        self.use_graph = use_cuda_graph
        self._graph = None
        self.step_launches = 0   # kernel launches inside one captured decode step
        self.replays = 0         # graph replays so far (bench.py: launches = eager C-ABI calls + replays * step_launches)

    # ---- one decode step: token of selection c-1 at position P+c-1 -> selection c -----------------------------------
    def _step(self, temperature: float, stop_token: int):
        eng, st, p = self.eng, self.st, self.eng.p
        d, eps = eng.d, eng.cfg.layer_norm_epsilon
        ops.decode_embed(st, p["gpt.transformer.wte.weight"], p["gpt.transformer.wpe.weight"], self.h[0], self.P)
        ops.add_ln_fwd(self.h[0], None, None, self.x1, self.st1, p["gpt.transformer.h.0.ln_1.weight"],
                       p["gpt.transformer.h.0.ln_1.bias"], eps=eps)
        hin, hout = self.h
        for l in range(eng.nl):
            pre = f"gpt.transformer.h.{l}."
            ops.linear_fwd(self.x1, p[pre + "attn.c_attn.weight"], "conv1d", p[pre + "attn.c_attn.bias"], self.qkv)
            ops.decode_attention(self.qkv, self.kc[l], self.vc[l], st, self.ctx, eng.H, eng.hd, self.P, self.Tmax,
                                 eng.hd ** -0.5)
            ops.linear_fwd(self.ctx, p[pre + "attn.c_proj.weight"], "conv1d", p[pre + "attn.c_proj.bias"], self.y)
            ops.add_ln_fwd(hin, self.y, self.h1, self.x2, self.st1, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"], eps=eps)
            ops.linear_fwd(self.x2, p[pre + "mlp.c_fc.weight"], "conv1d", p[pre + "mlp.c_fc.bias"], self.g,
                           act=ops.ACT_GELU_NEW)
            ops.linear_fwd(self.g, p[pre + "mlp.c_proj.weight"], "conv1d", p[pre + "mlp.c_proj.bias"], self.y)
            if l + 1 < eng.nl:
                nx = f"gpt.transformer.h.{l + 1}."
                ops.add_ln_fwd(self.h1, self.y, hout, self.x1, self.st1, p[nx + "ln_1.weight"], p[nx + "ln_1.bias"], eps=eps)
            else:
                ops.add_ln_fwd(self.h1, self.y, hout, self.xf, self.st1, p["gpt.transformer.ln_f.weight"],
                               p["gpt.transformer.ln_f.bias"], eps=eps)
            hin, hout = hout, hin
        lg = self.logits[:, : eng.V]
        ops.linear_fwd(self.xf, p["gpt.transformer.wte.weight"], "linear", None, lg)
        ops.row_topk(lg, eng.V, temperature, self.beam, st.cand_val, st.cand_idx, st.row_lse)
        ops.beam_select(st, self.n_img, self.beam, self.P, self.Tmax, eng.V, stop_token)

    def _prefill(self, embed: torch.Tensor, temperature: float, stop_token: int):
        """embed [n_img, P, d] (`model.clip_project(prefix).reshape(1, prefix_length, -1)`, gpt2_prefix_eval.py:271-272)."""
        eng, st = self.eng, self.st
        n_img, P, d = self.n_img, self.P, eng.d
        a = eng._arena(n_img, 0, P=P)
        a.pdrop = (0.0, 0.0, 0.0)  # decoding runs the model in eval mode (gpt2_prefix_eval.py:269 model.eval())
        ops.embed_fwd(None, embed, None, eng.p["gpt.transformer.wpe.weight"], a.h[0], n_img, P, 0)
        eng._trunk_fwd(a)
        ops.beam_init(st, n_img, self.beam, P, self.Tmax)
        for l in range(eng.nl):
            ops.kv_prefill(a.qkv[l], self.kc[l], self.vc[l], n_img, self.beam, P, self.Tmax, d)
        ops.rows_gather(a.xf, self.xlast, n_img, P, 1, P - 1)  # logits[:, -1, :] (:77)
        lg = self.logits0[:, : eng.V]
        ops.linear_fwd(self.xlast, eng.p["gpt.transformer.wte.weight"], "linear", None, lg)
        ops.row_topk(lg, eng.V, temperature, self.beam, st.cand_val, st.cand_idx, st.row_lse)
        ops.beam_select(st, n_img, self.beam, P, self.Tmax, eng.V, stop_token)

    @torch.no_grad()
    def run(self, embed: torch.Tensor, temperature: float = 1.0, stop_token: int = 13):
        embed = embed.detach().to(device=self.eng.dev, dtype=torch.float32).contiguous().view(self.n_img, self.P, self.eng.d)
        key = (float(temperature), int(stop_token))
        self._prefill(embed, *key)
        n_sel = 1
        for it in range(1, self.entry_length):
            if self.use_graph:
                if self._graph is None or self._graph[0] != key:
                    if it == 1:  # one eager step first (module loading / lazy init must not happen under capture)
                        self._step(*key)
                        n_sel += 1
                        continue
                    g = torch.cuda.CUDAGraph()
                    c0 = _launch_count()
                    with torch.cuda.graph(g):
                        self._step(*key)
                    self.step_launches = _launch_count() - c0   # kernels one replay of the decode-step graph launches
                    self._graph = (key, g)  # capture does not execute: fall through to the replay below
                self._graph[1].replay()
                self.replays += 1
            else:
                self._step(*key)
            n_sel += 1
            if it % _CHECK_EVERY == 0 and bool(self.st.img_done.all().item()):
                break
        return self._collect(n_sel)

    def _collect(self, n_sel: int):
        """Back-track the (token, parent) history into per-beam token lists ordered like gpt2_prefix_eval.py:110-114."""
        st, beam, n_img = self.st, self.beam, self.n_img
        tok = st.hist_tok[:n_sel].view(n_sel, n_img, beam).long()
        par = st.hist_parent[:n_sel].view(n_sel, n_img, beam).long()
        seq_len = st.seq_len.view(n_img, beam)
        scores = st.scores.view(n_img, beam) / seq_len                      # :110
        order = scores.argsort(dim=1, descending=True)                      # :111
        cur = order
        ids = torch.empty(n_img, beam, n_sel, dtype=torch.long, device=tok.device)
        for s in range(n_sel - 1, -1, -1):                                  # vectorised over images and beams
            ids[:, :, s] = tok[s].gather(1, cur)
            cur = par[s].gather(1, cur)
        lens = seq_len.gather(1, order)
        ids_l, sc_l, len_l = ids.tolist(), scores.gather(1, order).tolist(), lens.tolist()
        return [([row[: int(n)] for row, n in zip(ids_l[i], len_l[i])], sc_l[i], len_l[i]) for i in range(n_img)]


def generate_beam_ids(model, embed: torch.Tensor, beam_size: int = 5, entry_length: int = 67, temperature: float = 1.0,
                      stop_token_index: int = 13, use_cuda_graph: bool = True):
    """Batched id-level beam search.  embed [n_img, P, d] -> per image (beams ordered best-first, scores, lengths)."""
    eng = model.engine()
    embed = embed if embed.dim() == 3 else embed.view(-1, model.prefix_length, eng.d)
    n_img, P = embed.shape[0], embed.shape[1]
    cache = eng.__dict__.setdefault("_beam_decoders", {})
    key = (n_img, beam_size, P, entry_length, use_cuda_graph)
    dec = cache.get(key)
    if dec is None:
        cache.clear()  # one live configuration: the K/V cache is the big allocation
        dec = cache[key] = BeamDecoder(eng, n_img, beam_size, P, entry_length, use_cuda_graph)
    return dec.run(embed, temperature, stop_token_index)


def generate_beam(model, tokenizer, beam_size: int = 5, prompt=None, embed=None, entry_length: int = 67,
                  temperature: float = 1.0, stop_token: str = "."):
    """Same signature and return value as gpt2_prefix_eval.py:50-52 (list of decoded captions, best first)."""
    model.eval()
    stop_token_index = tokenizer.encode(stop_token)[0]
    head = []
    if embed is None:
        if prompt is None:
            raise ValueError("generate_beam needs `embed` or `prompt`")
        dev = model.engine().dev
        head = [int(t) for t in tokenizer.encode(prompt)]
        ids = torch.tensor(head, device=dev).unsqueeze(0)  # :65-68
        embed = model.gpt.transformer.wte(ids)
    (beams, _, _), = generate_beam_ids(model, embed.view(1, -1, embed.shape[-1]), beam_size, entry_length, temperature,
                                        stop_token_index)
    # prompt path of the reference (:82-88, :111-112): `tokens` starts as the prompt ids, the generated ids are appended,
    # and each beam is cut to `seq_length` tokens, a count that starts at 1 and only counts GENERATED tokens - so the
    # decoded text is the first `seq_length` ids of prompt + generated.  Reproduced as is.  (embed path: head is empty.)
    return [tokenizer.decode((head + ids)[: len(ids)]) for ids in beams]


def generate_beam_batch(model, tokenizer, embeds: torch.Tensor, beam_size: int = 5, entry_length: int = 67,
                        temperature: float = 1.0, stop_token: str = "."):
    """`generate_beam` for MANY images at once (BASELINE C5: the reference's make_preds loop, predictions_runner.py:213-233,
    decodes one image per call): embeds [n_img, P, d] = `model.clip_project(prefix).reshape(n_img, prefix_length, -1)`.
    Returns, per image, the list gpt2_prefix_eval.py:111-114 returns (decoded captions, best beam first); every image
    follows the reference's per-image recurrence exactly, so entry i equals `generate_beam(..., embed=embeds[i:i+1])`."""
    model.eval()
    stop_token_index = tokenizer.encode(stop_token)[0]
    embeds = embeds if embeds.dim() == 3 else embeds.view(-1, model.prefix_length, embeds.shape[-1])
    res = generate_beam_ids(model, embeds, beam_size, entry_length, temperature, stop_token_index)
    return [[tokenizer.decode(ids) for ids in beams] for beams, _, _ in res]


_GREEDY_EXTRA_STOP = 764   # gpt2_prefix_eval.py:187: generate2 also stops on token id 764


def generate_greedy_ids(model, embed: torch.Tensor, entry_length: int = 67, temperature: float = 1.0,
                        stop_token_index: int = 13):
    """Batched greedy decode = what gpt2_prefix_eval.generate2 (:118-198) computes: its top-p mask always keeps the best
    logit (:169) and the next token is the ARGMAX of the masked logits (:177, the multinomial draw is commented out), so
    `top_p` cannot change the result.  Runs the KV-cached decoder with one beam (a 1-beam search picks
    argmax((score + log p) / len) = argmax p at every step and stops right after the stop token) and cuts each sequence
    after the first stop token or token 764, both included, like :181-188.  embed [n_img, P, d] -> n_img id lists."""
    res = generate_beam_ids(model, embed, 1, entry_length, temperature, stop_token_index)
    out = []
    for beams, _, _ in res:
        ids = list(beams[0])
        for k, t in enumerate(ids):
            if t == stop_token_index or t == _GREEDY_EXTRA_STOP:
                ids = ids[: k + 1]
                break
        out.append(ids)
    return out


def generate2(model, tokenizer, tokens=None, prompt=None, embed=None, entry_count: int = 1, entry_length: int = 67,
              top_p: float = 0.8, temperature: float = 1.0, stop_token: str = "."):
    """Same signature and return value as gpt2_prefix_eval.py:118-198 (one decoded caption).  `top_p` and `entry_count`
    are accepted for compatibility: the reference's own code makes them no-ops (argmax after the mask; only
    generated_list[0] is returned)."""
    model.eval()
    stop_token_index = tokenizer.encode(stop_token)[0]
    head = []
    if embed is None:
        if tokens is None:
            if prompt is None:
                raise ValueError("generate2 needs `embed`, `tokens` or `prompt`")
            tokens = torch.tensor(tokenizer.encode(prompt)).unsqueeze(0)          # :143-145
        head = [int(t) for t in tokens.reshape(-1).tolist()]
        embed = model.gpt.transformer.wte(tokens.to(model.engine().dev))           # :150
    ids, = generate_greedy_ids(model, embed.reshape(1, -1, embed.shape[-1]), entry_length, temperature, stop_token_index)
    return tokenizer.decode(head + ids)                                             # :189-190
