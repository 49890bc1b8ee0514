"""CPU oracle for the CapDec training-step hot path — TEST INFRASTRUCTURE ONLY.

A plain-torch restatement (CPU, fp32 or fp64) of exactly what the reference executes per batch
(train.py:345-354): noise_injection -> ClipCaptionModel.forward -> logits slice -> masked cross entropy ->
backward (torch autograd here) -> HF AdamW + linear warm-up schedule.  GPT-2 is restated from the HuggingFace
source the reference calls (`HF:` = transformers/models/gpt2/modeling_gpt2.py, pytorch_utils.py, activations.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module, and only as the checker / the timed CPU baseline — never on the product path (capdec_b200 raises when
its CUDA library is missing instead of falling back here).

Pinning: the reference has no tests or golden vectors.  oracle/pin_against_reference.py runs the
reference's OWN classes (imported from /root/reference/train.py + transformers' GPT2LMHeadModel, shimmed as in
SURVEY §8c) on the deterministic weights/inputs below, checks this restatement against them and writes
tests/golden/*.json; tests/test_oracle_golden.py re-checks the restatement against those files on every run.

Parameters are passed as a flat dict keyed by the reference's state_dict names (train.py:359-371 layout).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

GPT2_SMALL = dict(n_layer=12, n_head=12, n_embd=768, vocab=50257, n_pos=1024, eps=1e-5)


# --------------------------------------------------------------------------------------------------------------
# deterministic weights / inputs shared by the reference run, the oracle and the CUDA model
# --------------------------------------------------------------------------------------------------------------
def make_state_dict(seed: int = 0, mapping_type: str = "mlp", prefix_length: int = 10, clip_length: int = 10,
                    prefix_size: int = 512, num_layers: int = 8, n_layer: int = 12, vocab: int = 50257,
                    n_embd: int = 768, n_pos: int = 1024, weight_std: float = 0.02, dtype=torch.float32) -> Params:
    """Seeded weights in the reference checkpoint layout.  Shapes follow SURVEY §8b.  Init follows HF GPT-2
    (HF:modeling_gpt2.py:433-458: N(0, 0.02), c_proj N(0, 0.02/sqrt(2 n_layer))) but LayerNorm gains/biases and
    linear biases are randomised too so that every parameter path is exercised by the parity tests."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *shape, std=1.0: (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)
    d = n_embd
    sd: Params = {}
    sd["gpt.transformer.wte.weight"] = rn(vocab, d, std=weight_std)
    sd["gpt.transformer.wpe.weight"] = rn(n_pos, d, std=weight_std)
    for i in range(n_layer):
        p = f"gpt.transformer.h.{i}."
        sd[p + "ln_1.weight"] = 1.0 + rn(d, std=0.1)
        sd[p + "ln_1.bias"] = rn(d, std=0.05)
        sd[p + "attn.c_attn.weight"] = rn(d, 3 * d, std=weight_std)
        sd[p + "attn.c_attn.bias"] = rn(3 * d, std=0.02)
        sd[p + "attn.c_proj.weight"] = rn(d, d, std=weight_std / math.sqrt(2 * n_layer))
        sd[p + "attn.c_proj.bias"] = rn(d, std=0.02)
        sd[p + "ln_2.weight"] = 1.0 + rn(d, std=0.1)
        sd[p + "ln_2.bias"] = rn(d, std=0.05)
        sd[p + "mlp.c_fc.weight"] = rn(d, 4 * d, std=weight_std)
        sd[p + "mlp.c_fc.bias"] = rn(4 * d, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(4 * d, d, std=weight_std / math.sqrt(2 * n_layer))
        sd[p + "mlp.c_proj.bias"] = rn(d, std=0.02)
    sd["gpt.transformer.ln_f.weight"] = 1.0 + rn(d, std=0.1)
    sd["gpt.transformer.ln_f.bias"] = rn(d, std=0.05)
    sd["gpt.lm_head.weight"] = sd["gpt.transformer.wte.weight"]  # tied (HF:modeling_gpt2.py:646-651)
    P = prefix_length
    if mapping_type == "mlp":  # train.py:269-270: MLP((D, d*P//2, d*P))
        hdim = (d * P) // 2
        sd["clip_project.model.0.weight"] = rn(hdim, prefix_size, std=1.0 / math.sqrt(prefix_size))
        sd["clip_project.model.0.bias"] = rn(hdim, std=0.02)
        sd["clip_project.model.2.weight"] = rn(d * P, hdim, std=1.0 / math.sqrt(hdim))
        sd["clip_project.model.2.bias"] = rn(d * P, std=0.02)
    else:  # train.py:238-243 TransformerMapper
        C = clip_length
        sd["clip_project.prefix_const"] = rn(P, d)
        sd["clip_project.linear.weight"] = rn(C * d, prefix_size, std=1.0 / math.sqrt(prefix_size))
        sd["clip_project.linear.bias"] = rn(C * d, std=0.02)
        hm = int(d * 2.0)  # mlp_ratio = 2.0 (train.py:212)
        for j in range(num_layers):
            p = f"clip_project.transformer.layers.{j}."
            sd[p + "norm1.weight"] = 1.0 + rn(d, std=0.1)
            sd[p + "norm1.bias"] = rn(d, std=0.05)
            sd[p + "attn.to_queries.weight"] = rn(d, d, std=1.0 / math.sqrt(d))
            sd[p + "attn.to_keys_values.weight"] = rn(2 * d, d, std=1.0 / math.sqrt(d))
            sd[p + "attn.project.weight"] = rn(d, d, std=1.0 / math.sqrt(d))
            sd[p + "attn.project.bias"] = rn(d, std=0.02)
            sd[p + "norm2.weight"] = 1.0 + rn(d, std=0.1)
            sd[p + "norm2.bias"] = rn(d, std=0.05)
            sd[p + "mlp.fc1.weight"] = rn(hm, d, std=1.0 / math.sqrt(d))
            sd[p + "mlp.fc1.bias"] = rn(hm, std=0.02)
            sd[p + "mlp.fc2.weight"] = rn(d, hm, std=1.0 / math.sqrt(hm))
            sd[p + "mlp.fc2.bias"] = rn(d, std=0.02)
    return sd


def make_batch(seed: int, B: int, L: int = 40, prefix_size: int = 512, vocab: int = 50257, min_len: int = 8,
               full_length: bool = False):
    """Synthetic batch of SURVEY §8d: normalised Gaussian 'CLIP' embeddings, token ids uniform in [1, vocab),
    per-row length ~ U{min_len..L}, right-padded with id 0; mask = cat(ones(P), tokens > 0) is built by callers
    (train.py:55-63).  Also returns the Gaussian draw for the noise injection (unit variance; scale by std)."""
    g = torch.Generator().manual_seed(seed)
    prefix = torch.randn(B, prefix_size, generator=g)
    prefix = prefix / prefix.norm(2, -1, keepdim=True)  # train.py:69-71
    tokens = torch.randint(1, vocab, (B, L), generator=g, dtype=torch.int64)
    if not full_length:
        lens = torch.randint(min_len, L + 1, (B,), generator=g)
        tokens[torch.arange(L)[None, :] >= lens[:, None]] = 0
    noise = torch.randn(B, prefix_size, generator=g)
    return tokens, prefix, noise


def make_caption_table(seed: int, n: int, n_emb: int, prefix_size: int = 512, vocab: int = 50257, min_len: int = 3,
                       max_len: int = 60, half: bool = False):
    """Synthetic stand-in for the pickled dataset of train.py:74-103: ragged int64 token lists (ids may include 0, the
    '!' token), a caption -> embedding index, and an un-normalised embedding table (fp32, or fp16 like CLIP's output)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, max_len + 1, (n,), generator=g)
    caps = [torch.randint(0, vocab, (int(k),), generator=g, dtype=torch.int64) for k in lens]
    cap2emb = torch.randint(0, n_emb, (n,), generator=g).tolist()
    table = torch.randn(n_emb, prefix_size, generator=g) * 0.7
    return caps, cap2emb, (table.half() if half else table)


def dataset_max_seq_len(captions_tokens) -> int:
    """train.py:102-103."""
    all_len = torch.tensor([len(t) for t in captions_tokens]).float()
    return min(int(all_len.mean() + all_len.std() * 10), int(all_len.max()))


def dataset_item(captions_tokens, caption2embedding, prefixes, item: int, max_seq_len: int, prefix_length: int,
                 normalize_prefix: bool):
    """ClipCocoDataset.pad_tokens + __getitem__, train.py:52-72 (without its in-place caching side effect)."""
    tokens = captions_tokens[item]
    padding = max_seq_len - tokens.shape[0]
    if padding > 0:
        tokens = torch.cat((tokens, torch.zeros(padding, dtype=torch.int64) - 1))
    elif padding < 0:
        tokens = tokens[:max_seq_len]
    tokens = tokens.clone()
    mask = tokens.ge(0)
    tokens[~mask] = 0
    mask = torch.cat((torch.ones(prefix_length), mask.float()), dim=0)
    prefix = prefixes[caption2embedding[item]]
    if normalize_prefix:
        prefix = prefix.float()
        prefix = prefix / prefix.norm(2, -1)
    return tokens, mask, prefix


def make_mask(tokens: torch.Tensor, prefix_length: int) -> torch.Tensor:
    """train.py:58-63 (mask is 0 on padding; synthetic padding == id 0)."""
    return torch.cat((torch.ones(tokens.shape[0], prefix_length), (tokens > 0).float()), dim=1)


# --------------------------------------------------------------------------------------------------------------
# restatement of the model
# --------------------------------------------------------------------------------------------------------------
def noise_injection(x, variance=0.001, modality_offset=None, uniform_noise=False, dont_norm=False, noise=None,
                    generator=None):
    """train.py:27-39.  `noise` = unit-variance Gaussian draw to use instead of torch.randn (parity mode)."""
    if variance == 0.0:
        return x  # train.py:28-29
    std = math.sqrt(variance)
    if not dont_norm:
        x = F.normalize(x, dim=1)  # train.py:31-32
    if uniform_noise:  # train.py:18-24
        gdraw = noise if noise is not None else torch.randn(x.shape, generator=generator, dtype=x.dtype)
        sphere = F.normalize(gdraw, dim=1)
        u = torch.rand(x.shape[0], generator=generator, dtype=x.dtype) ** (1.0 / x.shape[1])
        x = x + (sphere.T * u * std).T
    else:
        gdraw = noise if noise is not None else torch.randn(x.shape, generator=generator, dtype=x.dtype)
        x = x + gdraw * std  # train.py:36
    if modality_offset is not None:
        x = x + modality_offset  # train.py:37-38
    return F.normalize(x, dim=1)  # train.py:39


def gelu_new(x):
    """HF:activations.py:59-66."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def conv1d(x, w, b):
    """HF:pytorch_utils.py:97-123: y = x @ W + b with W [in, out]."""
    return torch.addmm(b, x.reshape(-1, x.shape[-1]), w).view(*x.shape[:-1], w.shape[1])


def gpt2_forward(sd: Params, inputs_embeds, attention_mask=None, n_head: int = 12, eps: float = 1e-5,
                 p_drop: float = 0.0, return_hidden: bool = False):
    """HF GPT2LMHeadModel.forward(inputs_embeds=..., attention_mask=...) in eval mode (or p_drop via torch RNG):
    HF:modeling_gpt2.py:579-636 (embeddings, blocks, ln_f), :262-309 (block), :54-72 + :185-226 (attention),
    :229-243 (MLP), :703-706 (tied lm_head)."""
    B, T, d = inputs_embeds.shape
    hd = d // n_head
    drop = (lambda t: F.dropout(t, p_drop, True)) if p_drop > 0 else (lambda t: t)
    h = inputs_embeds + sd["gpt.transformer.wpe.weight"][:T].unsqueeze(0)  # position_ids = arange(T)
    h = drop(h)
    # additive mask: causal (-inf above the diagonal) + padding ((1 - mask) * finfo.min on keys)
    neg = torch.finfo(h.dtype).min
    causal = torch.ones(T, T, dtype=torch.bool).tril()
    add_mask = torch.zeros(1, 1, T, T, dtype=h.dtype).masked_fill(~causal, neg)
    if attention_mask is not None:
        add_mask = add_mask + (1.0 - attention_mask.to(h.dtype))[:, None, None, :] * neg
        add_mask = add_mask.clamp_min(neg)
    n_layer = 0
    while f"gpt.transformer.h.{n_layer}.ln_1.weight" in sd:
        n_layer += 1
    for i in range(n_layer):
        p = f"gpt.transformer.h.{i}."
        x = F.layer_norm(h, (d,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps)
        qkv = conv1d(x, sd[p + "attn.c_attn.weight"], sd[p + "attn.c_attn.bias"])
        q, k, v = qkv.split(d, dim=2)
        q = q.view(B, T, n_head, hd).transpose(1, 2)
        k = k.view(B, T, n_head, hd).transpose(1, 2)
        v = v.view(B, T, n_head, hd).transpose(1, 2)
        w = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5) + add_mask
        w = drop(F.softmax(w, dim=-1))
        a = torch.matmul(w, v).transpose(1, 2).reshape(B, T, d)
        a = conv1d(a, sd[p + "attn.c_proj.weight"], sd[p + "attn.c_proj.bias"])
        h = h + drop(a)
        x = F.layer_norm(h, (d,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps)
        m = gelu_new(conv1d(x, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        m = conv1d(m, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        h = h + drop(m)
    hf = F.layer_norm(h, (d,), sd["gpt.transformer.ln_f.weight"], sd["gpt.transformer.ln_f.bias"], eps)
    logits = F.linear(hf, sd["gpt.transformer.wte.weight"])  # tied lm_head, no bias
    return (logits, hf) if return_hidden else logits


def mlp_mapper(sd: Params, x):
    """train.py:106-118 as built at :269-270: Linear -> Tanh -> Linear."""
    h = torch.tanh(F.linear(x, sd["clip_project.model.0.weight"], sd["clip_project.model.0.bias"]))
    return F.linear(h, sd["clip_project.model.2.weight"], sd["clip_project.model.2.bias"])


def transformer_mapper(sd: Params, x, clip_length: int, num_heads: int = 8):
    """train.py:229-243 -> :192-226 -> :170-189 -> :138-167 / :121-136 (no mask, dropout p=0)."""
    B = x.shape[0]
    pc = sd["clip_project.prefix_const"]
    P, d = pc.shape
    x = F.linear(x, sd["clip_project.linear.weight"], sd["clip_project.linear.bias"]).view(B, clip_length, d)
    h = torch.cat((x, pc.unsqueeze(0).expand(B, P, d)), dim=1)
    n = h.shape[1]
    hd = d // num_heads
    j = 0
    while f"clip_project.transformer.layers.{j}.norm1.weight" in sd:
        p = f"clip_project.transformer.layers.{j}."
        y = F.layer_norm(h, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        q = F.linear(y, sd[p + "attn.to_queries.weight"]).reshape(B, n, num_heads, hd)
        kv = F.linear(y, sd[p + "attn.to_keys_values.weight"]).reshape(B, n, 2, num_heads, hd)
        k, v = kv[:, :, 0], kv[:, :, 1]
        att = torch.einsum("bnhd,bmhd->bnmh", q, k) * (hd ** -0.5)
        att = att.softmax(dim=2)
        o = torch.einsum("bnmh,bmhd->bnhd", att, v).reshape(B, n, d)
        h = h + F.linear(o, sd[p + "attn.project.weight"], sd[p + "attn.project.bias"])
        y = F.layer_norm(h, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        y = F.relu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        h = h + F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        j += 1
    return h[:, clip_length:]


def _mapper_layer(sd: Params, p: str, x, y=None, num_heads: int = 8):
    """transformer_mapper.TransformerLayer.forward, transformer_mapper.py:63-66 -> MultiHeadAttention :34-51 + Mlp :13-19.
    `y` = key/value source; None -> norm1(x) (what MultiHeadAttention sees as its own `x`)."""
    B, n, c = x.shape
    hd = c // num_heads
    xn = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    y = xn if y is None else y
    m = y.shape[1]
    q = F.linear(xn, sd[p + "attn.to_queries.weight"]).reshape(B, n, num_heads, hd)
    kv = F.linear(y, sd[p + "attn.to_keys_values.weight"]).reshape(B, m, 2, num_heads, hd)
    k, v = kv[:, :, 0], kv[:, :, 1]
    att = (torch.einsum("bnhd,bmhd->bnmh", q, k) * (hd ** -0.5)).softmax(dim=2)
    o = torch.einsum("bnmh,bmhd->bnhd", att, v).reshape(B, n, c)
    x = x + F.linear(o, sd[p + "attn.project.weight"], sd[p + "attn.project.bias"])
    h = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    h = F.relu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def encdec_mapper(sd: Params, x, clip_length: int, num_heads: int = 8):
    """transformer_mapper.TransformerEncoderDecoder.forward, transformer_mapper.py:132-137 (gpt2_prefix.py:167-168,
    MappingType.TransformerDecoder): ref = ref_encoder(linear(x) as clip_length x 512 tokens);
    prefix = prefix_decoder(prefix_const, ref) with layers alternating cross (keys/values from ref, :87-88) and 'self'
    (:89-90 - keys/values from the UN-normalised stream x, queries from norm1(x))."""
    B = x.shape[0]
    pc = sd["clip_project.prefix_const"]
    de = sd["clip_project.ref_encoder.layers.0.norm1.weight"].shape[0]
    ref = F.linear(x, sd["clip_project.linear.weight"], sd["clip_project.linear.bias"]).view(B, clip_length, de)
    j = 0
    while f"clip_project.ref_encoder.layers.{j}.norm1.weight" in sd:
        ref = _mapper_layer(sd, f"clip_project.ref_encoder.layers.{j}.", ref, None, num_heads)   # :91-92 (enc_dec False)
        j += 1
    h = pc.unsqueeze(0).expand(B, *pc.shape)
    j = 0
    while f"clip_project.prefix_decoder.layers.{j}.norm1.weight" in sd:
        p = f"clip_project.prefix_decoder.layers.{j}."
        h = _mapper_layer(sd, p, h, ref if j % 2 == 0 else h, num_heads)
        j += 1
    return h


def make_encdec_state_dict(seed: int, prefix_length: int = 10, clip_length: int = 10, prefix_size: int = 512,
                           num_layers: int = 2, d: int = 768, de: int = 512) -> Params:
    """Seeded parameters with the key layout of transformer_mapper.TransformerEncoderDecoder (under `clip_project.`)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    sd: Params = {"clip_project.prefix_const": rn(prefix_length, d, std=1.0),
                  "clip_project.linear.weight": rn(clip_length * de, prefix_size, std=1.0 / math.sqrt(prefix_size)),
                  "clip_project.linear.bias": rn(clip_length * de, std=0.02)}

    def layer(p, c, c_ref):
        hm = int(c * 2.0)
        sd[p + "norm1.weight"] = 1.0 + rn(c, std=0.1); sd[p + "norm1.bias"] = rn(c, std=0.05)
        sd[p + "attn.to_queries.weight"] = rn(c, c, std=1.0 / math.sqrt(c))
        sd[p + "attn.to_keys_values.weight"] = rn(2 * c, c_ref, std=1.0 / math.sqrt(c_ref))
        sd[p + "attn.project.weight"] = rn(c, c, std=1.0 / math.sqrt(c)); sd[p + "attn.project.bias"] = rn(c, std=0.02)
        sd[p + "norm2.weight"] = 1.0 + rn(c, std=0.1); sd[p + "norm2.bias"] = rn(c, std=0.05)
        sd[p + "mlp.fc1.weight"] = rn(hm, c, std=1.0 / math.sqrt(c)); sd[p + "mlp.fc1.bias"] = rn(hm, std=0.02)
        sd[p + "mlp.fc2.weight"] = rn(c, hm, std=1.0 / math.sqrt(hm)); sd[p + "mlp.fc2.bias"] = rn(c, std=0.02)

    for j in range(num_layers):
        layer(f"clip_project.ref_encoder.layers.{j}.", de, de)
    for j in range(2 * num_layers):
        layer(f"clip_project.prefix_decoder.layers.{j}.", d, de if j % 2 == 0 else d)
    return sd


def make_bridger_state_dict(seed: int, dim: int = 640, num_layers: int = 8, dtype=torch.float32) -> Params:
    """Seeded weights in the layout of others/weights_modality_mapper.pt (`layers.{i}.weight [dim,dim]`, `.bias [dim]`):
    near-identity matrices plus noise, so that the ReLU stack neither dies nor explodes."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(num_layers):
        sd[f"layers.{i}.weight"] = (torch.eye(dim) + 0.05 * torch.randn(dim, dim, generator=g)).to(dtype)
        sd[f"layers.{i}.bias"] = (0.02 * torch.randn(dim, generator=g)).to(dtype)
    return sd


def modality_bridger(sd: Params, x, num_layers: int = 8):
    """others/supervised_embedding_bridger.py:103-108: x <- relu(layer_i(x)) for every layer but the last, which is linear
    (used by predictions_runner.py:225-227 before clip_project)."""
    for i in range(num_layers):
        x = F.linear(x, sd[f"layers.{i}.weight"], sd[f"layers.{i}.bias"])
        if i < num_layers - 1:
            x = F.relu(x)
    return x


def clip_project(sd: Params, prefix, clip_length: Optional[int] = None):
    if "clip_project.model.0.weight" in sd:
        return mlp_mapper(sd, prefix)
    if "clip_project.ref_encoder.layers.0.norm1.weight" in sd:
        return encdec_mapper(sd, prefix, clip_length)
    return transformer_mapper(sd, prefix, clip_length)


def clipcap_forward(sd: Params, tokens, prefix, mask=None, prefix_length: int = 10,
                    clip_length: Optional[int] = None, p_drop: float = 0.0):
    """ClipCaptionModel.forward, train.py:251-260 -> logits [B, P+L, V]."""
    d = sd["gpt.transformer.wte.weight"].shape[1]
    emb_text = sd["gpt.transformer.wte.weight"][tokens]  # train.py:253
    proj = clip_project(sd, prefix, clip_length).view(-1, prefix_length, d)  # train.py:254
    emb = torch.cat((proj, emb_text), dim=1)  # train.py:255
    return gpt2_forward(sd, emb, attention_mask=mask, p_drop=p_drop)  # train.py:259


def caption_loss(logits, tokens, prefix_length: int):
    """train.py:349-350."""
    lg = logits[:, prefix_length - 1: -1]
    return F.cross_entropy(lg.reshape(-1, lg.shape[-1]), tokens.flatten(), ignore_index=0)


def loss_and_grads(sd: Params, tokens, prefix, mask, prefix_length, clip_length=None, trainable=None):
    """One forward/backward of train.py:348-351 through autograd. Returns (loss, logits, {name: grad})."""
    leaves = {}
    for k, v in sd.items():
        if k == "gpt.lm_head.weight":
            continue
        if trainable is None or trainable(k):
            leaves[k] = v.detach().clone().requires_grad_(True)
        else:
            leaves[k] = v.detach()
    leaves["gpt.lm_head.weight"] = leaves["gpt.transformer.wte.weight"]
    logits = clipcap_forward(leaves, tokens, prefix, mask, prefix_length, clip_length)
    loss = caption_loss(logits, tokens, prefix_length)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if k != "gpt.lm_head.weight" and v.requires_grad and v.grad is not None}
    return loss.detach(), logits.detach(), grads


# --------------------------------------------------------------------------------------------------------------
# decoding: beam search exactly as gpt2_prefix_eval.py:50-115, full re-forward per step like the reference
# --------------------------------------------------------------------------------------------------------------
def generate_beam(sd: Params, embed, beam_size: int = 5, entry_length: int = 67, temperature: float = 1.0,
                  stop_token_index: int = 13):
    """Returns (token id lists ordered by final score, scores / seq_lengths in that order, seq_lengths in that order).
    The reference decodes the ids with the GPT-2 tokenizer (gpt2_prefix_eval.py:111-112); ids are the comparable part."""
    wte = sd["gpt.transformer.wte.weight"]
    tokens = None
    scores = None
    seq_lengths = torch.ones(beam_size)
    is_stopped = torch.zeros(beam_size, dtype=torch.bool)
    generated = embed  # [1, P, d]
    for _ in range(entry_length):
        logits = gpt2_forward(sd, generated)                                         # :75-76
        logits = logits[:, -1, :] / (temperature if temperature > 0 else 1.0)         # :77
        logits = logits.softmax(-1).log()                                            # :78-79
        if scores is None:
            scores, next_tokens = logits.topk(beam_size, -1)                         # :81
            generated = generated.expand(beam_size, *generated.shape[1:])
            next_tokens, scores = next_tokens.permute(1, 0), scores.squeeze(0)
            tokens = next_tokens
        else:
            logits[is_stopped] = -float("inf")                                       # :90
            logits[is_stopped, 0] = 0                                                # :91
            scores_sum = scores[:, None] + logits
            seq_lengths[~is_stopped] += 1
            scores_sum_average = scores_sum / seq_lengths[:, None]
            scores_sum_average, next_tokens = scores_sum_average.view(-1).topk(beam_size, -1)
            next_tokens_source = next_tokens // scores_sum.shape[1]                  # :96
            seq_lengths = seq_lengths[next_tokens_source]
            next_tokens = (next_tokens % scores_sum.shape[1]).unsqueeze(1)
            tokens = torch.cat((tokens[next_tokens_source], next_tokens), dim=1)
            generated = generated[next_tokens_source]
            scores = scores_sum_average * seq_lengths
            is_stopped = is_stopped[next_tokens_source]
        nxt = wte[next_tokens.squeeze()].view(generated.shape[0], 1, -1)             # :105
        generated = torch.cat((generated, nxt), dim=1)
        is_stopped = is_stopped + next_tokens.eq(stop_token_index).squeeze()
        if is_stopped.all():
            break
    scores = scores / seq_lengths                                                     # :110
    order = scores.argsort(descending=True)
    out = [tokens[i, : int(seq_lengths[i])].tolist() for i in order]
    return out, scores[order].tolist(), seq_lengths[order].tolist()


def generate2(sd: Params, embed, entry_length: int = 67, top_p: float = 0.8, temperature: float = 1.0,
              stop_token_index: int = 13):
    """gpt2_prefix_eval.generate2 (:118-198) for `embed` given, entry_count = 1: nucleus filtering followed by ARGMAX
    (:176 — the multinomial draw is commented out at :177), so the top-p mask never changes the chosen token (the best
    logit is always kept, :169) and the function is greedy decoding with a full re-forward per token; it stops after
    emitting the stop token or token 764 (:185).  Returns the generated token ids (the reference decodes them, :189-190)."""
    wte = sd["gpt.transformer.wte.weight"]
    generated = embed                                                                 # [1, P, d]
    tokens = []
    for _ in range(entry_length):
        logits = gpt2_forward(sd, generated)[:, -1, :] / (temperature if temperature > 0 else 1.0)      # :163-165
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)                            # :166
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)                      # :167
        remove = cumulative_probs > top_p
        remove[..., 1:] = remove[..., :-1].clone()                                                     # :169-171
        remove[..., 0] = 0
        logits[:, sorted_indices[remove]] = -float("inf")                                             # :174-175
        nxt = int(torch.argmax(logits, -1))                                                            # :177
        tokens.append(nxt)
        generated = torch.cat((generated, wte[nxt].view(1, 1, -1)), dim=1)                             # :181-186
        if nxt == stop_token_index or nxt == 764:                                                     # :187
            break
    return tokens


# --------------------------------------------------------------------------------------------------------------
# optimizer restatement
# --------------------------------------------------------------------------------------------------------------
def hf_adamw_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0):
    """transformers==4.24 optimization.AdamW.step (correct_bias=True), used at train.py:326,352. In place."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


def linear_warmup_lr(base_lr: float, step: int, warmup: int, total: int) -> float:
    """get_linear_schedule_with_warmup (HF:optimization.py:101-104), value used for optimizer step number `step`
    (0-based count of scheduler.step() calls already made; train.py:328-330,353)."""
    if step < warmup:
        return base_lr * float(step) / float(max(1, warmup))
    return base_lr * max(0.0, float(total - step) / float(max(1, total - warmup)))


class HFAdamW(torch.optim.Optimizer):
    """Drop-in for `transformers.AdamW` (removed from transformers >= 4.5x; train.py:6 imports it)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                hf_adamw_step(p, p.grad, st["exp_avg"], st["exp_avg_sq"], st["step"], group["lr"], *group["betas"],
                              group["eps"], group["weight_decay"])
