"""The reference's Python class surface on top of the sm_100a engine.

Mirrors train.py / gpt2_prefix.py of DavidHuji/CapDec: `MappingType`, `noise_injection`, `MLP`, `TransformerMapper`,
`ClipCaptionModel`, `ClipCaptionPrefix`, with the reference's constructor signatures, attribute surface
(`model.gpt.transformer.wte`, `model.clip_project`, `model.prefix_length`, ...) and checkpoint layout
(state_dict keys/shapes of train.py:359-371, incl. the tied `gpt.lm_head.weight` duplicate).  The modules only HOLD
parameters (as views into one flat fp32 buffer); all arithmetic runs in `engine.Engine` through the C-ABI kernels.
"""
from __future__ import annotations

import math
from enum import Enum
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import CapdecError


class MappingType(Enum):
    """train.py:42-44 (values) + gpt2_prefix.py:15-18 aliases."""
    MLP = "mlp"
    Transformer = "transformer"
    TransformerEncoder = "transformer"  # gpt2_prefix.py spelling 'transformer_encoder' maps to the same mapper
    TransformerDecoder = "transformer_decoder"  # gpt2_prefix.py:18 -> transformer_mapper.TransformerEncoderDecoder

    @classmethod
    def parse(cls, v):
        if isinstance(v, cls):
            return v
        v = getattr(v, "value", v)
        if v == "mlp":
            return cls.MLP
        if v in ("transformer", "transformer_encoder"):
            return cls.Transformer
        if v == "transformer_decoder":
            return cls.TransformerDecoder
        raise ValueError(f"unsupported mapping type {v!r}")


_noise_state = {"seed": None, "step": 0}


def noise_injection(x, variance=0.001, modality_offset=None, uniform_noise=False, dont_norm=False):
    """train.py:27-39 as one fused kernel (normalise -> Philox noise -> offset -> normalise).  variance == 0 returns
    `x` itself, un-normalised, exactly like the reference (train.py:28-29)."""
    if variance == 0.0:
        return x
    if not x.is_cuda:
        raise CapdecError("noise_injection: capdec_b200 runs on CUDA tensors only (no CPU fallback)")
    x = x.contiguous().float()
    if _noise_state["seed"] is None or _noise_state["seed"].device != x.device:
        _noise_state["seed"] = ops.make_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, x.device)
    out = torch.empty_like(x)
    off = None
    if modality_offset is not None:
        off = modality_offset.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
    _noise_state["step"] += 1
    ops.noise_injection(x, out, variance, offset=off, uniform_ball=uniform_noise, dont_norm=dont_norm,
                        seed=_noise_state["seed"], step=_noise_state["step"])
    return out


# ----------------------------------------------------------------------------------------------------------------
# parameter holders with the reference's names
# ----------------------------------------------------------------------------------------------------------------
class Conv1D(nn.Module):
    """HF Conv1D parameter layout: weight [in, out], bias [out] (HF:pytorch_utils.py:97-123)."""

    def __init__(self, nf: int, nx: int):
        super().__init__()
        self.nf = nf
        self.weight = nn.Parameter(torch.empty(nx, nf))
        self.bias = nn.Parameter(torch.zeros(nf))


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.c_attn = Conv1D(3 * d, d)
        self.c_proj = Conv1D(d, d)


class _Mlp(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.c_fc = Conv1D(f, d)
        self.c_proj = Conv1D(d, f)


class _Block(nn.Module):
    def __init__(self, d, f, eps):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d, eps=eps)
        self.attn = _Attn(d)
        self.ln_2 = nn.LayerNorm(d, eps=eps)
        self.mlp = _Mlp(d, f)


class _Transformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.wte = nn.Embedding(cfg.vocab_size, cfg.n_embd)
        self.wpe = nn.Embedding(cfg.n_positions, cfg.n_embd)
        self.h = nn.ModuleList([_Block(cfg.n_embd, 4 * cfg.n_embd, cfg.layer_norm_epsilon) for _ in range(cfg.n_layer)])
        self.ln_f = nn.LayerNorm(cfg.n_embd, eps=cfg.layer_norm_epsilon)


class GPT2Config(SimpleNamespace):
    """Subset of transformers.GPT2Config defaults the path depends on (HF:configuration_gpt2.py:78-103)."""

    def __init__(self, **kw):
        base = dict(vocab_size=50257, n_positions=1024, n_embd=768, n_layer=12, n_head=12, layer_norm_epsilon=1e-5,
                    resid_pdrop=0.1, embd_pdrop=0.1, attn_pdrop=0.1, initializer_range=0.02)
        base.update(kw)
        super().__init__(**base)


class GPT2LMHead(nn.Module):
    """Parameter layout + call surface of transformers.GPT2LMHeadModel as the reference uses it:
    `gpt.transformer.wte(ids)`, `gpt(inputs_embeds=..., attention_mask=..., labels=...)` -> `.logits`/`.loss`,
    `gpt.get_input_embeddings()`, `state_dict()` keys `transformer.*` + tied `lm_head.weight`."""

    def __init__(self, config: Optional[GPT2Config] = None):
        super().__init__()
        self.config = config or GPT2Config()
        c = self.config
        self.transformer = _Transformer(c)
        self.lm_head = nn.Linear(c.n_embd, c.vocab_size, bias=False)
        self.lm_head.weight = self.transformer.wte.weight  # tie (HF:modeling_gpt2.py:646-651)
        self._owner = None  # set by ClipCaptionModel; a bare GPT2LMHead owns its own engine
        self.reset_parameters()

    def reset_parameters(self):
        """HF GPT-2 init (HF:modeling_gpt2.py:433-458)."""
        c = self.config
        std = c.initializer_range
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name.endswith("ln_f.weight"):
                    p.fill_(1.0)
                elif name.endswith("bias"):
                    p.zero_()
                elif name.endswith("c_proj.weight"):
                    p.normal_(0.0, std / math.sqrt(2 * c.n_layer))
                else:
                    p.normal_(0.0, std)

    def get_input_embeddings(self):
        return self.transformer.wte

    def forward(self, input_ids=None, inputs_embeds=None, attention_mask=None, labels=None, **_unused):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise CapdecError("GPT2LMHead must be attached to a ClipCaptionModel (its engine owns the activations)")
        if inputs_embeds is None:
            inputs_embeds = self.transformer.wte(input_ids)
        return owner._gpt_forward(inputs_embeds, attention_mask, labels)


class MLP(nn.Module):
    """train.py:106-118 — same Sequential layout so the keys are `model.0.*` / `model.2.*`."""

    def __init__(self, sizes: Tuple[int, ...], bias=True, act=nn.Tanh):
        super().__init__()
        if len(sizes) != 3 or act is not nn.Tanh or not bias:
            raise CapdecError("capdec_b200.MLP implements the mapper CapDec builds: 2 Linear layers, Tanh, bias")
        layers = []
        for i in range(len(sizes) - 1):
            layers.append(nn.Linear(sizes[i], sizes[i + 1], bias=bias))
            if i < len(sizes) - 2:
                layers.append(act())
        self.model = nn.Sequential(*layers)
        self._owner = None

    def forward(self, x):
        return self._owner()._mapper_forward(x)


class _MapperAttn(nn.Module):
    def __init__(self, d, d_ref=None):
        super().__init__()
        d_ref = d if d_ref is None else d_ref
        self.to_queries = nn.Linear(d, d, bias=False)        # train.py:144 (bias=False via :183,187)
        self.to_keys_values = nn.Linear(d_ref, 2 * d, bias=False)  # train.py:145 / transformer_mapper.py:30 (dim_ref)
        self.project = nn.Linear(d, d)                         # train.py:146


class _MapperMlp(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.fc1 = nn.Linear(d, h)
        self.fc2 = nn.Linear(h, d)


class _MapperLayer(nn.Module):
    def __init__(self, d, h, d_ref=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(d)
        self.attn = _MapperAttn(d, d_ref)
        self.norm2 = nn.LayerNorm(d)
        self.mlp = _MapperMlp(d, h)


class _MapperTransformer(nn.Module):
    """Parameter holder with the layout of train.py:192-226 / transformer_mapper.py:71-108 (`enc_dec` doubles the layer
    count and alternates cross-attention layers, whose keys/values read `d_ref` features, with self-attention ones)."""

    def __init__(self, d, num_layers, mlp_ratio=2.0, d_ref=None, enc_dec=False):
        super().__init__()
        self.enc_dec = enc_dec
        n = num_layers * 2 if enc_dec else num_layers
        self.layers = nn.ModuleList([
            _MapperLayer(d, int(d * mlp_ratio), d_ref if (not enc_dec or i % 2 == 0) else None) for i in range(n)])


class TransformerMapper(nn.Module):
    """train.py:229-243 (8 heads, mlp_ratio 2, pre-LN, no mask, no dropout)."""

    num_heads = 8

    def __init__(self, dim_clip: int, dim_embedding: int, prefix_length: int, clip_length: int, num_layers: int = 8):
        super().__init__()
        self.clip_length = clip_length
        self.transformer = _MapperTransformer(dim_embedding, num_layers)
        self.linear = nn.Linear(dim_clip, clip_length * dim_embedding)
        self.prefix_const = nn.Parameter(torch.randn(prefix_length, dim_embedding), requires_grad=True)
        self._owner = None

    def forward(self, x):
        return self._owner()._mapper_forward(x)


class TransformerEncoderDecoder(nn.Module):
    """transformer_mapper.py:130-145 (gpt2_prefix.py:167-168, MappingType.TransformerDecoder): a 512-wide encoder over
    the CLIP embedding re-shaped to `clip_length` tokens, and a decoder whose `prefix_length` learned query tokens
    cross-attend to it.  Inference-side mapper of predictions_runner.py:457-460; train.py cannot construct it, but
    gpt2_prefix.py:219-243 trains it, so forward AND backward run on the kernels (Engine._encdec_fwd / _encdec_bwd)."""

    num_heads = 8
    dim_ref = 512

    def __init__(self, dim_clip: int, dim_embedding: int, prefix_length: int, clip_length: int, num_layers: int = 4):
        super().__init__()
        self.clip_length = clip_length
        self.ref_encoder = _MapperTransformer(self.dim_ref, num_layers)
        self.prefix_decoder = _MapperTransformer(dim_embedding, num_layers, d_ref=self.dim_ref, enc_dec=True)
        self.linear = nn.Linear(dim_clip, clip_length * self.dim_ref)
        self.prefix_const = nn.Parameter(torch.randn(prefix_length, dim_embedding), requires_grad=True)
        self._owner = None

    def forward(self, x):
        return self._owner()._mapper_forward(x)


def _pretrained_gpt2_state_dict(name: str = "gpt2"):
    """`GPT2LMHeadModel.from_pretrained('gpt2')` exactly as the reference calls it (train.py:266, gpt2_prefix.py:162),
    returned as a state_dict with the 4.24-era mask buffers dropped.  Raises whatever transformers raises (no local copy
    while offline, transformers missing, ...)."""
    from transformers import GPT2LMHeadModel   # the reference's own dependency; only its checkpoint loader is used
    hf = GPT2LMHeadModel.from_pretrained(name)
    sd = {k: v.detach().to(torch.float32) for k, v in hf.state_dict().items()
          if not (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"))}
    if "lm_head.weight" not in sd:
        sd["lm_head.weight"] = sd["transformer.wte.weight"]
    return sd


_HF_MASK_BUFFERS = None


def hf_expects_mask_buffers() -> bool:
    """Does the INSTALLED transformers keep GPT-2's causal-mask buffers (`attn.bias`, `attn.masked_bias`) in state_dict?
    transformers 4.24 - the reference's pin, requirments.txt:12 - does (persistent buffers), current releases do not.  The
    reference loads checkpoints with a strict `model.load_state_dict(torch.load(...))` (predictions_runner.py:461), so
    a checkpoint must carry exactly the keys the environment's own GPT2LMHeadModel has.  Probed once on a tiny model;
    CAPDEC_CKPT_MASK_BUFFERS=1/0 overrides."""
    global _HF_MASK_BUFFERS
    import os
    env = os.environ.get("CAPDEC_CKPT_MASK_BUFFERS")
    if env in ("0", "1"):
        return env == "1"
    if _HF_MASK_BUFFERS is None:
        try:
            import transformers
            tiny = transformers.GPT2LMHeadModel(transformers.GPT2Config(n_layer=1, n_embd=8, n_head=1, vocab_size=8,
                                                                        n_positions=8))
            _HF_MASK_BUFFERS = "transformer.h.0.attn.bias" in tiny.state_dict()
        except Exception:
            _HF_MASK_BUFFERS = False
    return _HF_MASK_BUFFERS


class _Output(SimpleNamespace):
    """Stand-in for transformers' CausalLMOutput: `.logits`, `.loss`."""


class ClipCaptionModel(nn.Module):
    """train.py:246-273 / gpt2_prefix.py:139-171.  Accepts both constructor spellings (`prefix_size` / `prefix_dim`)."""

    def __init__(self, prefix_length: int, clip_length: Optional[int] = None, prefix_size: int = 512,
                 num_layers: int = 8, mapping_type: MappingType = MappingType.MLP, prefix_dim: Optional[int] = None,
                 gpt_config: Optional[GPT2Config] = None, pretrained: Optional[bool] = None):
        """`gpt_config=None` (the reference's call, train.py:262-273): GPT-2 starts from the pretrained 'gpt2' checkpoint
        like `GPT2LMHeadModel.from_pretrained('gpt2')` (train.py:266); if that checkpoint cannot be loaded (offline, no
        cache) the model falls back to HF-style random init AND SAYS SO on stderr.  Passing a `gpt_config` (tests, the
        synthetic benchmark) or `pretrained=False` asks for a randomly initialised GPT-2 of that architecture explicitly."""
        super().__init__()
        import weakref
        if prefix_dim is not None:
            prefix_size = prefix_dim
        self.prefix_length = prefix_length
        self.prefix_size = prefix_size
        self.mapping_type = MappingType.parse(mapping_type)
        if pretrained is None:
            pretrained = gpt_config is None and __import__("os").environ.get("CAPDEC_GPT2_PRETRAINED", "1") != "0"
        self.gpt = GPT2LMHead(gpt_config)
        self.gpt_init = "random (HF-style init, requested)"
        if pretrained:
            try:
                self.gpt.load_state_dict(_pretrained_gpt2_state_dict("gpt2"), strict=True)
                self.gpt_init = "pretrained 'gpt2' (GPT2LMHeadModel.from_pretrained)"
            except Exception as ex:   # offline box / empty cache: keep running, but never silently
                import sys
                self.gpt_init = "random (HF-style init): pretrained 'gpt2' unavailable"
                print(f"capdec_b200: WARNING - GPT2LMHeadModel.from_pretrained('gpt2') failed ({type(ex).__name__}: "
                      f"{str(ex)[:160]}); GPT-2 starts from RANDOM weights. Load a checkpoint (--pretrain_weights / "
                      f"load_state_dict) or make the 'gpt2' files available to train the reference's model.", file=sys.stderr)
        self.gpt_embedding_size = self.gpt.transformer.wte.weight.shape[1]
        d = self.gpt_embedding_size
        if self.mapping_type == MappingType.MLP:
            self.clip_project = MLP((prefix_size, (d * prefix_length) // 2, d * prefix_length))
        else:
            if clip_length is None:
                clip_length = prefix_length           # gpt2_prefix.py:160 (train.py:272 would fail on None)
            cls = TransformerMapper if self.mapping_type == MappingType.Transformer else TransformerEncoderDecoder
            self.clip_project = cls(prefix_size, d, prefix_length, clip_length, num_layers)
        ref = weakref.ref(self)
        self.gpt._owner = ref
        self.clip_project._owner = ref
        self._engine = None
        self._flat = None  # (flat_params, layout) once on a CUDA device

    # ---- reference-visible behaviour -----------------------------------------------------------------------
    def get_dummy_token(self, batch_size: int, device: torch.device) -> torch.Tensor:
        return torch.zeros(batch_size, self.prefix_length, dtype=torch.int64, device=device)

    def gpt_trainable(self) -> bool:
        return True

    def forward(self, tokens: torch.Tensor, prefix: torch.Tensor, mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None):
        """train.py:251-260 -> object with `.logits` [B, P+L, V] (autograd-connected to the parameters)."""
        eng = self.engine()
        logits = eng.logits_autograd(tokens, prefix, mask)
        out = _Output(logits=logits, loss=None)
        if labels is not None:  # train.py:256-259: labels = cat(zeros, tokens); HF shifted CE (ignore_index=-100)
            lab = torch.cat((self.get_dummy_token(tokens.shape[0], tokens.device), tokens), dim=1)
            out.loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]),
                                                         lab[:, 1:].reshape(-1))
        return out

    # ---- engine plumbing -----------------------------------------------------------------------------------
    def engine(self):
        from .engine import Engine
        if self._engine is None:
            dev = next(super().parameters()).device
            if dev.type != "cuda":
                raise CapdecError("capdec_b200 has no CPU path: move the model to a CUDA device first (model.to('cuda'))")
            self._flatten()
            self._engine = Engine(self)
        return self._engine

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None
        self._flat = None
        return out

    def _all_named_params(self):
        """Mapper parameters first, then GPT-2 (so `--only_prefix` trains a contiguous prefix of the flat buffer)."""
        mapper = [(n, p) for n, p in nn.Module.named_parameters(self) if n.startswith("clip_project.")]
        gpt = [(n, p) for n, p in nn.Module.named_parameters(self) if n.startswith("gpt.")]
        return mapper, gpt

    def _flatten(self):
        """Re-home every parameter as a view of one flat fp32 buffer: [tail(4) | mapper | gpt]."""
        if self._flat is not None:
            return self._flat
        mapper, gpt = self._all_named_params()
        dev = mapper[0][1].device
        n_map = sum(p.numel() for _, p in mapper)
        n_gpt = sum(p.numel() for _, p in gpt)
        for _, p in mapper + gpt:
            if p.dtype != torch.float32 or p.numel() % 4:
                raise CapdecError("parameters must be fp32 with sizes that are multiples of 4")
        tail = 4
        flat = torch.zeros(tail + n_map + n_gpt, dtype=torch.float32, device=dev)
        layout = {}
        off = tail
        with torch.no_grad():
            for n, p in mapper + gpt:
                v = flat[off: off + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                layout[n] = (off, p.numel(), tuple(p.shape))
                off += p.numel()
        self._flat = SimpleNamespace(params=flat, layout=layout, tail=tail, n_mapper=n_map, n_gpt=n_gpt, grads=None)
        return self._flat

    def _gpt_forward(self, inputs_embeds, attention_mask, labels):
        eng = self.engine()
        logits = eng.gpt_logits_from_embeds(inputs_embeds, attention_mask)
        out = _Output(logits=logits, loss=None)
        if labels is not None:
            out.loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(),
                                                         labels[:, 1:].reshape(-1))
        return out

    def _mapper_forward(self, x):
        return self.engine().mapper_infer(x)

    # ---- checkpoint compatibility ---------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts checkpoints written under transformers 4.24 (extra `attn.bias` / `attn.masked_bias` buffers)."""
        sd = {k: v for k, v in state_dict.items()
              if not (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"))}
        return super().load_state_dict(sd, strict=strict, **kw)

    def state_dict(self, *args, **kw):
        """The reference's checkpoint layout (train.py:359-371).  When the installed transformers keeps GPT-2's causal-mask
        buffers in its state_dict (4.24, the reference's pin), they are emitted too, so that `torch.save(model.state_dict())`
        - in the reference's own train() or in capdec_b200.fit.train - loads with strict=True in that environment."""
        sd = super().state_dict(*args, **kw)
        if hf_expects_mask_buffers():
            self._add_mask_buffers(sd, kw.get("prefix", args[1] if len(args) > 1 else ""))
        return sd

    def _add_mask_buffers(self, sd, prefix=""):
        n_pos = self.gpt.config.n_positions
        tril = torch.tril(torch.ones(n_pos, n_pos, dtype=torch.uint8)).view(1, 1, n_pos, n_pos)
        for i in range(self.gpt.config.n_layer):
            sd[f"{prefix}gpt.transformer.h.{i}.attn.bias"] = tril       # shared storage: HF never writes to it
            sd[f"{prefix}gpt.transformer.h.{i}.attn.masked_bias"] = torch.tensor(-1e4)
        return sd

    def state_dict_hf424(self):
        """state_dict plus the causal-mask buffers an unmodified transformers-4.24 strict loader expects, whatever
        transformers is installed here."""
        sd = nn.Module.state_dict(self)
        return self._add_mask_buffers(sd)


class ClipCaptionPrefix(ClipCaptionModel):
    """train.py:276-284: only the mapper trains; GPT-2 stays in eval mode (no dropout) and gets no weight gradients."""

    def parameters(self, recurse: bool = True):
        return self.clip_project.parameters()

    def train(self, mode: bool = True):
        super().train(mode)
        self.gpt.eval()
        return self

    def gpt_trainable(self) -> bool:
        return False
