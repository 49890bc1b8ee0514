"""Forward / backward orchestration of the CapDec train step over the C-ABI kernels (host plumbing, no arithmetic).

One `Engine` per model.  Activations live in a preallocated arena keyed by (B, P, L); every launch goes to torch's
current stream, so a whole step is CUDA-graph capturable.  Two entry points share the same kernels:

  * `loss_and_grads(tokens, prefix)` — the fast path used by `Trainer`: LM head only on the L consumed positions
    (train.py:349), fused CE fwd+bwd, gradients written straight into the flat gradient buffer;
  * `logits_autograd(tokens, prefix, mask)` — the drop-in path behind `ClipCaptionModel.forward` (train.py:251-260):
    materialises `.logits` [B, P+L, V] and hooks the hand-written backward into torch autograd so the reference's
    `loss.backward()` (train.py:351) works unmodified.

Reference map: embedding assembly train.py:253-255; GPT-2 block HF:modeling_gpt2.py:262-309; attention :54-72,185-226;
MLP :229-243; ln_f/lm_head :628,703-706; mapper MLP train.py:106-118; TransformerMapper train.py:229-243.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace
from typing import Optional

import torch

from . import ops
from ._lib import CapdecError

# dropout site ids (Philox stream): 1 = embeddings, 16 + 4*layer + {0: attention probs, 1: attn resid, 2: mlp resid}
_SITE_EMBD = 1


def _site(layer: int, k: int) -> int:
    return 16 + 4 * layer + k


class Engine:
    def __init__(self, model):
        self.m = model
        self.flat = model._flatten()
        self.dev = self.flat.params.device
        cfg = model.gpt.config
        self.cfg = cfg
        self.d = cfg.n_embd
        self.H = cfg.n_head
        self.hd = self.d // self.H
        self.F = 4 * self.d
        self.V = cfg.vocab_size
        self.Vp = (self.V + 127) // 128 * 128  # logits pitch (16-byte aligned rows for TMA)
        self.nl = cfg.n_layer
        self.P = model.prefix_length
        self.D = model.prefix_size
        self.is_mlp = model.mapping_type.value == "mlp"
        self.is_encdec = model.mapping_type.value == "transformer_decoder"
        if self.d % 128 or self.hd not in (64, 96):
            raise CapdecError(f"unsupported GPT-2 geometry d={self.d} heads={self.H}")
        self._params()
        self.seed = ops.make_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, self.dev)
        self.arenas = {}
        self._mapper_arenas = {}
        # packed-row execution of the fused train step (csrc/packed.cu): on unless CAPDEC_PACKED=0
        self.packed = os.environ.get("CAPDEC_PACKED", "1") != "0" and self.P >= 1
        # expected live rows of a packed batch (trunk rows / non-ignored targets): the GEMM planner sizes its tiles for
        # them (capdec_gemm_set_row_hint).  0 = plan for the dense extent.  `measure_row_hints` fills them in.
        self.hint_rows = 0
        self.hint_targets = 0
        self.lm_dgrad_splitk = os.environ.get("CAPDEC_LM_DGRAD_SPLITK", "1") != "0"
        # weight-gradient GEMMs of the GPT-2 blocks on a second stream (they are off the backward's critical path): their
        # CTAs fill the SMs a dgrad GEMM leaves idle in its last, partly empty wave (18.38 -> 18.20 / 18.23 ms/step on one
        # box, gpurun_out/s13_bench_*.log).  CAPDEC_BWD_STREAMS=0 serialises.
        self.bwd_streams = os.environ.get("CAPDEC_BWD_STREAMS", "1") != "0"
        # N = 768 GEMMs with a linear epilogue (attn / mlp c_proj forward, the three plain dgrads): 105-150 tiles of 256 x 256
        # on 74 CTA pairs are 1.4-2.03 waves that run as 2-3 (profiles/r2_gemm_waves.md: the raw tcgen05 issue rate of
        # fc_proj is 586 TF/s against 763 for fc for this reason alone).  Their outputs are zeroed and accumulated instead
        # (TMA reduce-add), which lets the planner split the reduction in two and fill the last wave.  tcgen05 modes only.
        # Measured (profiles/r2_gemm_waves.md): the mlp c_proj forward gains 14 us, but reduce-add (a read-modify-write in L2)
        # and the 39 MB memset cost the other four as much: 16.82-16.95 ms/step with it, 16.86-16.94 without.  Opt-in.
        self.splitk_linear = os.environ.get("CAPDEC_SPLITK_LINEAR", "0") == "1"
        self.serial_backward = False          # set while GEMM plans are being measured (Trainer.autotune)
        self._side = None

    # ------------------------------------------------------------------------------------------------------------
    # parameter / gradient views
    # ------------------------------------------------------------------------------------------------------------
    def _params(self):
        fl = self.flat
        self.p = {n: fl.params[o: o + k].view(s) for n, (o, k, s) in fl.layout.items()}
        if fl.grads is None:
            fl.grads = torch.zeros_like(fl.params)
        self.g = {n: fl.grads[o: o + k].view(s) for n, (o, k, s) in fl.layout.items()}
        if self.is_encdec:
            cp = self.m.clip_project
            self.C = cp.clip_length
            self.mH = cp.num_heads
            self.de = cp.dim_ref
            self.n_enc_layers = len(cp.ref_encoder.layers)
            self.n_dec_layers = len(cp.prefix_decoder.layers)
        elif not self.is_mlp:
            # to_queries [d,d] and to_keys_values [2d,d] are adjacent in the flat buffer -> one fused [3d,d] weight
            self.wqkv, self.gqkv = [], []
            j = 0
            while f"clip_project.transformer.layers.{j}.attn.to_queries.weight" in fl.layout:
                oq, kq, _ = fl.layout[f"clip_project.transformer.layers.{j}.attn.to_queries.weight"]
                ok, kk, _ = fl.layout[f"clip_project.transformer.layers.{j}.attn.to_keys_values.weight"]
                assert ok == oq + kq, "mapper q / kv weights must be adjacent in the flat buffer"
                self.wqkv.append(fl.params[oq: oq + kq + kk].view(3 * self.d, self.d))
                self.gqkv.append(fl.grads[oq: oq + kq + kk].view(3 * self.d, self.d))
                j += 1
            self.n_mapper_layers = j
            self.C = self.m.clip_project.clip_length
            self.mH = self.m.clip_project.num_heads
            self.mhd = self.d // self.mH

    def grad_views(self):
        return self.g

    def layer_grad_slices(self):
        """Flat-gradient slices in the order backward completes them: block L-1 (+ ln_f) ... block 0, then the rest
        ([tail | mapper | wte | wpe], final only after the embedding scatter).  Used to overlap the data-parallel
        all-reduce with the backward pass."""
        fl = self.flat
        lo = lambda n: fl.layout[n][0]
        slices = []
        for l in range(self.nl):
            start = lo(f"gpt.transformer.h.{l}.ln_1.weight")
            end = lo(f"gpt.transformer.h.{l + 1}.ln_1.weight") if l + 1 < self.nl else fl.grads.numel()
            slices.append(fl.grads[start:end])
        head = fl.grads[: lo("gpt.transformer.h.0.ln_1.weight")]
        return slices, head

    def zero_grads(self, mapper_only: bool = False):
        fl = self.flat
        end = fl.tail + fl.n_mapper if mapper_only else fl.grads.numel()
        fl.grads[:end].zero_()

    # ------------------------------------------------------------------------------------------------------------
    # arenas
    # ------------------------------------------------------------------------------------------------------------
    def _arena(self, B: int, L: int, P: Optional[int] = None):
        P = self.P if P is None else P
        key = (B, P, L)
        a = self.arenas.get(key)
        if a is not None:
            return a
        T = P + L
        if T > 128:
            raise CapdecError(f"sequence of {T} positions exceeds the single-tile attention limit (128)")
        M = B * T
        d, F = self.d, self.F
        # zero-initialised: the packed path lets GEMM tiles run over (finite) stale rows past the live row count
        e = lambda *s: torch.zeros(*s, device=self.dev, dtype=torch.float32)
        a = SimpleNamespace(B=B, P=P, L=L, T=T, M=M, packed=False)
        a.cu = torch.zeros(B + 1, device=self.dev, dtype=torch.int32)       # first packed row of each caption
        a.rows = torch.zeros(2, device=self.dev, dtype=torch.int32)         # {live rows, rounded up to 32}
        a.row_bt = torch.zeros(M, device=self.dev, dtype=torch.int32)       # packed row -> (b << 8) | t
        a.h = [e(M, d) for _ in range(self.nl + 1)]       # residual stream entering layer l (h[nl] = input of ln_f)
        a.x1 = [e(M, d) for _ in range(self.nl)]
        a.st1 = [e(M, 2) for _ in range(self.nl)]
        a.qkv = [e(M, 3 * d) for _ in range(self.nl)]
        a.lse = [e(B * self.H * T) for _ in range(self.nl)]
        a.ctx = [e(M, d) for _ in range(self.nl)]
        a.h1 = [e(M, d) for _ in range(self.nl)]
        a.x2 = [e(M, d) for _ in range(self.nl)]
        a.st2 = [e(M, 2) for _ in range(self.nl)]
        a.u = [e(M, F) for _ in range(self.nl)]
        a.g = [e(M, F) for _ in range(self.nl)]
        a.y = e(M, d)
        a.xf = e(M, d)
        a.stf = e(M, 2)
        a.pp = e(B, P * d) if P > 0 else None
        # backward scratch
        a.dh = e(M, d)
        a.dy = e(M, d)
        a.dx = e(M, d)
        a.dF = e(M, F)
        a.dqkv = e(M, 3 * d)
        a.dctx = e(M, d)
        a.dpp = e(B, P * d) if P > 0 else None
        # LM head on the consumed rows only (fast path)
        if L > 0:
            z = lambda *s_, dt=torch.float32: torch.zeros(*s_, device=self.dev, dtype=dt)
            a.xsel = z(B * L, d)
            a.dxsel = z(B * L, d)
            a.logits_sel = z(B * L, self.Vp)
            # compaction of the non-ignored target rows (train.py:350 ignore_index=0)
            a.row_src = z(B * L, dt=torch.int32)
            a.dst_of = z(M, dt=torch.int32)
            a.targets_c = z(B * L, dt=torch.int64)
            a.counts = z(2, dt=torch.int32)
        a.logits_full = None
        a.tokens = None
        self.arenas[key] = a
        return a

    def _mapper_arena(self, B: int):
        a = self._mapper_arenas.get(B)
        if a is not None:
            return a
        d, P = self.d, self.P
        e = lambda *s: torch.empty(*s, device=self.dev, dtype=torch.float32)
        a = SimpleNamespace(B=B)
        a.x_in = e(B, self.D)
        if self.is_mlp:
            hdim = (d * P) // 2
            a.a1 = e(B, hdim)
            a.da1 = e(B, hdim)
        elif self.is_encdec:
            C, de = self.C, self.de
            a.lin = e(B, C * de)
            mk = lambda rows, c, rows_kv: SimpleNamespace(
                y=e(rows, c), st=e(rows, 2), q=e(rows, c), kv=e(rows_kv, 2 * c), o=e(rows, c), t=e(rows, c),
                x1=e(rows, c), f=e(rows, 2 * c), xa=e(rows, c), xb=e(rows, c))
            a.enc = mk(B * C, de, B * C)
            a.dec = mk(B * P, d, B * max(C, P))
        else:
            C, S = self.C, self.C + P
            Mm = B * S
            nl = self.n_mapper_layers
            a.S, a.Mm = S, Mm
            a.lin = e(B, C * d)
            a.dlin = e(B, C * d)
            a.hm = [e(Mm, d) for _ in range(nl + 1)]   # residual entering layer j; hm[nl] = mapper output stream
            a.y1 = [e(Mm, d) for _ in range(nl)]
            a.s1 = [e(Mm, 2) for _ in range(nl)]
            a.qkv = [e(Mm, 3 * d) for _ in range(nl)]
            a.lse = [e(B * self.mH * S) for _ in range(nl)]
            a.o = [e(Mm, d) for _ in range(nl)]
            a.hm1 = [e(Mm, d) for _ in range(nl)]
            a.y2 = [e(Mm, d) for _ in range(nl)]
            a.s2 = [e(Mm, 2) for _ in range(nl)]
            a.f = [e(Mm, 2 * d) for _ in range(nl)]
            a.t = e(Mm, d)
            a.scratch = e(Mm, d)
            a.sscr = e(Mm, 2)
            a.dhm = e(Mm, d)
            a.dt = e(Mm, d)
            a.df = e(Mm, 2 * d)
            a.dqkv = e(Mm, 3 * d)
            a.do = e(Mm, d)
        self._mapper_arenas[B] = a
        return a

    # ------------------------------------------------------------------------------------------------------------
    # mapper
    # ------------------------------------------------------------------------------------------------------------
    def _mapper_fwd(self, x, pp):
        """x [B, D] -> pp [B, P*d] (train.py:254 `clip_project(prefix)`)."""
        B = x.shape[0]
        ma = self._mapper_arena(B)
        p = self.p
        if self.is_mlp:
            ops.linear_fwd(x, p["clip_project.model.0.weight"], "linear", p["clip_project.model.0.bias"], ma.a1,
                           act=ops.ACT_TANH)
            ops.linear_fwd(ma.a1, p["clip_project.model.2.weight"], "linear", p["clip_project.model.2.bias"], pp)
            return
        if self.is_encdec:
            return self._encdec_fwd(x, pp, ma)
        d, C, P, S, Mm = self.d, self.C, self.P, ma.S, ma.Mm
        ops.linear_fwd(x, p["clip_project.linear.weight"], "linear", p["clip_project.linear.bias"], ma.lin)
        ops.mapper_concat_fwd(ma.lin, p["clip_project.prefix_const"], ma.hm[0], B, C, P)
        nl = self.n_mapper_layers
        for j in range(nl):
            pre = f"clip_project.transformer.layers.{j}."
            if j == 0:
                ops.add_ln_fwd(ma.hm[0], None, None, ma.y1[0], ma.s1[0], p[pre + "norm1.weight"], p[pre + "norm1.bias"])
            ops.gemm(ma.y1[j], 0, self.wqkv[j], 0, ma.qkv[j], Mm, 3 * d, d)
            q, k, v = ma.qkv[j][:, :d], ma.qkv[j][:, d:2 * d], ma.qkv[j][:, 2 * d:]
            ops.attention_fwd(q, k, v, ma.o[j], ma.lse[j], B, self.mH, S, S, self.mhd, S * 3 * d, 3 * d, S * 3 * d, 3 * d,
                              S * d, d, self.mhd ** -0.5, 0)
            ops.linear_fwd(ma.o[j], p[pre + "attn.project.weight"], "linear", p[pre + "attn.project.bias"], ma.t)
            ops.add_ln_fwd(ma.hm[j], ma.t, ma.hm1[j], ma.y2[j], ma.s2[j], p[pre + "norm2.weight"], p[pre + "norm2.bias"])
            ops.linear_fwd(ma.y2[j], p[pre + "mlp.fc1.weight"], "linear", p[pre + "mlp.fc1.bias"], ma.f[j], act=ops.ACT_RELU)
            ops.linear_fwd(ma.f[j], p[pre + "mlp.fc2.weight"], "linear", p[pre + "mlp.fc2.bias"], ma.t)
            if j + 1 < nl:
                nx = f"clip_project.transformer.layers.{j + 1}."
                ops.add_ln_fwd(ma.hm1[j], ma.t, ma.hm[j + 1], ma.y1[j + 1], ma.s1[j + 1], p[nx + "norm1.weight"],
                               p[nx + "norm1.bias"])
            else:  # last residual add (no LayerNorm follows, train.py:178): reuse the fused kernel, discard its LN output
                ops.add_ln_fwd(ma.hm1[j], ma.t, ma.hm[nl], ma.scratch, ma.sscr, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        ops.rows_gather(ma.hm[nl], pp.view(B * P, d), B, S, P, C)  # out = transformer(prefix)[:, clip_length:]

    def _mapper_layer_fwd(self, pre, ws, x_in, x_out, B, Tq, c, kv_src=None, Tk=None, c_ref=None, sv=None):
        """One transformer_mapper.TransformerLayer (transformer_mapper.py:63-66): x_out = x1 + mlp(norm2(x1)),
        x1 = x_in + project(attention(q = to_queries(norm1(x_in)), k|v = to_keys_values(kv_src))).
        kv_src None -> norm1(x_in) (the layer's own normalised input), else a [B*Tk, c_ref] matrix.
        sv: per-layer buffers that keep what backward needs (training); None -> the shared scratch `ws` (inference)."""
        p = self.p
        H = self.mH
        hd = c // H
        Mq = B * Tq
        k_ = sv if sv is not None else ws              # where y1 / st1 / q / kv / o / x1 / f live
        y2, st2 = (sv.y2, sv.st2) if sv is not None else (ws.y, ws.st)
        ops.add_ln_fwd(x_in, None, None, k_.y, k_.st, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        ops.gemm(k_.y, 0, p[pre + "attn.to_queries.weight"], 0, k_.q, Mq, c, c)
        self_norm = kv_src is None
        if self_norm:
            kv_src, Tk, c_ref = k_.y, Tq, c
        Mk = B * Tk
        kv = k_.kv[:Mk]
        ops.gemm(kv_src, 0, p[pre + "attn.to_keys_values.weight"], 0, kv, Mk, 2 * c, c_ref)
        ops.attention_fwd(k_.q, kv[:, :c], kv[:, c:], k_.o, sv.lse if sv is not None else None, B, H, Tq, Tk, hd, Tq * c, c,
                          Tk * 2 * c, 2 * c, Tq * c, c, hd ** -0.5, 0)
        ops.linear_fwd(k_.o, p[pre + "attn.project.weight"], "linear", p[pre + "attn.project.bias"], ws.t)
        ops.add_ln_fwd(x_in, ws.t, k_.x1, y2, st2, p[pre + "norm2.weight"], p[pre + "norm2.bias"])
        ops.linear_fwd(y2, p[pre + "mlp.fc1.weight"], "linear", p[pre + "mlp.fc1.bias"], k_.f, act=ops.ACT_RELU)
        ops.linear_fwd(k_.f, p[pre + "mlp.fc2.weight"], "linear", p[pre + "mlp.fc2.bias"], ws.t)
        # last residual add: reuse the fused add + LayerNorm kernel and discard its LayerNorm output
        ops.add_ln_fwd(k_.x1, ws.t, x_out, ws.y, ws.st, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        if sv is not None:
            sv.x_in, sv.kv_src, sv.Tk, sv.c_ref, sv.self_norm = x_in, kv_src, Tk, c_ref, self_norm

    def _encdec_train_buffers(self, ma, B):
        """Per-layer activations of the TransformerDecoder mapper kept for backward, plus gradient scratch."""
        if getattr(ma, "sv_enc", None) is not None:
            return
        C, P, d, de = self.C, self.P, self.d, self.de
        e = lambda *s: torch.empty(*s, device=self.dev, dtype=torch.float32)

        def mk(rows, c, rows_kv, T):
            return SimpleNamespace(y=e(rows, c), st=e(rows, 2), q=e(rows, c), kv=e(rows_kv, 2 * c), o=e(rows, c), x1=e(rows, c),
                                   y2=e(rows, c), st2=e(rows, 2), f=e(rows, 2 * c), lse=e(B * self.mH * T), x_out=e(rows, c))
        ma.sv_enc = [mk(B * C, de, B * C, C) for _ in range(self.n_enc_layers)]
        ma.sv_dec = [mk(B * P, d, B * max(C, P), P) for _ in range(self.n_dec_layers)]
        ma.h0 = e(B * P, d)
        n = B * max(C, P) * max(d, de)
        ma.g = SimpleNamespace(cur=e(n), t=e(n), f=e(2 * n), q=e(n), kv=e(2 * n), o=e(n), ref=e(B * C, de), lin=e(B, C * de))

    def _encdec_fwd(self, x, pp, ma):
        """transformer_mapper.TransformerEncoderDecoder.forward (transformer_mapper.py:132-137).  Inference reuses one set
        of scratch buffers; with the mapper in train mode every layer keeps its activations for `_encdec_bwd`."""
        B = x.shape[0]
        p = self.p
        C, P, d, de = self.C, self.P, self.d, self.de
        train = bool(self.m.clip_project.training)
        if train:
            self._encdec_train_buffers(ma, B)
        ops.linear_fwd(x, p["clip_project.linear.weight"], "linear", p["clip_project.linear.bias"], ma.lin)
        cur = ma.lin.view(B * C, de)
        for j in range(self.n_enc_layers):          # ref_encoder: plain self-attention layers over the C tokens
            sv = ma.sv_enc[j] if train else None
            nxt = sv.x_out if train else (ma.enc.xb if cur is ma.enc.xa else ma.enc.xa)
            self._mapper_layer_fwd(f"clip_project.ref_encoder.layers.{j}.", ma.enc, cur, nxt, B, C, de, sv=sv)
            cur = nxt
        ref = cur
        h = ma.h0 if train else ma.dec.xa
        h.view(B, P, d).copy_(p["clip_project.prefix_const"].unsqueeze(0).expand(B, P, d))   # broadcast, no arithmetic
        for j in range(self.n_dec_layers):          # even: cross-attention to ref; odd: keys/values from the raw stream
            pre = f"clip_project.prefix_decoder.layers.{j}."
            sv = ma.sv_dec[j] if train else None
            hn = sv.x_out if train else (ma.dec.xb if h is ma.dec.xa else ma.dec.xa)
            if j % 2 == 0:
                self._mapper_layer_fwd(pre, ma.dec, h, hn, B, P, d, kv_src=ref, Tk=C, c_ref=de, sv=sv)
            else:
                self._mapper_layer_fwd(pre, ma.dec, h, hn, B, P, d, kv_src=h, Tk=P, c_ref=d, sv=sv)
            h = hn
        pp.view(B * P, d).copy_(h)
        ma.trained_fwd = train

    def _mapper_layer_bwd(self, pre, sv, dcur, gs, B, Tq, c, dref=None):
        """Backward of one TransformerLayer.  dcur [B*Tq, c]: gradient of the layer output on entry, of the layer input
        on exit.  Cross layers add the gradient of their key/value source into `dref`; layers whose keys/values come from
        the un-normalised stream add it into dcur; self layers into the gradient of norm1's output."""
        p, g = self.p, self.g
        H = self.mH
        hd = c // H
        Mq, Tk = B * Tq, sv.Tk
        Mk = B * Tk
        view = lambda t, rows, cols: t[: rows * cols].view(rows, cols)
        dt, df, dq, do = view(gs.t, Mq, c), view(gs.f, Mq, 2 * c), view(gs.q, Mq, c), view(gs.o, Mq, c)
        dkv = view(gs.kv, Mk, 2 * c)
        # ---- x_out = x1 + fc2(relu(fc1(norm2(x1)))) ----
        ops.linear_wgrad(sv.f, dcur, g[pre + "mlp.fc2.weight"], "linear", g[pre + "mlp.fc2.bias"])
        ops.linear_dgrad_act(dcur, p[pre + "mlp.fc2.weight"], "linear", df, sv.f, ops.ACT_RELU, dbias=g[pre + "mlp.fc1.bias"])
        ops.linear_wgrad(sv.y2, df, g[pre + "mlp.fc1.weight"], "linear")
        ops.linear_dgrad(df, p[pre + "mlp.fc1.weight"], "linear", dt)
        ops.add_ln_bwd(dt, sv.x1, sv.st2, p[pre + "norm2.weight"], dcur, dcur, None, g[pre + "norm2.weight"],
                       g[pre + "norm2.bias"], dbias_branch=g[pre + "attn.project.bias"])
        # ---- x1 = x_in + project(attention(...)) : dcur is d(x1) now ----
        ops.linear_wgrad(sv.o, dcur, g[pre + "attn.project.weight"], "linear")
        ops.linear_dgrad(dcur, p[pre + "attn.project.weight"], "linear", do)
        kv = sv.kv[:Mk]
        ops.attention_bwd(sv.q, kv[:, :c], kv[:, c:], sv.o, do, sv.lse, dq, dkv[:, :c], dkv[:, c:], B, H, Tq, Tk, hd, Tq * c, c,
                          Tk * 2 * c, 2 * c, Tq * c, c, hd ** -0.5, 0)
        ops.linear_wgrad(sv.y, dq, g[pre + "attn.to_queries.weight"], "linear")
        ops.linear_dgrad(dq, p[pre + "attn.to_queries.weight"], "linear", dt)                # d(norm1 output) from the queries
        ops.linear_wgrad(sv.kv_src, dkv, g[pre + "attn.to_keys_values.weight"], "linear")
        wkv = p[pre + "attn.to_keys_values.weight"]
        if sv.self_norm:                       # keys/values from norm1(x_in) as well
            ops.linear_dgrad(dkv, wkv, "linear", dt, accumulate=True)
        elif dref is not None:                 # cross-attention: keys/values from the encoder output
            ops.linear_dgrad(dkv, wkv, "linear", dref, accumulate=True)
        else:                                  # keys/values from the un-normalised stream entering this layer
            ops.linear_dgrad(dkv, wkv, "linear", dcur, accumulate=True)
        ops.add_ln_bwd(dt, sv.x_in, sv.st, p[pre + "norm1.weight"], dcur, dcur, None, g[pre + "norm1.weight"],
                       g[pre + "norm1.bias"])

    def _encdec_bwd(self, x, dpp, ma):
        """Backward of TransformerEncoderDecoder (what autograd derives from transformer_mapper.py:132-137 when
        gpt2_prefix.py:219-243 trains a MappingType.TransformerDecoder model)."""
        if not getattr(ma, "trained_fwd", False):
            raise CapdecError("TransformerDecoder mapper: backward needs a forward pass run in train mode")
        B = x.shape[0]
        g = self.g
        C, P, d, de = self.C, self.P, self.d, self.de
        gs = ma.g
        dcur = gs.cur[: B * P * d].view(B * P, d)
        dcur.copy_(dpp.view(B * P, d))
        dref = gs.ref
        ops.zero_fill(dref)
        for j in reversed(range(self.n_dec_layers)):
            self._mapper_layer_bwd(f"clip_project.prefix_decoder.layers.{j}.", ma.sv_dec[j], dcur, gs, B, P, d,
                                   dref=dref if j % 2 == 0 else None)
        ops.colsum_acc(dcur.view(B, P * d), g["clip_project.prefix_const"].view(-1))       # prefix_const is broadcast over B
        dcur = gs.cur[: B * C * de].view(B * C, de)
        dcur.copy_(dref)
        for j in reversed(range(self.n_enc_layers)):
            self._mapper_layer_bwd(f"clip_project.ref_encoder.layers.{j}.", ma.sv_enc[j], dcur, gs, B, C, de)
        gs.lin.view(B * C, de).copy_(dcur)
        ops.linear_wgrad(x, gs.lin, g["clip_project.linear.weight"], "linear", g["clip_project.linear.bias"])

    def _mapper_bwd(self, x, dpp):
        B = x.shape[0]
        ma = self._mapper_arena(B)
        if self.is_encdec:
            return self._encdec_bwd(x, dpp, ma)
        p, g = self.p, self.g
        if self.is_mlp:
            w2, w1 = "clip_project.model.2.", "clip_project.model.0."
            ops.linear_wgrad(ma.a1, dpp, g[w2 + "weight"], "linear", g[w2 + "bias"])
            ops.linear_dgrad_act(dpp, p[w2 + "weight"], "linear", ma.da1, ma.a1, ops.ACT_TANH, dbias=g[w1 + "bias"])
            ops.linear_wgrad(x, ma.da1, g[w1 + "weight"], "linear")
            return
        d, C, P, S, Mm = self.d, self.C, self.P, ma.S, ma.Mm
        nl = self.n_mapper_layers
        ops.rows_scatter(dpp.view(B * P, d), ma.dhm, B, S, P, C)
        for j in reversed(range(nl)):
            pre = f"clip_project.transformer.layers.{j}."
            # fc2 (its bias gradient comes fused from the add_ln_bwd that produced dhm; last layer: explicit colsum)
            ops.linear_wgrad(ma.f[j], ma.dhm, g[pre + "mlp.fc2.weight"], "linear",
                             g[pre + "mlp.fc2.bias"] if j == nl - 1 else None)
            ops.linear_dgrad_act(ma.dhm, p[pre + "mlp.fc2.weight"], "linear", ma.df, ma.f[j], ops.ACT_RELU,
                                 dbias=g[pre + "mlp.fc1.bias"])
            ops.linear_wgrad(ma.y2[j], ma.df, g[pre + "mlp.fc1.weight"], "linear")
            ops.linear_dgrad(ma.df, p[pre + "mlp.fc1.weight"], "linear", ma.dt)
            ops.add_ln_bwd(ma.dt, ma.hm1[j], ma.s2[j], p[pre + "norm2.weight"], ma.dhm, ma.dhm, None,
                           g[pre + "norm2.weight"], g[pre + "norm2.bias"], dbias_branch=g[pre + "attn.project.bias"])
            # attention branch
            ops.linear_wgrad(ma.o[j], ma.dhm, g[pre + "attn.project.weight"], "linear")
            ops.linear_dgrad(ma.dhm, p[pre + "attn.project.weight"], "linear", ma.do)
            q, k, v = ma.qkv[j][:, :d], ma.qkv[j][:, d:2 * d], ma.qkv[j][:, 2 * d:]
            dq, dk, dv = ma.dqkv[:, :d], ma.dqkv[:, d:2 * d], ma.dqkv[:, 2 * d:]
            ops.attention_bwd(q, k, v, ma.o[j], ma.do, ma.lse[j], dq, dk, dv, B, self.mH, S, S, self.mhd, S * 3 * d, 3 * d,
                              S * 3 * d, 3 * d, S * d, d, self.mhd ** -0.5, 0)
            ops.linear_wgrad(ma.y1[j], ma.dqkv, self.gqkv[j], "linear")
            ops.linear_dgrad(ma.dqkv, self.wqkv[j], "linear", ma.dt)
            prev_fc2_bias = g[f"clip_project.transformer.layers.{j - 1}.mlp.fc2.bias"] if j > 0 else None
            ops.add_ln_bwd(ma.dt, ma.hm[j], ma.s1[j], p[pre + "norm1.weight"], ma.dhm, ma.dhm, None,
                           g[pre + "norm1.weight"], g[pre + "norm1.bias"], dbias_branch=prev_fc2_bias)
        ops.mapper_concat_bwd(ma.dhm, ma.dlin, g["clip_project.prefix_const"], B, C, P)
        ops.linear_wgrad(x, ma.dlin, g["clip_project.linear.weight"], "linear", g["clip_project.linear.bias"])

    def mapper_infer(self, x):
        """`model.clip_project(prefix)` (predictions_runner.py:228, gpt2_prefix_eval.py:271) — forward only."""
        x = x.detach().to(device=self.dev, dtype=torch.float32).contiguous()
        B = x.shape[0]
        pp = torch.empty(B, self.P * self.d, device=self.dev)
        self._mapper_fwd(x, pp)
        return pp if self.is_mlp else pp.view(B, self.P, self.d)

    # ------------------------------------------------------------------------------------------------------------
    # GPT-2 trunk
    # ------------------------------------------------------------------------------------------------------------
    def _drop_p(self):
        """(embd, attn, resid) dropout probabilities in effect: GPT-2's `training` flag decides (train.py:281-284)."""
        if self.m.gpt.training:
            c = self.cfg
            return float(c.embd_pdrop), float(c.attn_pdrop), float(c.resid_pdrop)
        return 0.0, 0.0, 0.0

    def measure_row_hints(self, B: int, L: int) -> None:
        """Read the live-row / target counts of the batch that last ran through arena (B, L) (one host sync; the
        Trainer calls this once after its eager warm-up steps, before the step is captured in a CUDA graph)."""
        a = self.arenas.get((B, self.P, L))
        if a is None or not a.packed:
            return
        self.hint_rows = int(a.rows[0].item())
        self.hint_targets = int(a.counts[0].item())

    def _trunk_fwd(self, a, key_len=None):
        """a.h[0] (embeddings + positions) -> a.xf = ln_f(h_L).  HF:modeling_gpt2.py:612-628."""
        p = self.p
        B, T, M, d, F = a.B, a.T, a.M, self.d, self.F
        _, p_attn, p_res = a.pdrop
        R = a.rows[0:1] if a.packed else None     # live row count (device scalar) of a packed batch
        cu = a.cu if a.packed else None
        ops.add_ln_fwd(a.h[0], None, None, a.x1[0], a.st1[0], p["gpt.transformer.h.0.ln_1.weight"],
                       p["gpt.transformer.h.0.ln_1.bias"], eps=self.cfg.layer_norm_epsilon, rows=R)
        for l in range(self.nl):
            pre = f"gpt.transformer.h.{l}."
            ops.linear_fwd(a.x1[l], p[pre + "attn.c_attn.weight"], "conv1d", p[pre + "attn.c_attn.bias"], a.qkv[l], rows=R)
            q, k, v = a.qkv[l][:, :d], a.qkv[l][:, d:2 * d], a.qkv[l][:, 2 * d:]
            ops.attention_fwd(q, k, v, a.ctx[l], a.lse[l], B, self.H, T, T, self.hd, T * 3 * d, 3 * d, T * 3 * d, 3 * d,
                              T * d, d, self.hd ** -0.5, 1, key_len=key_len, p_drop=p_attn, seed=self.seed,
                              stream_id=_site(l, 0), cu_rows=cu)
            sk = self.splitk_linear and ops.is_tc()
            if sk:
                ops.zero_fill(a.y)
            ops.linear_fwd(a.ctx[l], p[pre + "attn.c_proj.weight"], "conv1d", p[pre + "attn.c_proj.bias"], a.y, rows=R,
                           accumulate=sk)
            ops.add_ln_fwd(a.h[l], a.y, a.h1[l], a.x2[l], a.st2[l], p[pre + "ln_2.weight"], p[pre + "ln_2.bias"],
                           eps=self.cfg.layer_norm_epsilon, p_drop=p_res, seed=self.seed, stream_id=_site(l, 1), rows=R)
            # tcgen05 modes: a.u holds gelu_new'(pre-activation) (one tanh serves both), fp32 verification mode: the pre-activation
            ops.linear_fwd(a.x2[l], p[pre + "mlp.c_fc.weight"], "conv1d", p[pre + "mlp.c_fc.bias"], a.g[l],
                           act=ops.ACT_GELU_NEW_D if ops.is_tc() else ops.ACT_GELU_NEW, aux=a.u[l], rows=R)
            if sk:
                ops.zero_fill(a.y)
            ops.linear_fwd(a.g[l], p[pre + "mlp.c_proj.weight"], "conv1d", p[pre + "mlp.c_proj.bias"], a.y, rows=R,
                           accumulate=sk)
            if l + 1 < self.nl:
                nx = f"gpt.transformer.h.{l + 1}."
                ops.add_ln_fwd(a.h1[l], a.y, a.h[l + 1], a.x1[l + 1], a.st1[l + 1], p[nx + "ln_1.weight"],
                               p[nx + "ln_1.bias"], eps=self.cfg.layer_norm_epsilon, p_drop=p_res, seed=self.seed,
                               stream_id=_site(l, 2), rows=R)
            else:
                ops.add_ln_fwd(a.h1[l], a.y, a.h[self.nl], a.xf, a.stf, p["gpt.transformer.ln_f.weight"],
                               p["gpt.transformer.ln_f.bias"], eps=self.cfg.layer_norm_epsilon, p_drop=p_res,
                               seed=self.seed, stream_id=_site(l, 2), rows=R)

    def _trunk_bwd(self, a, dxf, train_gpt: bool, key_len=None, on_layer_done=None):
        """dxf = dL/d(ln_f output) [M, d] -> a.dh = dL/d(h[0]); GPT-2 parameter gradients accumulated if train_gpt."""
        p, g = self.p, self.g
        B, T, M, d, F = a.B, a.T, a.M, self.d, self.F
        _, p_attn, p_res = a.pdrop
        gw = (lambda n: g[n]) if train_gpt else (lambda n: None)
        dy = a.dy if p_res > 0 else None     # branch-output gradient buffer (mask * dh); aliases dh when p = 0
        cur = (lambda: a.dy) if p_res > 0 else (lambda: a.dh)
        R = a.rows[0:1] if a.packed else None
        cu = a.cu if a.packed else None
        if a.packed:  # attention_bwd writes live rows only; the K-limited c_attn weight gradient reads whole k-blocks
            ops.zero_tail_rows(a.dqkv, a.rows)
        sk = self.splitk_linear and ops.is_tc()
        use_side = self.bwd_streams and train_gpt and not self.serial_backward
        if use_side and self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        side = self._side

        def wgrad(x, dy_, gname):
            """dW += x^T dy_.  Two-stream mode: launched on the side stream after everything issued so far on the main
            stream; returns the event that marks its completion (the caller waits on it before x / dy_ are overwritten)."""
            if not train_gpt:
                return None
            if not use_side:
                ops.linear_wgrad(x, dy_, g[gname], "conv1d", rows=R)
                return None
            ev = torch.cuda.Event()
            ev.record()
            side.wait_event(ev)
            with torch.cuda.stream(side):
                ops.linear_wgrad(x, dy_, g[gname], "conv1d", rows=R)
                done = torch.cuda.Event()
                done.record(side)
            return done

        def wait(ev):
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

        e_fc = e_qkv = None    # pending weight gradients that still read a.dF / a.dqkv
        # bias gradients of attn.c_proj / mlp.c_proj / c_fc / c_attn come fused out of add_ln_bwd / act_bwd / attention_bwd
        ops.add_ln_bwd(dxf, a.h[self.nl], a.stf, p["gpt.transformer.ln_f.weight"], None, a.dh, dy,
                       gw("gpt.transformer.ln_f.weight"), gw("gpt.transformer.ln_f.bias"), p_drop=p_res, seed=self.seed,
                       stream_id=_site(self.nl - 1, 2),
                       dbias_branch=gw(f"gpt.transformer.h.{self.nl - 1}.mlp.c_proj.bias"), rows=R)
        for l in reversed(range(self.nl)):
            pre = f"gpt.transformer.h.{l}."
            dy2 = cur()
            # mlp.c_proj
            e_cproj = wgrad(a.g[l], dy2, pre + "mlp.c_proj.weight")
            wait(e_fc)                                       # layer l+1's c_fc weight gradient still reads a.dF
            ops.linear_dgrad_act(dy2, p[pre + "mlp.c_proj.weight"], "conv1d", a.dF, a.u[l],
                                 ops.ACT_GELU_NEW_D if ops.is_tc() else ops.ACT_GELU_NEW,
                                 dbias=gw(pre + "mlp.c_fc.bias"), rows=R)
            e_fc = wgrad(a.x2[l], a.dF, pre + "mlp.c_fc.weight")
            if sk:
                ops.zero_fill(a.dx)
            ops.linear_dgrad(a.dF, p[pre + "mlp.c_fc.weight"], "conv1d", a.dx, rows=R, accumulate=sk)
            wait(e_cproj)                                    # the next kernel overwrites dy2
            ops.add_ln_bwd(a.dx, a.h1[l], a.st2[l], p[pre + "ln_2.weight"], a.dh, a.dh, dy, gw(pre + "ln_2.weight"),
                           gw(pre + "ln_2.bias"), p_drop=p_res, seed=self.seed, stream_id=_site(l, 1),
                           dbias_branch=gw(pre + "attn.c_proj.bias"), rows=R)
            dy1 = cur()
            # attention
            e_aproj = wgrad(a.ctx[l], dy1, pre + "attn.c_proj.weight")
            if sk:
                ops.zero_fill(a.dctx)
            ops.linear_dgrad(dy1, p[pre + "attn.c_proj.weight"], "conv1d", a.dctx, rows=R, accumulate=sk)
            q, k, v = a.qkv[l][:, :d], a.qkv[l][:, d:2 * d], a.qkv[l][:, 2 * d:]
            dq, dk, dv = a.dqkv[:, :d], a.dqkv[:, d:2 * d], a.dqkv[:, 2 * d:]
            wait(e_qkv)                                      # layer l+1's c_attn weight gradient still reads a.dqkv
            ops.attention_bwd(q, k, v, a.ctx[l], a.dctx, a.lse[l], dq, dk, dv, B, self.H, T, T, self.hd, T * 3 * d, 3 * d,
                              T * 3 * d, 3 * d, T * d, d, self.hd ** -0.5, 1, key_len=key_len, p_drop=p_attn,
                              seed=self.seed, stream_id=_site(l, 0), dbias_qkv=gw(pre + "attn.c_attn.bias"), cu_rows=cu)
            e_qkv = wgrad(a.x1[l], a.dqkv, pre + "attn.c_attn.weight")
            if sk:
                ops.zero_fill(a.dx)
            ops.linear_dgrad(a.dqkv, p[pre + "attn.c_attn.weight"], "conv1d", a.dx, rows=R, accumulate=sk)
            wait(e_aproj)                                    # the next kernel overwrites dy1
            if l > 0:
                ops.add_ln_bwd(a.dx, a.h[l], a.st1[l], p[pre + "ln_1.weight"], a.dh, a.dh, dy, gw(pre + "ln_1.weight"),
                               gw(pre + "ln_1.bias"), p_drop=p_res, seed=self.seed, stream_id=_site(l - 1, 2),
                               dbias_branch=gw(f"gpt.transformer.h.{l - 1}.mlp.c_proj.bias"), rows=R)
            else:
                ops.add_ln_bwd(a.dx, a.h[0], a.st1[0], p[pre + "ln_1.weight"], a.dh, a.dh, None, gw(pre + "ln_1.weight"),
                               gw(pre + "ln_1.bias"), rows=R)
            if on_layer_done is not None and train_gpt:
                if use_side:
                    wait(e_fc), wait(e_qkv)
                    e_fc = e_qkv = None   # joined: a hook that cuts the CUDA graph here must not leave events behind
                on_layer_done(l)   # every gradient of GPT-2 block l (and ln_f when l is the last block) is final now
        wait(e_fc), wait(e_qkv)     # join the side stream

    # ------------------------------------------------------------------------------------------------------------
    # fast path: loss + gradients (sum-reduced CE gradients, divided by the token count inside AdamW)
    # ------------------------------------------------------------------------------------------------------------
    def forward_hidden(self, tokens, prefix, key_len=None, packed=False):
        """tokens int64 [B,L], prefix fp32 [B,D] (already noise-injected) -> arena with a.xf = ln_f output.
        packed: keep only the positions that can reach the loss (csrc/packed.cu); a.xf then has a.rows[0] live rows."""
        B, L = tokens.shape
        a = self._arena(B, L)
        a.pdrop = self._drop_p()
        a.tokens, a.prefix = tokens, prefix
        a.packed = bool(packed)
        p = self.p
        self._mapper_fwd(prefix, a.pp)
        if a.packed:
            ops.pack_plan(tokens, self.P, a.cu, a.rows, a.row_bt)
            ops.embed_fwd_packed(tokens, a.pp, p["gpt.transformer.wte.weight"], p["gpt.transformer.wpe.weight"], a.h[0],
                                 a.row_bt, a.rows, B, self.P, L, p_drop=a.pdrop[0], seed=self.seed, stream_id=_SITE_EMBD)
        else:
            ops.embed_fwd(tokens, a.pp, p["gpt.transformer.wte.weight"], p["gpt.transformer.wpe.weight"], a.h[0], B, self.P,
                          L, p_drop=a.pdrop[0], seed=self.seed, stream_id=_SITE_EMBD)
        self._trunk_fwd(a, key_len)
        return a

    def backward_hidden(self, a, dxf, train_gpt: bool, key_len=None, on_layer_done=None):
        """Everything below ln_f: trunk, embedding scatter, mapper."""
        g = self.g
        self._trunk_bwd(a, dxf, train_gpt, key_len, on_layer_done)
        d_wte = g["gpt.transformer.wte.weight"] if train_gpt else None
        d_wpe = g["gpt.transformer.wpe.weight"] if train_gpt else None
        if a.packed:
            ops.embed_bwd_packed(a.tokens, a.dh, a.dpp, d_wte, d_wpe, a.cu, a.B, self.P, a.L, self.V, p_drop=a.pdrop[0],
                                 seed=self.seed, stream_id=_SITE_EMBD)
        else:
            ops.embed_bwd(a.tokens, a.dh, a.dpp, d_wte, d_wpe, a.B, self.P, a.L, self.V, p_drop=a.pdrop[0], seed=self.seed,
                          stream_id=_SITE_EMBD)
        self._mapper_bwd(a.prefix, a.dpp)

    def loss_and_grads(self, tokens, prefix, train_gpt: Optional[bool] = None, mean_reduce: bool = False,
                       on_layer_done=None, before_backward=None):
        """One forward+backward of train.py:348-351.  Gradients are ACCUMULATED into the flat gradient buffer,
        sum-reduced over tokens unless `mean_reduce` (then divided by the local count of non-ignored targets).
        tail[0] <- number of non-ignored targets, tail[1] <- sum of token losses.  Returns the tail view.
        `before_backward()` is called right before the first launch that writes a parameter gradient (the data-parallel
        trainer clears the gradient buffer on a side stream during the forward pass and joins it there)."""
        if train_gpt is None:
            train_gpt = self.m.gpt_trainable()
        fl = self.flat
        B, L = tokens.shape
        fast = ops.is_tc()
        use_packed = fast and self.packed
        ops.set_row_hint(self.hint_rows if use_packed else 0)
        a = self.forward_hidden(tokens, prefix, packed=use_packed)
        p, g = self.p, self.g
        P, T, d = self.P, a.T, self.d
        tail = fl.grads[:4]
        n_valid, loss_sum = tail[0:1], tail[1:2]
        targets = tokens.reshape(-1)
        wte = p["gpt.transformer.wte.weight"]
        logits = a.logits_sel[:, : self.V]
        if fast:
            # LM head + CE only over the rows whose target is not ignored: compact them on device, keep every shape
            # static and let the GEMMs / CE read the row count from a device scalar (CUDA-graph friendly).
            ops.compact_targets(targets, B, L, T, P - 1, a.row_src, a.dst_of, a.targets_c, a.counts, n_valid, loss_sum,
                                cu_rows=a.cu if a.packed else None)
            nv = a.counts[0:1]
            ops.set_row_hint(self.hint_targets)
            ops.rows_gather_idx(a.xf, a.xsel, a.row_src, a.counts)
            ops.gemm(a.xsel, 0, wte, 0, logits, B * L, self.V, d, m_limit=nv)                       # tied lm_head
            ops.ce_fwd_bwd(logits, a.targets_c, self.V, loss_sum, n_valid=n_valid if mean_reduce else None, row_limit=nv)
            if before_backward is not None:
                before_backward()
            if train_gpt:
                ops.gemm(logits, 1, a.xsel, 1, g["gpt.transformer.wte.weight"], self.V, d, B * L, accumulate=True,
                         k_limit=nv)
            # d(hidden) = dlogits . wte: a handful of 256-row tiles with a 50257-long reduction each -> split-K over the
            # vocabulary (reduce-add into a zeroed buffer) so that the work spreads over every SM whatever the row count
            # (reduce-add order is not deterministic; CAPDEC_LM_DGRAD_SPLITK=0 / eng.lm_dgrad_splitk = False restores
            # the single-pass, bit-reproducible product)
            if self.lm_dgrad_splitk:
                ops.zero_fill(a.dxsel)
            ops.gemm(logits, 0, wte, 1, a.dxsel, B * L, d, self.V, m_limit=nv, accumulate=self.lm_dgrad_splitk)
            ops.rows_scatter_idx(a.dxsel, a.dx, a.dst_of)
            ops.set_row_hint(self.hint_rows if use_packed else 0)
        else:
            ops.ce_count(targets, n_valid, loss_sum)
            ops.rows_gather(a.xf, a.xsel, B, T, L, P - 1)                      # hidden states of logits[:, P-1:-1]
            ops.linear_fwd(a.xsel, wte, "linear", None, logits)               # tied lm_head
            ops.ce_fwd_bwd(logits, targets, self.V, loss_sum, n_valid=n_valid if mean_reduce else None)
            if before_backward is not None:
                before_backward()
            if train_gpt:
                ops.linear_wgrad(a.xsel, logits, g["gpt.transformer.wte.weight"], "linear")
            ops.linear_dgrad(logits, wte, "linear", a.dxsel)
            ops.rows_scatter(a.dxsel, a.dx, B, T, L, P - 1)
        # a.dx is reused as scratch inside the trunk; ln_f backward consumes it first
        self.backward_hidden(a, a.dx, train_gpt, on_layer_done=on_layer_done)
        ops.set_row_hint(0)
        return tail

    def loss_only(self, tokens, prefix, stats):
        """Forward + masked CE of train.py:383-385 (the validation pass): no backward, no gradient buffers touched.
        stats[0] <- number of non-ignored targets, stats[1] <- sum of their token losses (fp32 device tensor, >= 2).
        Dropout follows the modules' train/eval flags exactly like the training path."""
        B, L = tokens.shape
        fast = ops.is_tc()
        use_packed = fast and self.packed
        ops.set_row_hint(self.hint_rows if use_packed else 0)
        a = self.forward_hidden(tokens, prefix, packed=use_packed)
        P, T, d = self.P, a.T, self.d
        n_valid, loss_sum = stats[0:1], stats[1:2]
        targets = tokens.reshape(-1)
        wte = self.p["gpt.transformer.wte.weight"]
        logits = a.logits_sel[:, : self.V]
        if fast:
            ops.compact_targets(targets, B, L, T, P - 1, a.row_src, a.dst_of, a.targets_c, a.counts, n_valid, loss_sum,
                                cu_rows=a.cu if a.packed else None)
            nv = a.counts[0:1]
            ops.set_row_hint(self.hint_targets)
            ops.rows_gather_idx(a.xf, a.xsel, a.row_src, a.counts)
            ops.gemm(a.xsel, 0, wte, 0, logits, B * L, self.V, d, m_limit=nv)
            ops.ce_fwd_bwd(logits, a.targets_c, self.V, loss_sum, write_grad=False, row_limit=nv)
        else:
            ops.ce_count(targets, n_valid, loss_sum)
            ops.rows_gather(a.xf, a.xsel, B, T, L, P - 1)
            ops.linear_fwd(a.xsel, wte, "linear", None, logits)
            ops.ce_fwd_bwd(logits, targets, self.V, loss_sum, write_grad=False)
        ops.set_row_hint(0)
        return stats

    # ------------------------------------------------------------------------------------------------------------
    # drop-in path: full logits + autograd hook
    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _key_len(mask, T):
        if mask is None:
            return None
        m = mask.reshape(mask.shape[0], -1)[:, :T] > 0
        klen = m.sum(dim=1).to(torch.int32)
        # only right padding can be expressed as a key length (which is what ClipCocoDataset produces, train.py:55-63);
        # CAPDEC_CHECK_MASK=1 verifies it (one host sync per call, so it is off on the hot path)
        if os.environ.get("CAPDEC_CHECK_MASK", "0") == "1":
            want = torch.arange(m.shape[1], device=m.device)[None, :] < klen[:, None]
            if not torch.equal(m, want):
                raise CapdecError("attention mask is not right-padded: only masks of the form [1...1 0...0] "
                                  "(ClipCocoDataset, train.py:55-63) are supported")
        return klen.contiguous()

    def _full_logits(self, a):
        if a.logits_full is None:
            a.logits_full = torch.empty(a.M, self.Vp, device=self.dev, dtype=torch.float32)
        lg = a.logits_full[:, : self.V]
        ops.linear_fwd(a.xf, self.p["gpt.transformer.wte.weight"], "linear", None, lg)
        return lg

    def logits_autograd(self, tokens, prefix, mask=None):
        if not tokens.is_cuda:
            raise CapdecError("capdec_b200 has no CPU path: tokens/prefix must be CUDA tensors")
        tokens = tokens.contiguous()
        prefix = prefix.detach().to(torch.float32).contiguous()
        params = [v for _, v in sorted(self._trainable().items())]
        return _LogitsFn.apply(self, tokens, prefix, mask, *params)

    def _trainable(self):
        named = dict(torch.nn.Module.named_parameters(self.m))
        if not self.m.gpt_trainable():
            named = {k: v for k, v in named.items() if k.startswith("clip_project.")}
        return {k: v for k, v in named.items() if v.requires_grad}

    def gpt_logits_from_embeds(self, inputs_embeds, attention_mask=None):
        """`model.gpt(inputs_embeds=...)` for decoding (gpt2_prefix_eval.py:76,163) — forward only, no grad."""
        x = inputs_embeds.detach().to(device=self.dev, dtype=torch.float32).contiguous()
        B, T, d = x.shape
        a = self._arena(B, 0, P=T)
        a.pdrop = self._drop_p()
        ops.embed_fwd(None, x, None, self.p["gpt.transformer.wpe.weight"], a.h[0], B, T, 0, p_drop=a.pdrop[0],
                      seed=self.seed, stream_id=_SITE_EMBD)
        self._trunk_fwd(a, self._key_len(attention_mask, T))
        return self._full_logits(a).view(B, T, self.V).clone()


class _LogitsFn(torch.autograd.Function):
    """`ClipCaptionModel.forward` for torch autograd: forward materialises logits [B,T,V]; backward runs the
    hand-written backward chain and hands each parameter its gradient (a view of the flat gradient buffer)."""

    @staticmethod
    def forward(ctx, eng: Engine, tokens, prefix, mask, *params):
        B, L = tokens.shape
        key_len = Engine._key_len(mask, eng.P + L)
        # fresh Philox masks for every forward with dropout live (train.py:348 runs GPT-2 in train mode: p = 0.1 at 37
        # sites); the seed this forward used is kept for ITS backward, whatever runs in between (validation, a second
        # forward): backward regenerates the masks from it instead of storing them
        ctx.seed = None
        if any(p > 0.0 for p in eng._drop_p()):
            ops.step_clock(seed=eng.seed)
            ctx.seed = eng.seed.clone()
        a = eng.forward_hidden(tokens, prefix, key_len)
        lg = eng._full_logits(a)
        ctx.eng, ctx.arena, ctx.key_len = eng, a, key_len
        ctx.names = sorted(eng._trainable().keys())
        return lg.view(B, a.T, eng.V)

    @staticmethod
    def backward(ctx, dlogits):
        eng, a = ctx.eng, ctx.arena
        train_gpt = eng.m.gpt_trainable()
        g, p = eng.g, eng.p
        # fresh gradients for this backward (autograd accumulates them into .grad itself)
        eng.zero_grads(mapper_only=not train_gpt)
        dl = a.logits_full[:, : eng.V]
        dl.copy_(dlogits.reshape(a.M, eng.V))
        if train_gpt:
            ops.linear_wgrad(a.xf, dl, g["gpt.transformer.wte.weight"], "linear")
        ops.linear_dgrad(dl, p["gpt.transformer.wte.weight"], "linear", a.dx)
        live_seed = eng.seed
        if ctx.seed is not None:
            eng.seed = ctx.seed            # the masks of THIS forward
        try:
            eng.backward_hidden(a, a.dx, train_gpt, ctx.key_len)
        finally:
            eng.seed = live_seed
        grads = [g[n].clone() for n in ctx.names]
        return (None, None, None, None, *grads)
