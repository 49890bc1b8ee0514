"""Parity at BASELINE.json scale (GPU): loss, n_valid and every per-tensor gradient of ONE train step at the C1 / C2 / C3
configurations against the CPU oracle (pinned bit-exactly on the reference's own classes at small batch,
tests/golden) run on this box's host cores on the same seeded weights and inputs.

  C1  MLP mapper P=10, GPT-2 frozen (--only_prefix), bs 32, noise_variance 0.016 (train.py:276-284, 345-351)
  C2  MLP mapper P=10 + GPT-2 fine-tuned, bs 256, caption lengths ~ U{8..40}: the benchmarked configuration - packed rows,
      device row limits, measured GEMM plans (Trainer.autotune), split-K LM-head dgrad, two-stream backward, >= 5 GEMM waves
  C3  TransformerMapper (8 layers, P = C = 40) + GPT-2 fine-tuned, bs 256

Dropout is off (p = 0; the oracle cannot reproduce Philox masks).
Stated tolerances: "tf32" (1xTF32, the perf mode) loss rel <= 2e-4, per-tensor grad rel-L2 <= 2e-2;
"tf32x3" (3xTF32, the fp32-grade mode) loss rel <= 2e-6, per-tensor grad rel-L2 <= 1e-4.
These hold for every WELL-CONDITIONED gradient.  tests/golden/conditioning.json (tests/golden/make_conditioning.py) holds,
per tensor, what fp32 rounding alone does to the reference algorithm (oracle fp32 vs fp64): 0.8-1.2e-6 for every tensor of
C1 / C2 and for all 148 GPT-2 tensors of C3, but 3e-6 ... 3e-4 (up to 300x the median) for 97 of the 99 tensors of C3's
randomly initialised TransformerMapper: ReLU gates of mlp.fc1 flip under 1e-7 perturbations and the near-uniform softmax
amplifies the change; the reference itself on another BLAS differs by that much.  Tensors whose fp32-vs-fp64 error exceeds
3e-6 (3x the median) are held to the ill-conditioned bound instead: 5e-2 ("tf32"), 5e-3 ("tf32x3").  Measured values:
gpurun_out/parity_report.jsonl -> profiles/r2_parity_report.md.
"""
import json
import os
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import capdec_oracle as O  # noqa: E402  (checker only)

ROOT = Path(__file__).resolve().parent.parent
TOL = {"tf32": dict(loss=2e-4, grad=2e-2, grad_ill=5e-2), "tf32x3": dict(loss=2e-6, grad=1e-4, grad_ill=5e-3)}
ILL_CONDITIONED_ABOVE = 3e-6      # fp32-vs-fp64 rel-L2 of the reference algorithm itself (median over tensors: 1e-6)
COND = json.loads((ROOT / "tests" / "golden" / "conditioning.json").read_text())
CASES = {
    "c1": dict(B=32, P=10, C=10, mapping="mlp", only_prefix=True, noise=0.016),
    "c2": dict(B=256, P=10, C=10, mapping="mlp", only_prefix=False, noise=0.016),
    "c3": dict(B=256, P=40, C=40, mapping="transformer", only_prefix=False, noise=0.016),
}
_ORACLE = {}


def _oracle(case):
    """Oracle loss / gradients of the case, computed once per session on the host cores (~10-25 s at bs 256)."""
    if case in _ORACLE:
        return _ORACLE[case]
    c = CASES[case]
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    sd = O.make_state_dict(seed=11, mapping_type=c["mapping"], prefix_length=c["P"], clip_length=c["C"], prefix_size=512,
                           num_layers=8)
    tokens, prefix, draw = O.make_batch(seed=12, B=c["B"], L=40, prefix_size=512)
    pfx = O.noise_injection(prefix, c["noise"], noise=draw)
    trainable = (lambda k: k.startswith("clip_project")) if c["only_prefix"] else None
    loss, _, grads = O.loss_and_grads(sd, tokens, pfx, O.make_mask(tokens, c["P"]), c["P"], c["C"], trainable)
    _ORACLE[case] = (sd, tokens, prefix, draw, pfx, float(loss), grads)
    return _ORACLE[case]


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("mode", ["tf32", "tf32x3"])
@pytest.mark.parametrize("case", ["c1", "c2", "c3"])
def test_step_at_baseline_scale_matches_oracle(case, mode):
    import capdec_b200 as cb
    c = CASES[case]
    sd, tokens, prefix, draw, pfx_ref, o_loss, o_grads = _oracle(case)
    cb.ops.set_precision(mode)
    try:
        cls = cb.ClipCaptionPrefix if c["only_prefix"] else cb.ClipCaptionModel
        mt = cb.MappingType.MLP if c["mapping"] == "mlp" else cb.MappingType.Transformer
        model = cls(c["P"], clip_length=c["C"], prefix_size=512, num_layers=8, mapping_type=mt,
                    gpt_config=cb.GPT2Config(resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0))
        model.load_state_dict(sd)
        model = model.to("cuda").train()
        eng = model.engine()
        tok_d = tokens.cuda()
        pfx = torch.empty(c["B"], 512, device="cuda")
        cb.ops.noise_injection(prefix.cuda(), pfx, c["noise"], noise=(draw * c["noise"] ** 0.5).cuda())
        assert (pfx.cpu() - pfx_ref).abs().max() < 1e-6
        # the benchmarked machinery: two eager Trainer steps (lr = 0: weights untouched) measure the live-row hints and the
        # GEMM plans of exactly these problems; the step under test then runs with those plans
        tr = cb.Trainer(model, batch_size=c["B"], seq_len=40, lr=0.0, warmup_steps=0, total_steps=10, noise_variance=0.0)
        for _ in range(2):
            tr.step(tok_d, pfx)
        torch.cuda.synchronize()
        assert eng.packed and eng.hint_rows > 0
        for k, v in torch.nn.Module.named_parameters(model):      # lr = 0 left every weight bit-identical
            assert torch.equal(v.detach().cpu(), sd[k]), k
        eng.zero_grads()
        tail = eng.loss_and_grads(tok_d, pfx, mean_reduce=True)
        torch.cuda.synchronize()
        n_valid, loss_sum = tail[0].item(), tail[1].item()
        assert n_valid == (tokens != 0).sum().item()
        loss = loss_sum / n_valid
        tol = TOL[mode]
        g = eng.grad_views()
        rels = {k: rel_l2(g[k].cpu(), og) for k, og in o_grads.items()}
        cond = COND[case]["fp32_vs_fp64_rel_l2"]
        ill = {k for k in rels if cond[k] > ILL_CONDITIONED_ABOVE}
        well = {k: v for k, v in rels.items() if k not in ill}
        worst_well = max(well, key=well.get)
        worst_ill = max(ill, key=rels.get) if ill else None
        srt = sorted(rels.values())
        rec = dict(test="scale_parity", case=case, mode=mode, B=c["B"], n_valid=n_valid, live_rows=eng.hint_rows, loss=loss,
                   loss_ref=o_loss, loss_rel=abs(loss - o_loss) / abs(o_loss), worst_grad_rel_l2=well[worst_well],
                   worst_grad=worst_well, median_grad_rel_l2=srt[len(srt) // 2], n_tensors=len(rels), n_ill_conditioned=len(ill),
                   worst_ill_conditioned_rel_l2=rels[worst_ill] if ill else None, worst_ill_conditioned=worst_ill,
                   worst_ill_conditioned_fp32_vs_fp64=cond[worst_ill] if ill else None)
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        with open(out / "parity_report.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
        (out / f"scale_parity_{case}_{mode}.json").write_text(json.dumps({k: [rels[k], cond[k]] for k in rels}, indent=0))
        assert abs(loss - o_loss) <= tol["loss"] * abs(o_loss), (loss, o_loss)
        assert well[worst_well] <= tol["grad"], (worst_well, well[worst_well])
        if ill:
            assert rels[worst_ill] <= tol["grad_ill"], (worst_ill, rels[worst_ill], cond[worst_ill])
        if c["only_prefix"]:  # frozen GPT-2: no gradient may have been written (train.py:276-284)
            fl = eng.flat
            assert fl.grads[fl.tail + fl.n_mapper:].abs().max().item() == 0.0
    finally:
        cb.ops.set_precision("tf32")
        cb.ops.gemm_autotune(-1)
