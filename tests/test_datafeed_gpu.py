"""Device-resident data feed + validation pass on the GPU.

DeviceCaptionDataset.batch() must return what the reference's ClipCocoDataset + DataLoader return for the same items
(train.py:52-72): token ids and mask bit-exact, prefix within fp32 round-off of `x / x.norm(2, -1)` (1e-6 relative).
Checked against the CPU oracle restatement and against the fixture recorded from the reference's own class
(tests/golden/datafeed.json).  Trainer.evaluate (train.py:372-389) is checked against the oracle's eval-mode loss at the
tf32 tolerance of tests/test_model_gpu.py (loss rel <= 2e-4).
"""
import json
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import capdec_oracle as O  # noqa: E402  (checker only)

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("case", [0, 1, 2])
def test_batch_gather_matches_reference_dataset(case):
    import capdec_b200 as cb
    c = json.loads((GOLD / "datafeed.json").read_text())["cases"][case]
    caps, cap2emb, table = O.make_caption_table(seed=c["table_seed"], n=c["n"], n_emb=c["n_emb"], half=c["half"])
    L, P = c["max_seq_len"], c["prefix_length"]
    ds = cb.DeviceCaptionDataset(caps, cap2emb, table, P, normalize_prefix=c["normalize_prefix"], max_seq_len=L)
    assert len(ds) == c["n"] and ds.max_seq_len == L and ds.table.dtype == table.dtype
    tokens, mask, prefix = ds.batch(c["idx"])
    assert tokens.dtype == torch.int64 and mask.dtype == torch.float32 and prefix.dtype == torch.float32
    assert tokens.cpu().tolist() == c["tokens"]                                  # the reference's own output
    assert mask.sum(1).cpu().tolist() == c["mask_sum_rows"]
    ref = [O.dataset_item(caps, cap2emb, table, it, L, P, c["normalize_prefix"]) for it in c["idx"]]
    assert torch.equal(tokens.cpu(), torch.stack([r[0] for r in ref]))
    assert torch.equal(mask.cpu(), torch.stack([r[1] for r in ref]))
    rp = torch.stack([r[2] for r in ref]).float()
    assert (prefix.cpu() - rp).abs().max() <= 1e-6 * rp.abs().max()
    assert (prefix.cpu()[:, :6].double() - torch.tensor(c["prefix_head"])).abs().max() < 1e-6
    # every item of the table, default max_seq_len rule (train.py:102-103)
    ds2 = cb.DeviceCaptionDataset(caps, cap2emb, table, P, normalize_prefix=c["normalize_prefix"])
    assert ds2.max_seq_len == O.dataset_max_seq_len(caps)
    t2, m2, _ = ds2.batch(list(range(c["n"])))
    for it in range(c["n"]):
        t, m, _ = O.dataset_item(caps, cap2emb, table, it, ds2.max_seq_len, P, False)
        assert torch.equal(t2[it].cpu(), t) and torch.equal(m2[it].cpu(), m)


def test_epoch_order_is_a_sharded_permutation_with_drop_last():
    import capdec_b200 as cb
    caps, cap2emb, table = O.make_caption_table(seed=1, n=103, n_emb=5)
    ds = cb.DeviceCaptionDataset(caps, cap2emb, table, 10)
    order = ds.epoch_order(8, generator=torch.Generator().manual_seed(3))
    assert tuple(order.shape) == (12, 8) and order.unique().numel() == 96 and order.is_cuda
    parts = [ds.epoch_order(8, generator=torch.Generator().manual_seed(3), rank=r, world=2) for r in range(2)]
    both = torch.cat([p.flatten() for p in parts])
    assert all(tuple(p.shape) == (6, 8) for p in parts) and both.unique().numel() == 96
    assert torch.equal(ds.epoch_order(8, shuffle=False).flatten().cpu(), torch.arange(96))


def test_train_from_device_dataset_and_validation_pass():
    import capdec_b200 as cb
    P, D, B = 10, 512, 6
    sd = O.make_state_dict(seed=4, mapping_type="mlp", prefix_length=P, prefix_size=D)
    cfg = cb.GPT2Config()
    model = cb.ClipCaptionModel(P, prefix_size=D, gpt_config=cfg)
    model.load_state_dict(sd)
    model = model.to("cuda").train()
    caps, cap2emb, table = O.make_caption_table(seed=9, n=40, n_emb=11, min_len=5, max_len=40)
    caps = [c.clamp_min(1) for c in caps]
    ds = cb.DeviceCaptionDataset(caps, cap2emb, table, P, normalize_prefix=True, max_seq_len=40)
    tr = cb.Trainer(model, batch_size=B, seq_len=40, noise_variance=0.016, lr=1e-4, warmup_steps=2, total_steps=50)
    idx = torch.tensor([5, 1, 39, 17, 8, 22], device="cuda")
    # ---- validation pass first (weights still those of `sd`): eval mode, no noise, no dropout ----
    val = []
    for rep in range(3):     # eager, then captured graph replays
        s = tr.evaluate_from(ds, idx).tolist()
        val.append(s[1] / s[0])
    ref_items = [O.dataset_item(caps, cap2emb, table, int(i), 40, P, True) for i in idx.tolist()]
    tokens = torch.stack([r[0] for r in ref_items])
    prefix = torch.stack([r[2] for r in ref_items])
    assert torch.equal(tr.tokens_d.cpu(), tokens)
    logits = O.clipcap_forward(sd, tokens, prefix, None, P)
    ref_loss = float(O.caption_loss(logits, tokens, P))
    assert s[0] == float((tokens != 0).sum())
    for v in val:
        assert abs(v - ref_loss) <= 2e-4 * abs(ref_loss), (val, ref_loss)
    assert abs(tr.evaluate(tokens.cuda(), prefix.cuda()) - ref_loss) <= 2e-4 * abs(ref_loss)
    assert model.training and all(p.grad is None or True for p in model.parameters())
    assert float(tr.eng.flat.grads.abs().max()) == 0.0          # the validation pass never touches gradient buffers
    # ---- a few train steps straight from the device table ----
    order = ds.epoch_order(B, generator=torch.Generator().manual_seed(0))
    losses = []
    for step in range(5):
        tr.step_from(ds, order[step])
        losses.append(tr.loss())
    assert all(torch.isfinite(torch.tensor(losses))) and all(9.0 < x < 13.0 for x in losses), losses
    after = tr.evaluate_from(ds, idx).tolist()
    assert after[0] == s[0] and abs(after[1] / after[0] - ref_loss) > 1e-6   # the weights moved
