"""CPU checks of capdec_b200.fit.train — the reference's `train()` loop (train.py:307-393) on the fast path.  The GPU pieces
(Trainer, DeviceCaptionDataset) are replaced by recording fakes, so what is tested here is the HOST logic the reference
defines: drop_last step count, scheduler horizon = epochs x steps, checkpoint names and cadence (train.py:358-371),
loss_per_epoch.json, the validation pass being built from --val_pt with the reference's constructor arguments
(train.py:373-375) and averaged per batch (train.py:386-388).  The real classes are exercised on the GPU by
tests/test_datafeed_gpu.py (step_from / evaluate_from) and tests/test_fit_gpu.py."""
import json
from types import SimpleNamespace

import pytest
import torch


class FakeModel:
    def __init__(self):
        self.moved, self.training, self.w = None, False, torch.zeros(3)

    def to(self, device):
        self.moved = device
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def state_dict(self):
        return {"w": self.w.clone()}


class FakeDeviceDataset:
    made = []

    def __init__(self, ref):
        self.ref = ref
        self.max_seq_len = ref.max_seq_len
        FakeDeviceDataset.made.append(self)

    @classmethod
    def from_reference(cls, ref, device="cuda"):
        return cls(ref)

    def __len__(self):
        return self.ref.n

    def epoch_order(self, batch_size, shuffle=True, generator=None, rank=0, world=1):
        per = (len(self) // (batch_size * world)) * batch_size
        perm = torch.randperm(len(self), generator=generator)
        perm = perm[: per * world].view(per, world)[:, rank] if world > 1 else perm[:per]
        return perm.view(-1, batch_size)


class FakeRefDataset:
    """Stands in for train.ClipCocoDataset: `type(dataset)(val_pt, prefix_length, normalize_prefix=..., ...)` builds the
    validation set exactly as train.py:373-374 does."""
    ctor_calls = []

    def __init__(self, path, prefix_length, normalize_prefix=False, use_image_embedding_as_clipcap=False, n=23, max_seq_len=17):
        FakeRefDataset.ctor_calls.append((path, prefix_length, normalize_prefix, use_image_embedding_as_clipcap))
        self.n, self.max_seq_len, self.prefix_length = (9 if path == "val.pkl" else n), (11 if path == "val.pkl" else max_seq_len), prefix_length


class FakeTrainer:
    last = None

    def __init__(self, model, **kw):
        self.model, self.kw, self.steps, self.evals = model, kw, [], []
        FakeTrainer.last = self

    def step_from(self, ds, idx):
        self.steps.append(idx.clone())
        k = len(self.steps)
        return torch.tensor([10.0, 10.0 * k, 0.0, 0.0])      # per-step mean loss = k

    def evaluate_any(self, ds, idx):
        self.evals.append((ds, idx.clone()))
        return torch.tensor([4.0, 8.0, 0.0, 0.0])            # batch mean 2.0


def _args(**kw):
    base = dict(bs=4, epochs=3, lr=2e-5, noise_variance=0.016, uniform_noise=False, dont_norm=False,
                add_modality_offset=False, save_every=2, val_pt="", prefix_length=10, use_image_embedding_as_clipcap=False)
    base.update(kw)
    return SimpleNamespace(**base)


def test_epoch_bookkeeping_checkpoints_and_schedule(tmp_path):
    from capdec_b200 import fit
    FakeDeviceDataset.made.clear()
    model = FakeModel()
    out = fit.train(FakeRefDataset("train.pkl", 10), model, _args(), warmup_steps=123, output_dir=str(tmp_path / "ckpt"),
                    output_prefix="coco", trainer_cls=FakeTrainer, device_dataset_cls=FakeDeviceDataset)
    tr = FakeTrainer.last
    assert out is model and model.training and model.moved.type == "cuda"
    # 23 captions, bs 4, drop_last -> 5 steps per epoch, 15 optimizer steps = the scheduler's horizon (train.py:327-330)
    assert len(tr.steps) == 15 and all(tuple(s.shape) == (4,) for s in tr.steps)
    assert tr.kw["total_steps"] == 15 and tr.kw["warmup_steps"] == 123 and tr.kw["seq_len"] == 17 and tr.kw["batch_size"] == 4
    assert tr.kw["noise_variance"] == pytest.approx(0.016) and tr.kw["lr"] == pytest.approx(2e-5)
    assert tr.kw["modality_offset"] is None and tr.kw["uniform_noise"] is False and tr.kw["dont_norm"] is False
    for e in range(3):   # every epoch visits 20 distinct captions; epochs are shuffled differently
        seen = torch.cat(tr.steps[5 * e: 5 * e + 5])
        assert seen.unique().numel() == 20
    assert not torch.equal(torch.cat(tr.steps[0:5]), torch.cat(tr.steps[5:10]))
    # checkpoints: epoch % save_every == 0 or last epoch (train.py:366-370) -> 000 and 002; no _latest before 10000 steps
    files = sorted(p.name for p in (tmp_path / "ckpt").iterdir())
    assert files == ["coco-000.pt", "coco-002.pt", "loss_per_epoch.json"]
    assert list(torch.load(tmp_path / "ckpt" / "coco-002.pt").keys()) == ["w"]
    rec = json.loads((tmp_path / "ckpt" / "loss_per_epoch.json").read_text())
    assert rec["val"] == [] and rec["train"] == pytest.approx([3.0, 8.0, 13.0])    # means of step losses 1..5, 6..10, 11..15


def test_validation_pass_follows_the_reference(tmp_path):
    from capdec_b200 import fit
    FakeDeviceDataset.made.clear(); FakeRefDataset.ctor_calls.clear()
    fit.train(FakeRefDataset("train.pkl", 10), FakeModel(), _args(epochs=2, save_every=1, val_pt="val.pkl", dont_norm=True),
              output_dir=str(tmp_path), output_prefix="p", trainer_cls=FakeTrainer, device_dataset_cls=FakeDeviceDataset)
    tr = FakeTrainer.last
    # train.py:373-374: a fresh ClipCocoDataset(args.val_pt, args.prefix_length, normalize_prefix=not args.dont_norm, ...) per epoch
    assert FakeRefDataset.ctor_calls[1:] == [("val.pkl", 10, False, False)] * 2
    # 9 validation captions, bs 4, drop_last -> 2 batches per epoch, each from the validation device dataset
    assert len(tr.evals) == 4 and all(ds.ref.n == 9 and tuple(idx.shape) == (4,) for ds, idx in tr.evals)
    rec = json.loads((tmp_path / "loss_per_epoch.json").read_text())
    assert rec["val"] == pytest.approx([2.0, 2.0]) and len(rec["train"]) == 2
    assert sorted(p.name for p in tmp_path.iterdir()) == ["loss_per_epoch.json", "p-000.pt", "p-001.pt"]
    assert tr.kw["dont_norm"] is True


def test_dataset_smaller_than_a_batch_is_an_error(tmp_path):
    from capdec_b200 import fit
    with pytest.raises(ValueError):
        fit.train(FakeRefDataset("train.pkl", 10, n=3), FakeModel(), _args(), output_dir=str(tmp_path),
                  trainer_cls=FakeTrainer, device_dataset_cls=FakeDeviceDataset)
