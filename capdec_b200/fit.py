"""The reference's training loop (`train()`, train.py:307-393) on the fast path.

Same signature, same side effects (epoch loop, `{prefix}-{epoch:03d}.pt` / `{prefix}_latest.pt` checkpoints with the
reference's state_dict layout, `loss_per_epoch.json`, the optional validation pass), but every batch runs as
`Trainer.step_from(DeviceCaptionDataset, idx)`: the batch is gathered on the device, the whole step of train.py:345-354
replays as CUDA graphs, and the loss is accumulated on the device instead of a `loss.item()` sync per step.

    launchers/run_train_b200.py --fast <train.py flags>      # rebinds train.train to this function

What differs from the reference, deliberately:
  * shuffling uses an explicit `torch.Generator` (seed = CAPDEC_SEED or 0, + epoch) instead of the global RNG, so that
    the ranks of a data-parallel run agree on the permutation; batches are drop_last like train.py:327;
  * under torchrun every rank trains on its shard of each batch and rank 0 alone writes files;
  * tqdm's per-step loss postfix is refreshed every `CAPDEC_LOG_EVERY` steps (default 50): each refresh is a device sync.
Host code only: tensor plumbing, file I/O and the epoch bookkeeping of the reference; no arithmetic of the hot path.
"""
from __future__ import annotations

import json
import os
import pickle
import sys

import torch


def _rank_world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def _validation_loss(tr, val_ds, batch_size: int, generator, rank: int, world: int) -> float:
    """train.py:372-389: mean over the validation batches (shuffled, drop_last) of the batch-mean token loss, model in
    eval mode, no noise injection.  The reference slices the logits with the TRAIN dataset's prefix_length (:384), which
    is the model's own here."""
    order = val_ds.epoch_order(batch_size, generator=generator, rank=rank, world=world)
    if order.shape[0] == 0:
        return float("nan")
    total = None
    for i in range(order.shape[0]):
        stats = tr.evaluate_any(val_ds, order[i])      # [n_valid, loss_sum] of the GLOBAL batch (summed over the ranks)
        batch_mean = (stats[1] / stats[0].clamp_min(1.0)).double()
        total = batch_mean if total is None else total + batch_mean
    return float(total) / int(order.shape[0])


def train(dataset, model, args, warmup_steps: int = 5000, output_dir: str = ".", output_prefix: str = "",
          trainer_cls=None, device_dataset_cls=None):
    """Drop-in for the reference's `train(dataset, model, args, warmup_steps, output_dir, output_prefix)`.
    `dataset` is a loaded reference `ClipCocoDataset` (or anything with its attributes: captions_tokens,
    caption2embedding, prefixes, prefix_length, normalize_prefix, max_seq_len).  Returns the model."""
    from .data import DeviceCaptionDataset
    from .trainer import Trainer
    trainer_cls = trainer_cls or Trainer
    device_dataset_cls = device_dataset_cls or DeviceCaptionDataset
    rank, world = _rank_world()
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    batch_size, epochs = int(args.bs), int(args.epochs)
    if rank == 0 and not os.path.exists(output_dir):
        os.makedirs(output_dir)
    model = model.to(device)
    model.train()
    ds = device_dataset_cls.from_reference(dataset, device=device)
    steps_per_epoch = len(ds) // (batch_size * world)                      # DataLoader(..., drop_last=True), train.py:327
    if steps_per_epoch == 0:
        raise ValueError(f"dataset of {len(ds)} captions is smaller than one global batch ({batch_size} x {world})")
    modality_offset = None
    if getattr(args, "add_modality_offset", False):                        # train.py:332-334
        with open("others/CLIP_embeddings_centers_info.pkl", "rb") as f:
            modality_offset = pickle.load(f)["offset_to_add_in_training"]
    tr = trainer_cls(model, batch_size=batch_size, seq_len=ds.max_seq_len, lr=float(args.lr), warmup_steps=warmup_steps,
                     total_steps=epochs * steps_per_epoch, noise_variance=float(args.noise_variance),
                     uniform_noise=bool(getattr(args, "uniform_noise", False)), dont_norm=bool(getattr(args, "dont_norm", False)),
                     modality_offset=modality_offset)
    seed = int(os.environ.get("CAPDEC_SEED", "0"))
    log_every = max(1, int(os.environ.get("CAPDEC_LOG_EVERY", "50")))
    try:
        from tqdm import tqdm
    except Exception:                                                      # pragma: no cover
        tqdm = None
    loss_per_epoch_train, loss_per_epoch_val = [], []
    for epoch in range(epochs):
        if rank == 0:
            print(f">>> Training epoch {epoch} / {epochs}")
            sys.stdout.flush()
        progress = tqdm(total=steps_per_epoch, desc=output_prefix) if (tqdm is not None and rank == 0) else None
        order = ds.epoch_order(batch_size, generator=torch.Generator().manual_seed(seed + epoch), rank=rank, world=world)
        accumulated = None                                                 # device scalar: sum of the per-step mean losses
        for idx in range(steps_per_epoch):
            stats = tr.step_from(ds, order[idx])                           # [n_valid, loss_sum, ., .] of the GLOBAL batch
            step_loss = stats[1] / stats[0].clamp_min(1.0)
            accumulated = step_loss.clone() if accumulated is None else accumulated + step_loss
            if progress is not None:
                if (idx + 1) % log_every == 0 or idx + 1 == steps_per_epoch:
                    progress.set_postfix({"loss": float(step_loss)})
                progress.update()
            if (idx + 1) % 10000 == 0 and rank == 0:                       # train.py:358-362
                torch.save(model.state_dict(), os.path.join(output_dir, f"{output_prefix}_latest.pt"))
        if progress is not None:
            progress.close()
        loss_per_epoch_train.append(float(accumulated) / steps_per_epoch)
        if rank == 0:
            print("loss_per_epoch_train: ", loss_per_epoch_train)
            if epoch % int(args.save_every) == 0 or epoch == epochs - 1:   # train.py:366-370
                torch.save(model.state_dict(), os.path.join(output_dir, f"{output_prefix}-{epoch:03d}.pt"))
        if getattr(args, "val_pt", ""):                                    # train.py:372-389
            val_ref = type(dataset)(args.val_pt, args.prefix_length, normalize_prefix=not args.dont_norm,
                                    use_image_embedding_as_clipcap=getattr(args, "use_image_embedding_as_clipcap", False))
            val_ds = device_dataset_cls.from_reference(val_ref, device=device)
            loss_per_epoch_val.append(_validation_loss(tr, val_ds, batch_size, torch.Generator().manual_seed(seed + 7919 + epoch),
                                                       rank, world))
            del val_ds
            if rank == 0:
                print("loss_per_epoch_val: ", loss_per_epoch_val)
        if rank == 0:
            with open(os.path.join(output_dir, "loss_per_epoch.json"), "w") as f:
                json.dump({"train": loss_per_epoch_train, "val": loss_per_epoch_val}, f)
    return model
