"""Device-resident data feed for the train step.

`DeviceCaptionDataset` holds what `ClipCocoDataset` (train.py:47-103) holds — the tokenised captions, the
caption -> CLIP-embedding index and the embedding table — but resident in HBM in a gather-friendly layout: captions
pre-padded / truncated to `max_seq_len` as one int32 matrix with -1 in the padding (train.py:52-59), the table in its
pickled dtype.  `gather(idx)` then produces exactly what `DataLoader(dataset)` + `.to(device)` would for those items
(train.py:60-72, :346) in ONE kernel launch: tokens int64 [B, L], mask fp32 [B, P+L], prefix fp32 [B, D] (normalised
when `normalize_prefix`).  This removes the per-sample Python `__getitem__`, the collate and the H2D copies from the
step (~10 us of GPU time instead of milliseconds of host time per batch at B200 step rates).

Host code here is tensor plumbing only (the reference's pickle / tokenizer loading stays the reference's).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from ._lib import CapdecError


def reference_max_seq_len(lengths: torch.Tensor) -> int:
    """train.py:102-103: min(int(mean + 10 * std), max) over the caption lengths (torch's unbiased std)."""
    all_len = lengths.float()
    return min(int(all_len.mean() + all_len.std() * 10), int(all_len.max()))


class DeviceCaptionDataset:
    def __init__(self, captions_tokens: Sequence[torch.Tensor], caption2embedding: Sequence[int], prefixes: torch.Tensor,
                 prefix_length: int, normalize_prefix: bool = False, max_seq_len: Optional[int] = None, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise CapdecError("DeviceCaptionDataset lives in GPU memory: capdec_b200 has no CPU path")
        n = len(captions_tokens)
        if n == 0 or len(caption2embedding) != n:
            raise ValueError("captions_tokens and caption2embedding must be non-empty and of equal length")
        lengths = torch.tensor([int(t.shape[0]) for t in captions_tokens])
        self.max_seq_len = int(max_seq_len) if max_seq_len is not None else reference_max_seq_len(lengths)
        self.prefix_length = int(prefix_length)
        self.normalize_prefix = bool(normalize_prefix)
        L = self.max_seq_len
        host = torch.full((n, L), -1, dtype=torch.int32)
        for i, t in enumerate(captions_tokens):               # one-off host prep (the reference pads lazily per item)
            k = min(int(t.shape[0]), L)
            host[i, :k] = t[:k].to(torch.int32)
        self.tokens_all = host.to(dev)
        self.cap2emb = torch.as_tensor(list(caption2embedding), dtype=torch.int32).to(dev)
        tbl = prefixes if isinstance(prefixes, torch.Tensor) else torch.stack(list(prefixes))
        if tbl.dtype not in (torch.float16, torch.float32):
            tbl = tbl.float()
        self.table = tbl.reshape(tbl.shape[0], -1).contiguous().to(dev)
        if int(self.cap2emb.max()) >= self.table.shape[0] or int(self.cap2emb.min()) < 0:
            raise ValueError("caption2embedding points outside the embedding table")
        self.device = dev

    @classmethod
    def from_reference(cls, ds, device="cuda"):
        """Build from a loaded reference `ClipCocoDataset` (train.py:47-103) without touching its files again."""
        return cls(ds.captions_tokens, ds.caption2embedding, ds.prefixes, ds.prefix_length,
                   normalize_prefix=ds.normalize_prefix, max_seq_len=ds.max_seq_len, device=device)

    def __len__(self) -> int:
        return int(self.tokens_all.shape[0])

    @property
    def prefix_size(self) -> int:
        return int(self.table.shape[1])

    def gather(self, idx: torch.Tensor, tokens: torch.Tensor, prefix: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """Fill caller-owned device buffers with the batch `idx` (int64 CUDA tensor [B])."""
        ops.batch_gather(self.tokens_all, self.cap2emb, self.table, idx, tokens, prefix, self.prefix_length, mask=mask,
                         normalize=self.normalize_prefix)

    def batch(self, idx):
        """(tokens, mask, prefix) for the items `idx` — the tuple the reference's DataLoader yields (train.py:345)."""
        idx = torch.as_tensor(idx, dtype=torch.int64).to(self.device).reshape(-1).contiguous()
        B, L, P = idx.numel(), self.max_seq_len, self.prefix_length
        tokens = torch.empty(B, L, dtype=torch.int64, device=self.device)
        mask = torch.empty(B, P + L, dtype=torch.float32, device=self.device)
        prefix = torch.empty(B, self.prefix_size, dtype=torch.float32, device=self.device)
        self.gather(idx, tokens, prefix, mask)
        return tokens, mask, prefix

    def epoch_order(self, batch_size: int, shuffle: bool = True, generator: Optional[torch.Generator] = None, rank: int = 0,
                    world: int = 1) -> torch.Tensor:
        """Device int64 [steps, batch_size]: a shuffled epoch with drop_last=True (train.py:327), sharded
        DistributedSampler-style when world > 1 (every rank must pass the same generator seed)."""
        n = len(self)
        perm = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        per = (n // (batch_size * world)) * batch_size
        perm = perm[: per * world].view(per, world)[:, rank] if world > 1 else perm[:per]
        return perm.view(-1, batch_size).to(self.device)
