"""CPU test of the data-parallel reduction semantics with gloo, world_size 2:
each rank holds half of the captions, produces SUM-reduced CE gradients plus [n_valid, loss_sum], one all-reduce
over the flat payload [n_valid, loss_sum, 0, 0 | grads], then divides by the GLOBAL count — this must equal the
single-process mean over the whole batch (train.py:350), also when ranks hold different numbers of valid tokens.
This is the host-side protocol `capdec_b200.Trainer` runs around the CUDA kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import capdec_oracle as O

P = 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat_sum_grads(sd, tokens, prefix):
    """[n_valid, loss_sum, 0, 0 | mapper grads] with sum-reduced token losses (what one rank contributes)."""
    leaves = {k: (v.detach().clone().requires_grad_(k.startswith("clip_project")) if k != "gpt.lm_head.weight" else None)
              for k, v in sd.items()}
    leaves["gpt.lm_head.weight"] = leaves["gpt.transformer.wte.weight"]
    logits = O.clipcap_forward(leaves, tokens, prefix, None, P)
    lg = logits[:, P - 1:-1]
    loss_sum = torch.nn.functional.cross_entropy(lg.reshape(-1, lg.shape[-1]), tokens.flatten(), ignore_index=0, reduction="sum")
    loss_sum.backward()
    names = sorted(k for k in leaves if k.startswith("clip_project"))
    tail = torch.tensor([float((tokens != 0).sum()), float(loss_sum), 0.0, 0.0])
    return torch.cat([tail] + [leaves[k].grad.flatten() for k in names]), names


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = O.make_state_dict(seed=1, n_layer=2)
    tokens, prefix, _ = O.make_batch(seed=9, B=4)
    tokens[0, 12:] = 0  # make the shards unbalanced in valid tokens
    sl = slice(rank * 2, rank * 2 + 2)
    flat, _ = _flat_sum_grads(sd, tokens[sl], prefix[sl])
    dist.all_reduce(flat)                      # the ONE collective of the step
    grads = flat[4:] / flat[0]                 # AdamW kernel: g / *grad_denom_dev
    if rank == 0:
        torch.save({"flat": flat, "grads": grads}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sum_reduce_equals_global_mean(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    sd = O.make_state_dict(seed=1, n_layer=2)
    tokens, prefix, _ = O.make_batch(seed=9, B=4)
    tokens[0, 12:] = 0
    trainable = lambda k: k.startswith("clip_project")
    loss, _, grads = O.loss_and_grads(sd, tokens, prefix, None, P, None, trainable)
    ref = torch.cat([grads[k].flatten() for k in sorted(grads)])
    assert got["flat"][0].item() == (tokens != 0).sum().item()
    assert got["flat"][1].item() / got["flat"][0].item() == pytest.approx(float(loss), rel=1e-5)
    assert (got["grads"] - ref).norm() / ref.norm() < 1e-5
    # averaging per-rank MEANS instead would be wrong here (unequal token counts) — guard the design choice
    n0, n1 = (tokens[:2] != 0).sum().item(), (tokens[2:] != 0).sum().item()
    assert n0 != n1


def test_push_plan_covers_every_gradient_exactly_once():
    """Host logic of the copy-engine gradient pushes (capdec_b200.trainer.push_plan): over all buckets of a step, every
    rank sends every element that another rank owns exactly once, into its own slot of the owner's staging area, and the
    slots of the world - 1 sources tile that area without overlap."""
    from capdec_b200.trainer import push_plan
    for world in (2, 3, 8):
        shard = 4 * 37
        n = world * shard
        spans = [(n - 50 - 40 * i, n - 50 - 40 * (i - 1)) for i in range(1, 6)]        # "layers", back to front
        spans = [(max(0, a), b) for a, b in spans] + [(0, max(0, n - 50 - 200))] + [(n - 50, n)]
        assert sorted(x for a, b in spans for x in range(a, b)) == list(range(n))
        for owner in range(world):
            filled = {}
            for rank in range(world):
                if rank == owner:
                    assert all(r != rank for sp in spans for (r, _, _, _) in push_plan(sp, world, rank, shard))
                    continue
                for sp in spans:
                    for r, first, count, off in push_plan(sp, world, rank, shard):
                        if r != owner:
                            continue
                        assert owner * shard <= first and first + count <= (owner + 1) * shard
                        for i in range(count):
                            key = off + i
                            assert key not in filled, "two pushes land on the same staging element"
                            filled[key] = (rank, first + i)
            assert sorted(filled) == list(range((world - 1) * shard))
            for key, (rank, elem) in filled.items():       # slot = source rank with the owner left out, element order kept
                slot = rank if rank < owner else rank - 1
                assert key == slot * shard + (elem - owner * shard)
