"""Real multi-GPU data-parallel test (NCCL over NVLink): 2 ranks, each with half of the batch, must end bit-identical
to each other and match one process stepping the whole batch ('real 2/4/8-GPU NCCL test').
Skipped unless the box has >= 2 GPUs (`gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _make(cb, O, sd, B, lr, warm, total, graph):
    cfg = cb.GPT2Config(resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    m = cb.ClipCaptionModel(10, prefix_size=512, gpt_config=cfg)
    m.load_state_dict(sd)
    m = m.to("cuda").train()
    return m, cb.Trainer(m, batch_size=B, seq_len=40, lr=lr, warmup_steps=warm, total_steps=total, noise_variance=0.0,
                         use_cuda_graph=graph)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import capdec_b200 as cb
    from oracle import capdec_oracle as O
    sd = O.make_state_dict(seed=1)
    tokens, prefix, _ = O.make_batch(seed=2, B=8)
    tokens[0, 10:] = 0                                   # unequal valid-token counts per rank
    per = 8 // world
    sl = slice(rank * per, (rank + 1) * per)
    m, tr = _make(cb, O, sd, per, 1e-3, 1, 10, True)
    losses = []
    for _ in range(4):                                   # 2 eager warm-up steps + 2 CUDA-graph replays
        tr.step(tokens[sl].cuda(), prefix[sl].cuda())
        losses.append(tr.loss())
    flat = tr.eng.flat.params.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"losses": losses, "identical": all(torch.equal(gathered[0], g) for g in gathered[1:]),
                    "params": flat.cpu(), "peer": tr.peer, "push": tr.push, "sharded": tr.sharded}, os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _check_against_single_gpu(tmp_path, expect_peer=None, expect_push=None):
    import torch.multiprocessing as mp
    import capdec_b200 as cb
    from oracle import capdec_oracle as O
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    assert got["identical"], "ranks diverged"
    if expect_peer is not None:
        assert got["peer"] == expect_peer, "the data-parallel update did not take the expected path"
    if expect_push is not None:
        assert got["push"] == expect_push, "the gradients did not travel the expected way"
    sd = O.make_state_dict(seed=1)
    tokens, prefix, _ = O.make_batch(seed=2, B=8)
    tokens[0, 10:] = 0
    m, tr = _make(cb, O, sd, 8, 1e-3, 1, 10, False)
    losses = []
    for _ in range(4):
        tr.step(tokens.cuda(), prefix.cuda())
        losses.append(tr.loss())
    # the all-reduced (n_valid, loss_sum) give the GLOBAL mean loss on every rank
    assert got["losses"] == pytest.approx(losses, rel=2e-4)
    ref = tr.eng.flat.params.detach().cpu()
    rel = ((got["params"] - ref).norm() / ref.norm()).item()
    assert rel < 2e-4, rel


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_data_parallel_matches_single_gpu(tmp_path):
    """Default path: copy-engine pushes of every finished gradient bucket into the owners' staging areas during the
    backward pass, then the sharded update as ONE kernel that all-gathers over NVLink peer memory (csrc/peer.cu)."""
    _check_against_single_gpu(tmp_path, expect_peer=True, expect_push=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_peer_update_with_nvlink_gradient_loads_matches_single_gpu(tmp_path, monkeypatch):
    """CAPDEC_DP_PUSH=0: the update kernel loads the peers' gradients over NVLink itself (no staging)."""
    monkeypatch.setenv("CAPDEC_DP_PUSH", "0")
    _check_against_single_gpu(tmp_path, expect_peer=True, expect_push=False)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_nccl_sharded_update_matches_single_gpu(tmp_path, monkeypatch):
    """CAPDEC_DP_PEER=0: NCCL reduce-scatter -> AdamW on the slice -> NCCL all-gather."""
    monkeypatch.setenv("CAPDEC_DP_PEER", "0")
    _check_against_single_gpu(tmp_path, expect_peer=False)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_segmented_graph_overlap_matches_single_gpu(tmp_path, monkeypatch):
    """Per-block all-reduce issued eagerly between the 13 graph segments of the step (trainer._capture_segments)."""
    monkeypatch.setenv("CAPDEC_DP_OVERLAP", "2")
    _check_against_single_gpu(tmp_path)
