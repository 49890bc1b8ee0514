"""Time the peer-memory optimizer kernel (csrc/peer.cu) alone on N ranks: every rank owns 1/N of a 155.9 M-float buffer,
loads that slice of every rank's gradients over NVLink, updates, stores the slice into every rank's parameters.
Launch: python -m torch.distributed.run --nproc-per-node N tools/peer_probe.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import capdec_b200 as cb  # noqa: E402
from capdec_b200 import ops  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    n = 155_900_000 // (4 * world) * (4 * world)
    g = torch.randn(n, device="cuda") * 1e-3
    p = torch.randn(n, device="cuda")
    sh = n // world
    m, v = torch.zeros(sh, device="cuda"), torch.zeros(sh, device="cuda")
    lr, t, den = torch.full((1,), 1e-5, device="cuda"), torch.full((1,), 1.0, device="cuda"), torch.full((1,), 1.0, device="cuda")
    table = [None] * world
    dist.all_gather_object(table, (ops.peer_export(g), ops.peer_export(p)))
    gp, pp = [], []
    for r, ((gh, go), (ph, po)) in enumerate(table):
        if r == rank:
            gp.append(g.data_ptr()); pp.append(p.data_ptr())
        else:
            gp.append(ops.peer_open(gh, go)); pp.append(ops.peer_open(ph, po))
    fence = torch.zeros(4, device="cuda")

    def step():
        dist.all_reduce(fence)
        ops.adamw_peer_step([a + 4 * rank * sh for a in gp], pp, rank, rank * sh, sh, m, v, lr, t, grad_denom=den)
        dist.all_reduce(fence)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # kernel alone (the ranks start together after a barrier; stream order keeps the launches back to back)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.adamw_peer_step([a + 4 * rank * sh for a in gp], pp, rank, rank * sh, sh, m, v, lr, t, grad_denom=den)
    e1.record(); torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        step()
    e1.record(); torch.cuda.synchronize()
    s_ms = e0.elapsed_time(e1) / reps
    # the default (push) variant's kernel: all N gradient contributions are local (staging), only the parameter stores cross NVLink
    staging = torch.randn((world - 1) * sh, device="cuda") * 1e-3
    gl_slices = [g.data_ptr() + 4 * rank * sh if r == rank else staging.data_ptr() + 4 * sh * (r if r < rank else r - 1)
                 for r in range(world)]
    for _ in range(2):
        ops.adamw_peer_step(gl_slices, pp, rank, rank * sh, sh, m, v, lr, t, grad_denom=den)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.adamw_peer_step(gl_slices, pp, rank, rank * sh, sh, m, v, lr, t, grad_denom=den)
    e1.record(); torch.cuda.synchronize()
    p_ms = e0.elapsed_time(e1) / reps
    # the copy-engine push of one rank's gradients to its peers (what runs during the backward pass)
    sp = [None] * world
    dist.all_gather_object(sp, ops.peer_export(staging))
    s_ptrs = [staging.data_ptr() if r == rank else ops.peer_open(*sp[r]) for r in range(world)]
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        for r in range(world):
            if r != rank:
                slot = rank if rank < r else rank - 1
                ops.copy_async(s_ptrs[r] + 4 * slot * sh, g.data_ptr() + 4 * r * sh, 4 * sh)
    e1.record(); torch.cuda.synchronize()
    c_ms = e0.elapsed_time(e1) / reps
    # plain local AdamW on the slice for comparison
    pl, gl = p[rank * sh:(rank + 1) * sh], g[rank * sh:(rank + 1) * sh]
    e0.record()
    for _ in range(reps):
        ops.adamw_step(pl, gl, m, v, lr, t, grad_denom=den, zero_grad=False)
    e1.record(); torch.cuda.synchronize()
    l_ms = e0.elapsed_time(e1) / reps
    remote = (world - 1) * sh * 4 / 1e9
    if rank == 0:
        print(f"N={world}: peer kernel {k_ms:.3f} ms ({remote / k_ms * 1e3:.0f} GB/s in + the same out per rank over NVLink), "
              f"with its two 16-byte all-reduces {s_ms:.3f} ms, local AdamW on the 1/N slice {l_ms:.3f} ms", flush=True)
        print(f"N={world}: push-variant kernel (gradients local, parameter stores over NVLink) {p_ms:.3f} ms = {remote / p_ms * 1e3:.0f} GB/s out per rank; "
              f"copy-engine push of the gradients ({remote * 1e3:.0f} MB per rank, all ranks at once) {c_ms:.3f} ms = {remote / c_ms * 1e3:.0f} GB/s", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
