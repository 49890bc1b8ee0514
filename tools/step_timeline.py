"""Where the timed step's time goes beyond the sum of its kernels: a CUPTI kernel timeline (torch.profiler) of CUDA-graph
replays of the C2 train step.  Reports, per replay: wall span (first kernel start -> last kernel end), busy time (union of
all kernel intervals), idle gaps, time during which two of our kernels overlap (two-stream backward), and the per-kernel
sums - the numbers `ncu` cannot give because it serialises launches.

    python tools/step_timeline.py [--steps 3] > gpurun_out/step_timeline.md
"""
import argparse
import collections
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    a = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    import capdec_b200 as cb
    import bench
    torch.manual_seed(0)
    model = cb.ClipCaptionModel(10, prefix_size=512, gpt_config=cb.GPT2Config()).to("cuda").train()
    tr = cb.Trainer(model, batch_size=256, seq_len=40, noise_variance=0.016, use_cuda_graph=True)
    tok, pfx = bench.synth_batch(256, 1)
    tok, pfx = tok.cuda(), pfx.cuda()
    for _ in range(6):
        tr.step(tok, pfx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tr.step_device()
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / 10
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            tr.step_device()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    if not ks:
        print("no CUDA kernel events captured (CUPTI unavailable?)")
        return
    # split into replays at the step_clock kernel
    starts = [i for i, k in enumerate(ks) if "step_clock" in k[2]]
    print(f"# r2 — kernel timeline of the CUDA-graph'ed C2 step (CUPTI via torch.profiler; {len(ks)} kernel records, "
          f"{len(starts)} steps; un-profiled step time {ms_plain:.3f} ms)\n")
    print("| step | span ms | busy (union) ms | idle gaps ms | two kernels overlapping ms | sum of kernel durations ms | kernels | "
          "gap to the next step's first kernel ms |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    per_name = collections.Counter()
    for si, s in enumerate(starts):
        e = starts[si + 1] if si + 1 < len(starts) else len(ks)
        seg = ks[s:e]
        t0, t1 = seg[0][0], max(k[1] for k in seg)
        busy = overlap = 0.0
        cur_end = t0
        total = 0.0
        for st, en, name in seg:
            total += en - st
            if si == len(starts) - 1 or True:
                per_name[name.split("(")[0].split("<")[0]] += (en - st) / len(starts)
            if st >= cur_end:
                busy += en - st
                cur_end = en
            else:
                overlap += min(en, cur_end) - st
                if en > cur_end:
                    busy += en - cur_end
                    cur_end = en
        span = t1 - t0
        nxt = (ks[e][0] - t1) / 1e3 if e < len(ks) else float("nan")
        print(f"| {si} | {span / 1e3:.3f} | {busy / 1e3:.3f} | {(span - busy) / 1e3:.3f} | {overlap / 1e3:.3f} | {total / 1e3:.3f} | {len(seg)} | {nxt:.3f} |")
    print("\n| kernel | us per step (sum of durations) |\n|---|---:|")
    for n, v in per_name.most_common(12):
        print(f"| `{n}` | {v:.1f} |")


if __name__ == "__main__":
    main()
