"""Run the C2 train step eagerly (no CUDA graph) for profiling under ncu:

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 1

cudaProfilerStart/Stop bracket exactly `--steps` steps after warm-up.  `--summarise FILE.csv` prints the per-kernel
share table that goes into profiles/.
"""
import argparse
import collections
import csv
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def summarise(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0]
        agg[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    tot = sum(agg.values())
    print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in agg.most_common():
        print(f"| `{k}` | {cnt[k]} | {v / 1e3:.1f} | {100 * v / tot:.1f}% |")
    print(f"| **all** | {sum(cnt.values())} | {tot / 1e3:.1f} | 100% |")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--bs", type=int, default=256)
    ap.add_argument("--mapping", default="mlp")
    ap.add_argument("--only_prefix", action="store_true")
    ap.add_argument("--summarise", default=None)
    a = ap.parse_args()
    if a.summarise:
        return summarise(a.summarise)
    import torch
    import capdec_b200 as cb
    import bench
    torch.manual_seed(0)
    if a.mapping == "mlp":
        cls = cb.ClipCaptionPrefix if a.only_prefix else cb.ClipCaptionModel
        model = cls(10, prefix_size=512, gpt_config=cb.GPT2Config())
    else:
        model = cb.ClipCaptionModel(40, clip_length=40, prefix_size=512, mapping_type=cb.MappingType.Transformer,
                                    gpt_config=cb.GPT2Config())
    model = model.to("cuda").train()
    tr = cb.Trainer(model, batch_size=a.bs, seq_len=40, noise_variance=0.016, use_cuda_graph=False)
    tok, pfx = bench.synth_batch(a.bs, 1)
    tok, pfx = tok.cuda(), pfx.cuda()
    for _ in range(3):   # 2 eager warm-up steps (row hints + GEMM plan measurement happen after the 2nd) + 1 with the final plans
        tr.step(tok, pfx)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        tr.step(tok, pfx)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", tr.loss())


if __name__ == "__main__":
    main()
