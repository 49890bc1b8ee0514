"""C5 probe: KV-cached batched beam search throughput (captions/s) at several batch sizes.
Usage: python tools/decode_probe.py [n_img ...].  The CPU baseline (the oracle restatement of the reference's
re-forward-everything generate_beam) is timed by `python bench.py --workload c5 [--impl reference]`."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import capdec_b200 as cb
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 64, 256]
    P, D, L = 10, 512, 67
    torch.manual_seed(0)
    model = cb.ClipCaptionModel(P, prefix_size=D, mapping_type=cb.MappingType.MLP, gpt_config=cb.GPT2Config()).to("cuda").eval()
    out = []
    for n_img in sizes:
        x = torch.randn(n_img, D, device="cuda")
        x = x / x.norm(2, -1, keepdim=True)
        embed = model.clip_project(x).view(n_img, P, -1)
        for rep in range(3):  # first call allocates + captures the step graph
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            l0 = cb._lib.load().capdec_launch_count()
            res = cb.generate_beam_ids(model, embed, 5, L, 1.0, -1)  # stop token never emitted: all 67 steps (worst case)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        rec = {"n_img": n_img, "beam": 5, "entry_length": L, "s_per_batch": dt, "captions_per_s": n_img / dt,
               "ms_per_decode_step": dt / L * 1e3, "tokens_len": len(res[0][0][0])}
        print(json.dumps(rec), flush=True)
        out.append(rec)


if __name__ == "__main__":
    main()
