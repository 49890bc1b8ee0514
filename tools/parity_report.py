"""gpurun_out/parity_report.jsonl (appended by the GPU parity tests) -> a markdown table for profiles/.
    python tools/parity_report.py gpurun_out/r2d_parity_report.jsonl > profiles/r2_parity_report.md"""
import json
import sys


def f(x, fmt="%.1e"):
    return "-" if x is None else fmt % x


def main():
    recs = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
    print("| test | case | mode | loss (ours) | loss (reference / oracle) | loss rel | worst grad rel-L2 (well-conditioned) | median | "
          "worst ill-conditioned (its fp32-vs-fp64 noise) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in recs:
        t = r.get("test")
        if t == "scale_parity":
            ill = "-"
            if r.get("worst_ill_conditioned"):
                ill = f"{f(r['worst_ill_conditioned_rel_l2'])} {r['worst_ill_conditioned'].replace('clip_project.', '')} ({f(r['worst_ill_conditioned_fp32_vs_fp64'])}), {r['n_ill_conditioned']} tensors"
            print(f"| scale B={r['B']} ({int(r['n_valid'])} targets, {r['live_rows']} live rows) | {r['case']} | {r['mode']} | {r['loss']:.9f} | "
                  f"{r['loss_ref']:.9f} | {f(r['loss_rel'])} | {f(r['worst_grad_rel_l2'])} {r['worst_grad'].replace('gpt.transformer.', '')} | "
                  f"{f(r['median_grad_rel_l2'])} | {ill} |")
        elif t == "fast_path":
            print(f"| fast path vs golden | {r['case']} | {r['mode']} | {r['loss']:.9f} | {r['loss_ref']:.9f} | {f(r['loss_rel'])} | "
                  f"{f(r['worst_grad_rel_l2'])} {str(r['worst_grad']).replace('gpt.transformer.', '')} | - | - |")
        elif t == "drop_in":
            print(f"| drop-in (train.py:348-351 verbatim) | {r['case']} | {r['mode']} | {r['loss']:.9f} | {r['loss_ref']:.9f} | "
                  f"{f(abs(r['loss'] - r['loss_ref']) / abs(r['loss_ref']))} | logits rel-L2 {f(r['logits_rel_l2'])}; sampled grads "
                  f"{f(r['worst_sampled_grad_err_over_norm'])} | - | - |")
        elif t == "trainer_trajectory":
            print(f"| Trainer, 3 AdamW steps vs oracle | mlp_full_b4 | fp32 | {r['losses'][-1]:.6f} | {r['ref_losses'][-1]:.6f} | - | "
                  f"worst parameter rel-L2 {f(r['worst_param_rel_l2'])} | - | - |")


if __name__ == "__main__":
    main()
