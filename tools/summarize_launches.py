"""Turn the ncu launch lists (gpurun_out/launches.csv, launches_train.csv) into profiles/*.md summaries.

usage: python tools/summarize_launches.py <tag>      e.g. r02a
Eval list: per-layer table with algorithmic FLOPs -> effective TFLOP/s (2500 images = 25 five-shot episodes per forward).
"""
import collections
import csv
import sys

IMGS = 2500


def load(fn):
    rows = []
    with open(fn) as f:
        lines = [l for l in f if not l.startswith("==")]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((int(row["ID"]), row["Kernel Name"], row["Grid Size"], float(row["Metric Value"].replace(",", ""))))
    return rows


def short(name):
    return name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]


# MFLOP per image per launch
LAYERS = ([("stem conv1+downsample", 16.59), ("stem conv2", 235.93), ("stem conv3 + shortcut + LeakyReLU + maxpool + pos1 (fused epilogue)", 471.86)]
          + [x for i in range(4) for x in ((f"stage1.{i} mlp.conv1", 26.21),
                                           (f"stage1.{i} grouped 3x3 + GELU + conv3 + residual (fused)", 58.98 + 26.21))]
          + [("patch_embed2", 26.21)]
          + [x for i in range(2) for x in ((f"stage2.{i} qkv", 38.71), (f"stage2.{i} attention", 10.08), (f"stage2.{i} proj", 12.90),
                                           (f"stage2.{i} mlp conv1 + GELU + conv3 + residual (fused)", 104.86))]
          + [("patch_embed3", 26.21)]
          + [x for i in range(3) for x in ((f"stage3.{i} qkv", 39.17), (f"stage3.{i} attention", 1.28), (f"stage3.{i} proj", 13.06),
                                           (f"stage3.{i} mlp.conv1", 52.43), (f"stage3.{i} mlp.conv3", 52.43))]
          + [("final BN + pool", 0), ("episode head", 0)])


def main():
    tag = sys.argv[1]
    rows = load("gpurun_out/launches.csv")
    idx = [i for i, r in enumerate(rows) if "stem_in" in r[1]]
    seg = rows[idx[-1]: idx[-1] + len(LAYERS)]
    tot = sum(r[3] for r in seg)
    with open(f"profiles/{tag}_launches_eval_step.md", "w") as f:
        f.write(f"# {tag} -- ncu launch list of one eval forward (25 five-shot episodes = {IMGS} images)\n\n"
                "Command: `SUNB_BENCH_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "
                "gpurun_out/launches.csv python bench.py --steps 1 --warmup 1` (B200, 1 GPU; last forward of the run).\n"
                "Times under ncu are serialised and cold-cache: compare SHARES.  TFLOP/s = algorithmic FLOPs of the layer / its time.\n\n"
                "| layer | kernel | us | share | TFLOP/s |\n|---|---|---|---|---|\n")
        for (name, mflop), r in zip(LAYERS, seg):
            tf = f"{mflop * 1e6 * IMGS / (r[3] * 1e-9) / 1e12:.0f}" if mflop else "-"
            f.write(f"| {name} | `{short(r[1])}` | {r[3] / 1000:.1f} | {100 * r[3] / tot:.1f}% | {tf} |\n")
        f.write(f"\nTotal {tot / 1e6:.2f} ms for {len(seg)} launches = {2030.6e6 * IMGS / (tot * 1e-9) / 1e12:.0f} TFLOP/s over the forward.\n\n")
        by = collections.defaultdict(lambda: [0, 0.0])
        for r in seg:
            by[short(r[1])][0] += 1
            by[short(r[1])][1] += r[3]
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t / 1000:.1f} | {100 * t / tot:.1f}% |\n")
    rows = load("gpurun_out/launches_train.csv")
    idx = [i for i, r in enumerate(rows) if "pack_multi" in r[1]]          # first launch of a step: the weight-operand pack
    seg = rows[idx[-1]:]
    tot = sum(r[3] for r in seg)
    by = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        by[short(r[1])[:70]][0] += 1
        by[short(r[1])[:70]][1] += r[3]
    with open(f"profiles/{tag}_launches_train_step.md", "w") as f:
        f.write(f"# {tag} -- ncu launch list of one SUN-M meta-tuning step (480 images, 1 GPU, eager launches)\n\n"
                "Command: `SUNB_BENCH_PROFILE=train ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "
                "gpurun_out/launches_train.csv python bench.py --steps 1 --warmup 1` (last step of the run).\n\n"
                "| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t / 1000:.1f} | {100 * t / tot:.1f}% |\n")
        f.write(f"\n{len(seg)} launches, {tot / 1e6:.2f} ms = {3 * 2030.6e6 * 480 / (tot * 1e-9) / 1e12:.0f} TFLOP/s (3x forward FLOPs).  "
                "`at::` kernels are torch plumbing (zero fill of the gradient buffer / accumulator arena, gradient layout permutes).\n")
    print("wrote", tag)


if __name__ == "__main__":
    main()
