"""Summarise an ncu launch list (gpu__time_duration.sum per launch, csv) by kernel: count, total time, share.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md
"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        agg[row["Kernel Name"]].append(v)
    return agg


def main(path):
    agg = load(path)
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu launch list: {path}\n")
    print("cold-cache, serialised per-launch times (ncu --metrics gpu__time_duration.sum --clock-control none): compare SHARES\n")
    print("| kernel | launches | total ms | share | mean us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        name = k.replace("|", "/")
        if len(name) > 90:
            name = name[:87] + "..."
        print(f"| `{name}` | {len(v)} | {sum(v):.3f} | {100 * sum(v) / tot:.1f}% | {1e3 * sum(v) / len(v):.1f} | {1e3 * max(v):.1f} |")
    print(f"\ntotal {tot:.3f} ms over {sum(len(v) for v in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
