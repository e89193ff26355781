"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`)
into a per-kernel table (markdown).    python tools/ncu_summary.py X.csv [title] > profiles/....md"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)  # drop the argument list
        if not name.startswith(("wb::", "void wb::")):
            name = "(torch / other: outside the library)"
        rows.append((name + " grid=" + r["Grid Size"], float(r["Metric Value"].replace(",", ""))))
    agg = defaultdict(lambda: [0.0, 0])
    for k, ns in rows:
        agg[k][0] += ns
        agg[k][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"# {title}\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
    print("| total ms | share | launches | avg us | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {ns / 1e6:.2f} | {100 * ns / total:.1f}% | {n} | {ns / n / 1e3:.1f} | `{k}` |")
    print(f"\ntotal {total / 1e6:.2f} ms over {len(rows)} launches")


if __name__ == "__main__":
    main()
