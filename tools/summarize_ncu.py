"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launches per kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None, title=""):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, ns in rows:
        key = re.sub(r"\(.*", "", name)
        key = re.sub(r"^void ", "", key)
        agg[key][0] += 1
        agg[key][1] += ns
    total = sum(v[1] for v in agg.values())
    lines = [f"# {title}", "", f"source: `{path}` — {len(rows)} launches, {total / 1e6:.3f} ms total "
             "(ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)", "",
             "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:110]}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "ncu launch summary")
