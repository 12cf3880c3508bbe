"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
time, launches and (when captured) DRAM traffic / achieved GB/s per kernel name; `--per-launch` also lists every launch."""
import csv
import re
import sys
from collections import defaultdict

UNIT = {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1, "B": 1}


def load(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = {}
    for r in csv.DictReader(lines):
        m = r.get("Metric Name")
        if m is None:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        rec = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "ns": 0.0, "rd": None, "wr": None})
        if m == "gpu__time_duration.sum":
            rec["ns"] = v * UNIT.get(unit, 1)
        elif m == "dram__bytes_read.sum":
            rec["rd"] = v * BYTES.get(unit, 1)
        elif m == "dram__bytes_write.sum":
            rec["wr"] = v * BYTES.get(unit, 1)
    return [launches[k] for k in sorted(launches)]


def short(name):
    key = re.sub(r"\(.*", "", name)
    return re.sub(r"^void ", "", key)


def main(path, out=None, title="", per_launch=False):
    rows = load(path)
    has_dram = any(r["rd"] is not None for r in rows)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        a = agg[short(r["name"])]
        a[0] += 1
        a[1] += r["ns"]
        a[2] += (r["rd"] or 0.0) + (r["wr"] or 0.0)
    total = sum(v[1] for v in agg.values())
    lines = [f"# {title}", "", f"source: `{path}` — {len(rows)} launches, {total / 1e6:.3f} ms total "
             "(ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)", ""]
    if has_dram:
        lines += ["| kernel | launches | total ms | share | avg us | DRAM MB/launch | DRAM GB/s |", "|---|---:|---:|---:|---:|---:|---:|"]
    else:
        lines += ["| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        row = f"| `{k[:110]}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |"
        if has_dram:
            row += f" {by / n / 1e6:.1f} | {by / ns:.0f} |"
        lines.append(row)
    if per_launch:
        lines += ["", "## every launch, in order", "", "| # | kernel | grid | us | DRAM MB | GB/s |", "|---:|---|---|---:|---:|---:|"]
        for i, r in enumerate(rows):
            by = (r["rd"] or 0.0) + (r["wr"] or 0.0)
            lines.append(f"| {i} | `{short(r['name'])[:70]}` | {r['grid']} | {r['ns'] / 1e3:.1f} | {by / 1e6:.1f} | {by / max(r['ns'], 1):.0f} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0], args[1] if len(args) > 1 else None, args[2] if len(args) > 2 else "ncu launch summary", "--per-launch" in sys.argv)
