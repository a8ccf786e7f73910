"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel-family table (markdown)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        rows.append((row["Kernel Name"], v, row["Grid Size"]))
    agg = collections.OrderedDict()
    for k, v, g in rows:
        name = re.sub(r"\(.*", "", k).replace("void ", "").strip()
        name = re.sub(r"vmv::", "", name)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(v for _, v, _ in rows)
    lines = [f"launches: {len(rows)}   total device time (serialised, cold-cache): {tot / 1e3:.3f} ms", "",
             "| kernel | launches | total ms | share | avg us | max us |", "|---|---:|---:|---:|---:|---:|"]
    for k, (n, t, mx) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| `{k[:70]}` | {n} | {t / 1e3:.3f} | {100 * t / tot:.1f}% | {t / n:.1f} | {mx:.1f} |")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
