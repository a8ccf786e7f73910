"""ncu launch list with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch of ONE forward ->
per-kernel-family DRAM traffic (JSON for bench.py's roofline.traffic + a markdown table)."""
import collections
import csv
import json
import re
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)


def main(path, out_json, out_md):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = (row["ID"], row["Kernel Name"])
        d = per.setdefault(k, {})
        m = row["Metric Name"]
        if m.startswith("dram__bytes"):
            d[m] = to_bytes(row["Metric Value"], row["Metric Unit"])
        elif m == "gpu__time_duration.sum":
            d[m] = to_us(row["Metric Value"], row["Metric Unit"])
    fam = collections.OrderedDict()
    for (_, name), d in per.items():
        n = re.sub(r"\(.*", "", name).replace("void ", "").replace("vmv::", "").strip()
        n = re.sub(r"<.*", "", n)
        a = fam.setdefault(n, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("dram__bytes_read.sum", 0.0)
        a[2] += d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("gpu__time_duration.sum", 0.0)
    js = {k: {"launches": v[0], "dram_read_bytes": v[1], "dram_write_bytes": v[2], "time_us": v[3],
              "dram_bytes_per_launch": (v[1] + v[2]) / v[0]} for k, v in fam.items()}
    json.dump(js, open(out_json, "w"), indent=1)
    lines = ["| kernel family | launches | DRAM read MB | DRAM write MB | time ms (ncu, serialised) | DRAM GB/s |", "|---|---:|---:|---:|---:|---:|"]
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][3]):
        lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.1f} | {v[2] / 1e6:.1f} | {v[3] / 1e3:.3f} | {(v[1] + v[2]) / max(v[3], 1e-9) / 1e3:.0f} |")
    open(out_md, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:4])
