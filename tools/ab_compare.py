"""Compare per-shape GEMM tables (tools/gemm_breakdown.py output) of two builds: python tools/ab_compare.py A1.md,A2.md B1.md,B2.md"""
import re
import sys


def load(files):
    acc = {}
    for f in files.split(","):
        for l in open(f):
            m = re.match(r"\| (mode\d M\d+ N\d+ K\d+ act\d res\d rb\d split\d) \| (\d+) \| ([\d.]+) \|", l)
            if m:
                acc.setdefault(m.group(1), [int(m.group(2)), []])[1].append(float(m.group(3)))
    return {k: (n, min(v)) for k, (n, v) in acc.items()}


a, b = load(sys.argv[1]), load(sys.argv[2])
ta = sum(n * us for n, us in a.values()) / 1e3
tb = sum(n * us for n, us in b.values()) / 1e3
print(f"A total {ta:.3f} ms   B total {tb:.3f} ms   B/A = {tb / ta:.4f}")
rows = sorted(((b[k][1] / a[k][1], k, a[k][0], a[k][1], b[k][1]) for k in a if k in b), key=lambda r: (r[4] - r[3]) * r[2])
print("largest gains (B faster):")
for r in rows[:10]:
    print(f"  {r[1]:58s} n={r[2]:3d}  A {r[3]:7.1f}  B {r[4]:7.1f} us  ({(r[4] - r[3]) * r[2] / 1e3:+.3f} ms)")
print("largest losses (B slower):")
for r in rows[-10:]:
    print(f"  {r[1]:58s} n={r[2]:3d}  A {r[3]:7.1f}  B {r[4]:7.1f} us  ({(r[4] - r[3]) * r[2] / 1e3:+.3f} ms)")
