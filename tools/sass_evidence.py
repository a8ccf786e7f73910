"""SASS evidence (no GPU needed): per kernel of the built library, the Blackwell-native instruction counts
(UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA tensor / bulk copies, UTCBAR = tcgen05.commit,
PREEXIT / ACQBULK = griddepcontrol.launch_dependents / wait, HMMA = legacy mma.sync) and registers / static smem.
Usage: python tools/sass_evidence.py > profiles/<name>.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "videomv_b200", "lib", "obj")
PATS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "PREEXIT", "ACQBULK", "HMMA", "LDGSTS", "MUFU", "STG.E.ENL2.256",
        "RED.E", "ATOMG"]


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except FileNotFoundError:
        return n


def main():
    print("| kernel | regs | static smem | " + " | ".join(PATS) + " |")
    print("|---|---:|---:|" + "---:|" * len(PATS))
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith(".o"):
            continue
        path = os.path.join(OBJ, f)
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        usage = {}
        for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:\d+ SHARED:(\d+)", res):
            usage[m.group(1)] = (m.group(2), m.group(3))
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        cur, counts = None, collections.OrderedDict()
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                for p in PATS:
                    if op.startswith(p):
                        counts[cur][p] += 1
        for k, c in counts.items():
            name = demangle(k)
            name = re.sub(r"\(.*", "", name.replace("(int)", "")).replace("void ", "").replace("vmv::", "")
            r, sm = usage.get(k, ("?", "?"))
            print(f"| `{name[:60]}` | {r} | {sm} | " + " | ".join(str(c.get(p, 0)) for p in PATS) + " |")


if __name__ == "__main__":
    main()
