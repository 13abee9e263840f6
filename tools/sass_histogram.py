#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native paths: disassembles cim_b200/libcimhead.so with
`cuobjdump -sass` and counts, per kernel, the mnemonics /opt/skills/guides/B200_PROFILING.md lists as proof:

    UTC*MMA          tcgen05.mma                 LDTM / STTM       tcgen05.ld / tcgen05.st (tensor memory)
    UTMALDG/UTMASTG  TMA tensor copies           UBLKCP            cp.async.bulk (1-D bulk copies)
    UTMAREDG/UBLKRED TMA / bulk reductions       LDGSTS            cp.async
    FFMA2            packed fp32 FMA             RED / ATOM        global reductions / atomics
    HMMA             legacy mma.sync (there must be none)

    python tools/sass_histogram.py [out.txt]      (`make -C cim_b200/csrc sass` writes profiles/sass_histogram_r2.txt)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cim_b200", "libcimhead.so")
COLS = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "LDGSTS", "FFMA2", "FFMA", "POPC", "RED",
        "ATOM", "HMMA", "SYNCS", "total"]


def classify(op):
    if op.startswith("UTC") and "MMA" in op:
        return "UTC*MMA"
    for k in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "UTMAREDG", "LDGSTS", "FFMA2", "POPC", "HMMA",
              "SYNCS"):
        if op.startswith(k):
            return "UBLKRED" if k == "UTMAREDG" else k
    if op.startswith("FFMA"):
        return "FFMA"
    if op.startswith("RED"):
        return "RED"
    if op.startswith("ATOM"):
        return "ATOM"
    return None


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except (OSError, subprocess.CalledProcessError):
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, mma_kinds = collections.OrderedDict(), None, collections.defaultdict(collections.Counter)
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = ins.match(line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["total"] += 1
        k = classify(op)
        if k:
            counts[cur][k] += 1
        if k == "UTC*MMA":
            mma_kinds[cur][op + m.group(2)] += 1
    names = demangle(list(counts))
    def short(n):
        d = re.sub(r"\((?:bool|int|unsigned int)\)", "", names[n]).replace("(anonymous namespace)::", "")
        return re.sub(r"\(.*", "", d.replace("<unnamed>::", "")).replace("void ", "")
    lines = ["# SASS mnemonic counts per kernel of cim_b200/libcimhead.so (cuobjdump -sass, sm_100a); written by "
             "tools/sass_histogram.py", "# " + " ".join(f"{c:>8s}" for c in COLS) + "  kernel"]
    tot = collections.Counter()
    for n, c in sorted(counts.items(), key=lambda kv: short(kv[0])):
        tot.update(c)
        lines.append("  " + " ".join(f"{c.get(k, 0):8d}" for k in COLS) + "  " + short(n))
        for kind, cnt in sorted(mma_kinds[n].items()):
            lines.append(" " * (9 * len(COLS) + 4) + f"  {cnt} x {kind}")
    lines.append("  " + " ".join(f"{tot.get(k, 0):8d}" for k in COLS) + "  ALL KERNELS")
    lines.append(f"# legacy tensor path (HMMA / mma.sync): {tot.get('HMMA', 0)} instructions")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
