#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma
(UTC*MMA), tcgen05.ld / st (LDTM / STTM), the bulk-copy TMA engine (UBLKCP), tcgen05.commit (UTCBAR), packed fp32
(FFMA2 / FADD2 / FMUL2), mbarrier (SYNCS) — from ``cuobjdump -sass`` of the built library.  Runs on the CPU box.

    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bgflow_b200", "libbgflow_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "MUFU",
        "HMMA", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, it = {}, [], None, iter(names)
    for line in sass.splitlines():
        if "Function :" in line:
            cur = next(it)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    print("# cuobjdump -sass bgflow_b200/libbgflow_b200.so : instruction counts per kernel (static)")
    print("# " + " ".join(f"{k:>8s}" for k in ["total"] + KEYS) + "  kernel")
    # one line per kernel family (template instances share a prefix): show the first instance and the number of instances
    fam = collections.OrderedDict()
    for n in order:
        base = re.sub(r"<.*", "", n)
        fam.setdefault(base, []).append(n)
    for base, inst in fam.items():
        n = inst[0]
        c = counts[n]
        print("  " + " ".join(f"{c[k]:8d}" for k in ["_total"] + KEYS) + f"  {n}   [{len(inst)} instance(s)]")
    tc = [n for n in order if counts[n]["UTCHMMA"]]
    print(f"# kernels with tcgen05.mma (UTCHMMA): {len(tc)} instances in {len({re.sub(r'<.*', '', n) for n in tc})} families; "
          f"tensor-map TMA (UTMALDG/UTMASTG): {sum(counts[n]['UTMALDG'] + counts[n]['UTMASTG'] for n in order)} "
          f"(weight / tile traffic uses 1-D bulk copies, UBLKCP: pre-swizzled tiles need no tensor map); "
          f"legacy HMMA: {sum(counts[n]['HMMA'] for n in order)}")


if __name__ == "__main__":
    sys.exit(main())
