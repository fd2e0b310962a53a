#!/usr/bin/env python
"""SASS evidence for the tcgen05 / TMA kernels: per kernel of libxfeat_b200.so, the counts of the Blackwell mnemonics
(B200_PROFILING.md "What proves a Blackwell-native kernel") -> profiles/rNN_sass_mnemonics.md.

  python tools/sass_summary.py > profiles/r02_sass_mnemonics.md"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
SO = REPO / "xfeatslam_b200" / "lib" / "libxfeat_b200.so"
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UBLKPF", "UTMASTG", "UTMALDG", "UTMACMDFLUSH", "SYNCS", "ELECT",
        "LDG.E.ENL2.256", "STG.E.ENL2.256", "FFMA2", "HMMA", "REDG", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            kernels[cur]["_n"] += 1
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w in ("LDG.E.ENL2.256", "STG.E.ENL2.256") and op.startswith(w)):
                    kernels[cur][w] += 1
    dem = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonics per kernel of `xfeatslam_b200/lib/libxfeat_b200.so` (`cuobjdump -sass`, sm_100a)\n")
    print("`UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st, `UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk, `UBLKPF` = "
          "cp.async.bulk.prefetch, `UTMASTG` / `UTMALDG` = cp.async.bulk.tensor store / load (TMA), `SYNCS` = mbarrier ops, `FFMA2` = fma.rn.f32x2, `LDG/STG.E.ENL2.256` = 256-bit global access.\n")
    cols = [w for w in WANT if any(k[w] for k in kernels.values())]
    print("| kernel | instructions | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for (name, c), d in zip(kernels.items(), dem):
        if not any(c[w] for w in ("UTCHMMA", "LDTM", "UBLKCP", "UTMASTG", "STTM", "FFMA2")):
            continue
        short = re.sub(r"\((?:int|bool)\)", "", d)
        short = re.sub(r"\((?:xfb::|const |CUtensorMap).*$", "", short).replace("void xfb::", "").replace("xfb::", "")
        print("| `%s` | %d | " % (short[:120], c["_n"]) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    tot = Counter()
    for c in kernels.values():
        tot.update(c)
    print("\nTotals over the library: " + ", ".join("%s %d" % (w, tot[w]) for w in cols) + ".")


if __name__ == "__main__":
    sys.exit(main())
