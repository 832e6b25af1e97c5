#!/usr/bin/env python
"""Opcode histogram of the built kernels (cuobjdump -sass), per kernel function: the SASS mnemonics that prove a
Blackwell-native path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, tcgen05.commit ->
UTCBAR) and the legacy ones that must be absent (HMMA = mma.sync, HGMMA = wgmma).

    python tools/sass_summary.py > profiles/r2_sass_k1.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "ah_rag_b200", "build")
NATIVE = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMACCTL", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UCGABAR", "SYNCS", "UBLKCP")
LEGACY = ("HMMA", "HGMMA", "QGMMA", "IGMMA", "IMMA")


def main():
    objs = sys.argv[1:] or ["topk_tc2.o", "topk_tc.o", "select.o", "graph.o", "graph_shard.o", "env.o", "exchange.o", "lca.o", "topk_exact.o"]
    for o in objs:
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, o)], capture_output=True, text=True).stdout
        fn, hist = None, collections.OrderedDict()
        for line in txt.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
                hist[fn] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]+)", line)
            if m and fn:
                hist[fn][m.group(1)] += 1
        print(f"== {o}")
        for fn, h in hist.items():
            total = sum(h.values())
            nat = {k: v for k, v in h.items() if k.split(".")[0] in NATIVE}
            leg = {k: v for k, v in h.items() if k.split(".")[0] in LEGACY}
            top = ", ".join(f"{k}:{v}" for k, v in h.most_common(8))
            print(f"  {fn}\n    instructions {total}; Blackwell-native: {nat or 'none'}; legacy tensor ops: {leg or 'none'}\n    top: {top}")


if __name__ == "__main__":
    main()
