"""profiles/rNN_sass_summary.txt: which tensor-core / TMA / TMEM instructions each kernel of libishape_b200.so contains
(cuobjdump -sass, no GPU needed).   python tools/sass_summary.py profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ishapediting_b200", "libishape_b200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UTMASTG", "UBLKPF", "SYNCS", "HMMA", "IMMA", "FFMA",
            "MUFU", "SHFL", "ATOM", "RED", "LDG", "LDS", "STS", "STG", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[kern]["_total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    counts[kern][p + ("." + ".".join(op.split(".")[1:3]) if p in ("UTCHMMA", "HMMA", "LDTM", "UTMALDG", "UTCBAR") else "")] += 1
    lines = ["SASS instruction census of ishapediting_b200/libishape_b200.so (sm_100a), by kernel.",
             "UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKPF = bulk L2 prefetch,",
             "UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync.", ""]
    for k, c in counts.items():
        if c["_total"] == 0:
            continue
        tc = {p: n for p, n in c.items() if p != "_total" and any(p.startswith(x) for x in ("UTC", "LDTM", "STTM", "UTMA", "UBLK", "HMMA", "IMMA"))}
        rest = {p: n for p, n in c.items() if p in ("FFMA", "MUFU", "SHFL", "ATOM", "RED", "SYNCS")}
        lines.append(f"{k}\n    {c['_total']} instructions; tensor/TMA/TMEM: {tc or 'none'}; other: {rest}")
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.txt")
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", dst, len(counts), "kernels")


if __name__ == "__main__":
    main()
