"""Per-kernel SASS evidence for the tcgen05 / TMA / DSMEM claims: counts of the mnemonics B200_PROFILING.md lists
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk
copy, SYNCS = mbarrier, HMMA = mma.sync) in the built library.  No GPU needed:
    python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "transformer-inertial-poser_b200", "lib", "libtip_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU", "LDGSTS",
             "LDSM", "ST.E", "STS", "LDS", "BAR.SYNC", "UCGABAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + ".") or (mn.endswith(".E") and op.startswith(mn)):
                    counts[cur][mn] += 1
            counts[cur]["_total"] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass), sm_100a")
    print("# " + " ".join(f"{m:>8}" for m in ["total"] + MNEMONICS) + "  kernel")
    tot = collections.Counter()
    for k, c in counts.items():
        name = demangle(k)
        name = re.sub(r"\(.*", "", name)
        print("  " + " ".join(f"{c[m]:>8}" for m in ["_total"] + MNEMONICS) + "  " + name)
        tot.update(c)
    print("  " + " ".join(f"{tot[m]:>8}" for m in ["_total"] + MNEMONICS) + "  ALL")


if __name__ == "__main__":
    sys.exit(main())
