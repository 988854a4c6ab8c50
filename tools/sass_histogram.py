"""Opcode histogram of the shipped library's SASS, per kernel family — the evidence that the hot kernels are
Blackwell-native (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st,
UTCBAR = tcgen05.commit, SYNCS = mbarrier).  No GPU needed:

    python tools/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rlaif-v_b200", "librlaifv_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU", "HMMA", "RED", "ATOMG")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fam = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::", "")
            cur = fam.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur[op] += 1
            if op == "UTCHMMA" and ".2CTA" in m.group(2):
                cur["UTCHMMA.2CTA"] += 1
    print("%-58s %s" % ("kernel (template instance)", " ".join("%8s" % k for k in KEYS + ("total",))))
    tot = collections.Counter()
    for name, c in fam.items():
        if not any(c[k] for k in KEYS[:7]):
            continue
        print("%-58s %s" % (name[:58], " ".join("%8d" % c[k] for k in KEYS) + " %8d" % sum(v for k, v in c.items() if k != "UTCHMMA.2CTA")))
        tot.update(c)
    print("%-58s %s" % ("ALL tensor / TMA kernels", " ".join("%8d" % tot[k] for k in KEYS)))
    print("UTCHMMA with .2CTA (cta_group::2):", tot["UTCHMMA.2CTA"])
    print("kernels in the library:", len(fam))


if __name__ == "__main__":
    sys.exit(main())
