#!/usr/bin/env python
"""SASS evidence that the shipped library is Blackwell-native: per kernel of vit-ad_b200/lib/libvitad.so the counts of
tcgen05.mma (UTC*MMA), tcgen05.ld/st (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP) and legacy tensor-core (HMMA) instructions.
CPU only:  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vit-ad_b200", "lib", "libvitad.so")
PATTERNS = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"\bUTCHMMA\.2CTA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM",
            "UTMALDG": r"\bUTMALDG", "UTMALDG.MULTICAST": r"\bUTMALDG\S*MULTICAST", "UTMASTG": r"\bUTMASTG", "UBLKCP": r"\bUBLKCP",
            "HMMA (legacy)": r"\bHMMA", "HGMMA (sm_90)": r"\bHGMMA"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                per[cur][name] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (kernels without any of them omitted)")
    print("# " + " | ".join(PATTERNS))
    for (fn, cnt), name in zip(per.items(), demangled):
        total.update(cnt)
        if sum(cnt.values()) == 0:
            continue
        short = re.sub(r"CUtensorMap_st", "TMap", name)
        short = short if len(short) <= 150 else short[:147] + "..."
        print(f"{short}\n    " + "  ".join(f"{k}={cnt[k]}" for k in PATTERNS if cnt[k]))
    print("\nTOTAL  " + "  ".join(f"{k}={total[k]}" for k in PATTERNS))
    print(f"kernels: {len(per)}; with tcgen05.mma: {sum(1 for c in per.values() if c['UTCHMMA'])}; with legacy HMMA: {sum(1 for c in per.values() if c['HMMA (legacy)'])}")


if __name__ == "__main__":
    main()
