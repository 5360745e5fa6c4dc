"""SASS mnemonic counts per kernel of the built library (profiles/rN_sass_counts.txt): proof of tcgen05 / TMA use.
usage: python scripts/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "denet_b200", "libdenet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.splitlines()
cols = [("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTMALDG", r"\bUTMALDG(?!\S*\.2CTA)"),
        ("UTMALDG.2CTA", r"\bUTMALDG\S*\.2CTA"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"),
        ("UTCBAR", r"\bUTCBAR(?!\S*MULTICAST)"), ("UTCBAR.MULTICAST", r"\bUTCBAR\S*MULTICAST"), ("HMMA", r"\bHMMA"),
        ("MEMBAR.ALL.GPU", r"\bMEMBAR\.ALL\.GPU")]
print("SASS mnemonic counts per kernel of denet_b200/libdenet_b200.so (cuobjdump -sass; sm_100a)")
print("UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; UTMALDG / UTMASTG = TMA load / store; LDTM = tcgen05.ld; "
      "UTCBAR = tcgen05.commit (.MULTICAST to both CTAs of a pair);")
print("HMMA (legacy mma.sync) must be 0 everywhere\n")
print("%-62s" % "kernel" + "".join("%17s" % c for c, _ in cols))
tot = collections.Counter()
for name, body in zip(names, re.split(r"Function : \S+", sass)[1:]):
    counts = [len(re.findall(rx, body)) for _, rx in cols]
    if sum(counts[:8]) == 0 and counts[8] == 0:
        continue
    short = re.sub(r"\(.*", "", name)
    print("%-62s" % short[:62] + "".join("%17d" % c for c in counts))
    for (c, _), v in zip(cols, counts):
        tot[c] += v
print("%-62s" % "total" + "".join("%17d" % tot[c] for c, _ in cols))
