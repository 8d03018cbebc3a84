"""Count the Blackwell-native SASS mnemonics per kernel of the built library -> profiles/rNN_sass_summary.txt.

    python tools/sass_summary.py [out.txt]

UTCIMMA / UTCHMMA = tcgen05.mma kind::i8 / kind::f16 (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA
loads, UTCBAR = tcgen05.commit, DMMA = mma.sync f64, REDG.E.ADD.F64 = float64 atomics, SYNCS = mbarrier operations.
"""
import collections
import re
import shutil
import subprocess
import sys

SO = "svgp_vae_b200/libsvgp_b200.so"
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM(?:\.x\d+)?|STTM(?:\.x\d+)?|UTMALDG(?:\.\dD)?(?:\.MULTICAST)?|"
                 r"UTMASTG|UBLKCP|DMMA(?:\.\dx\dx\d)?|HMMA|IMMA|SYNCS[A-Z.]*|REDG\.E\.ADD\.F64[A-Z.]*)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kern, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
        elif kern:
            for t in PAT.findall(line):
                counts[kern][re.sub(r"^SYNCS.*", "SYNCS(mbarrier)", t)] += 1
    names = list(counts)
    if shutil.which("cu++filt"):
        names = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS evidence of the native sm_100a code paths: cuobjdump -sass %s, mnemonics counted per kernel (tools/sass_summary.py)" % SO]
    for (k, c), d in zip(counts.items(), names):
        if c:
            lines.append("%-110s %s" % (d[:110], "  ".join("%s=%d" % kv for kv in sorted(c.items()))))
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
