"""Count the SASS mnemonics that prove the Blackwell paths (tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit ->
UTCBAR, cp.async.bulk[.tensor] -> UBLKCP / UTMALDG / UTMASTG) in the shipped library, in total and per kernel.  CPU-only:

    python tools/sass_digest.py [path/to/libursa_b200.so] > profiles/<tag>_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ursabench_b200", "libursa_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|UTCBAR[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|UBLKPF[.\w]*|"
                 r"HMMA[.\w]*|UTMAPF[.\w]*|SYNCS[.\w]*)")
total, per, fn = collections.Counter(), collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = pat.search(line)
    if m and fn:
        k = "SYNCS (mbarrier ops)" if m.group(1).startswith("SYNCS") else m.group(1)
        total[k] += 1
        per[fn][k] += 1


def demangle(name):
    try:
        return re.sub(r"\(.*", "", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
    except OSError:
        return name


print("SASS mnemonic counts of %s (cuobjdump -sass, sm_100a)\n" % os.path.relpath(lib, ROOT))
for k in sorted(total):
    print("%7d %s" % (total[k], k))
print("\nper kernel (kernels with tcgen05 / TMA / bulk-copy instructions):")
for f in sorted(per, key=lambda f: -sum(per[f].values())):
    c = per[f]
    if any(k.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLK")) for k in c):
        print("  %s: %s" % (demangle(f), ", ".join("%s x%d" % (k, c[k]) for k in sorted(c) if not k.startswith("SYNCS"))))
