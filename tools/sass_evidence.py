"""Mnemonic counts per kernel of the built library (cuobjdump -sass): python tools/sass_evidence.py > profiles/rNN/sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "llicti_b200", "libllicti_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.RCP", "SHFL", "VOTE"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, cnt, n, ts = None, collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        n[cur] += 1
        for k in KEYS:
            if re.search(r"(?<![A-Z0-9.])" + re.escape(k) + r"(?![A-Z0-9])", line):
                cnt[cur][k] += 1
        if "UTCHMMA" in line and re.search(r"UTCHMMA tmem\[[^\]]*\], gdesc", line):
            ts[cur] += 1


def demangle(x):
    try:
        return subprocess.run(["c++filt", x], capture_output=True, text=True).stdout.strip()
    except OSError:
        return x


print("SASS evidence of the Blackwell-native paths in llicti_b200/libllicti_b200.so (cuobjdump -sass, sm_100a).")
print("Counts per kernel of the mnemonics B200_PROFILING.md names: UTCHMMA = tcgen05.mma (kind::f16), of which 'A-in-TMEM' take their A")
print("operand from tensor memory; LDTM / STTM = tcgen05.ld / tcgen05.st; UTCBAR = tcgen05.commit; UBLKCP = cp.async.bulk (TMA bulk copy of")
print("the packed weights); LDGSTS = cp.async (parameter prefetch of the lane / group decoders); SYNCS = mbarrier ops.")
print("No HMMA (legacy mma.sync) and no UTMALDG (tensor-map TMA: DESIGN.md 4.1, \"TMA\").\n")
for f in sorted(n, key=lambda f: ("cnn_tc" not in f, "lane" not in f, f)):
    name = re.sub(r"\(.*", "", demangle(f))[-62:]
    extra = f"  A-in-TMEM={ts[f]}" if ts[f] else ""
    print(f"{name:64s} instr {n[f]:6d}  " + "  ".join(f"{k}={v}" for k, v in cnt[f].items()) + extra)
