#!/usr/bin/env bash
# Instruction census of the shipped library (no GPU needed): per kernel, the Blackwell-specific SASS mnemonics that show
# tensor memory (LDTM / STTM), tensor-map TMA (UTMALDG / UTMASTG), bulk copies (UBLKCP), packed fp32 (FADD2 / FMUL2 /
# FFMA2), mbarriers (SYNCS), register reallocation (USETMAXREG) and cluster barriers (UCGABAR).
# Usage: tools/sass_census.sh > profiles/r02_sass_census.txt
cd "$(dirname "$0")/.."
lib=vndecorrelate_b200/_lib/libvnd_b200.so
echo "# cuobjdump -sass $lib ($(date -u +%F)), nvcc $(/usr/local/cuda/bin/nvcc --version | grep -o 'release [0-9.]*')"
cuobjdump -sass "$lib" | python3 -c '
import re, sys, collections
want = ["LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "FADD2", "FMUL2", "FFMA2", "SYNCS", "USETMAXREG", "UCGABAR", "MUFU", "LDS", "STS", "LDL", "STL", "DMUL", "HMMA", "UTCHMMA", "UTCMMA"]
cur, tot = None, collections.OrderedDict()
for ln in sys.stdin:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); tot[cur] = collections.Counter(); continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1); tot[cur]["_all"] += 1
        for w in want:
            if op.startswith(w): tot[cur][w] += 1
import subprocess
def demangle(n):
    try: return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:110]
    except Exception: return n[:110]
grand = collections.Counter()
for k, c in tot.items():
    grand.update(c)
    print("%-112s %6d instr  " % (demangle(k), c["_all"]) + " ".join("%s=%d" % (w, c[w]) for w in want if c[w]))
print("TOTAL " + " ".join("%s=%d" % (w, grand[w]) for w in want) + "  (no MMA instruction anywhere: the path is a sparse gather, BASELINE.json north_star)")
'
