"""Opcode histogram per kernel of libsrba_b200.so (cuobjdump -sass), for profiles/: which pipes / data-movement instructions the
kernels really use (DFMA/DMUL/DADD = FP64 pipe, DMMA = FP64 tensor pipe, LDG/STG/LDS/STS/LDL/STL, BAR, UCGABAR = cluster barrier,
MEMBAR, SHFL ...).  Usage: python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "srba_b200/libsrba_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?\s", line)
    if m and kern:
        hist[kern][m.group(2)] += 1
KEY = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "LDG", "LD", "STG", "ST", "LDS", "STS", "LDL", "STL", "ATOMG", "RED", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT", "MEMBAR", "CCTL", "SHFL",
       "BSSY", "BSYNC", "CALL", "LDGSTS", "UBLKCP", "SYNCS"]
print("%-64s %7s  %s" % ("kernel", "instr", " ".join("%s" % k for k in KEY)))
for k, h in hist.items():
    if "wave_" not in k and "srba_" not in k:
        continue
    print("%-64s %7d  %s" % (k[:64], sum(h.values()), " ".join("%s=%d" % (q, h[q]) for q in KEY if h[q])))
