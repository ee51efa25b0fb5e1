"""SASS evidence table kept under profiles/: for every kernel of the shipped library, the instruction count and the opcodes that
prove which hardware path it uses -- UBLKCP (1-D cp.async.bulk), UTMALDG / UTMASTG (tensor-map TMA loads / stores), SYNCS
(mbarrier), ACQBULK (griddepcontrol.wait), UCGABAR_ARV/_WAIT (cluster barrier), FFMA2 (packed fp32 FMA), plus the FFMA / LDS /
LDG / STG / SHFL counts.  Usage: python tools/sass_evidence.py [lib.so] > profiles/rN_sass_opcodes.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "wt-pse-code_b200", "libwtpse_b200.so")
COLS = ["UBLKCP", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "ACQBULK", "UCGABAR_ARV", "UCGABAR_WAIT", "FFMA2", "FFMA", "FADD", "LDS", "STS",
        "LDG", "STG", "SHFL", "ATOMG", "RED", "MUFU", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
fn, rows = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        rows[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        rows[fn][m.group(1)] += 1
        rows[fn]["_n"] += 1
print("# cuobjdump -sass %s (sm_100a), opcode counts per kernel (static instruction stream)" % os.path.basename(path))
print("%-78s %6s " % ("kernel", "instr") + " ".join("%7s" % c[:7] for c in COLS))
for (fn, cnt), name in sorted(zip(rows.items(), names), key=lambda t: t[1]):
    short = re.sub(r"\(anonymous namespace\)::", "", name)
    short = re.sub(r"^void ", "", short).split("(")[0].replace("wtpse::", "")
    print("%-78s %6d " % (short[:78], cnt["_n"]) + " ".join("%7s" % (cnt[c] or ".") for c in COLS))
