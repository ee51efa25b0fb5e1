"""Per-kernel SASS opcode histogram and a run-length view of the instruction stream (where the spills / shared loads sit
relative to the FMA blocks).  Usage: python tools/sass_ops.py <binary or .so> <substring of the mangled kernel name> [--stream]"""
import collections
import re
import subprocess
import sys

path, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
fn, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        rows[fn] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        rows[fn].append(m.group(1))
for fn, ops in rows.items():
    if pat not in fn:
        continue
    print("==", fn, len(ops), "instructions")
    hist = collections.Counter(o.split(".")[0] for o in ops)
    print("  " + "  ".join("%s:%d" % kv for kv in hist.most_common(24)))
    if "--stream" in sys.argv:
        runs, prev, n = [], None, 0
        for o in ops:
            k = o.split(".")[0]
            if k == prev:
                n += 1
            else:
                if prev:
                    runs.append("%s%s" % (prev, "x%d" % n if n > 1 else ""))
                prev, n = k, 1
        runs.append("%s x%d" % (prev, n))
        print("  " + " ".join(runs))
