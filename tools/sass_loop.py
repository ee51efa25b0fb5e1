"""Opcode histogram of the innermost loop that contains a given opcode (default STG) in one kernel's SASS.
Usage: python tools/sass_loop.py <obj or .so> <kernel name substring> [opcode]"""
import collections
import re
import subprocess
import sys

path, pat = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else "STG"
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
fn, body = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        body[fn] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and fn:
        body[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ins in body.items():
    if pat not in fn:
        continue
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, text) in enumerate(ins):
        m = re.search(r"BRA(?:\.\S+)? (0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            loop = ins[addr[int(m.group(1), 16)]:i + 1]
            if any(want in t for _, t in loop) and (best is None or len(loop) < len(best)):
                best = loop
    if best is None:
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in best)
    print("==", fn[-90:], "| innermost loop with %s: %d instructions" % (want, len(best)))
    print("  " + "  ".join("%s:%d" % kv for kv in ops.most_common(24)))
