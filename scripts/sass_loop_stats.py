"""static look at the hot kernel: registers / stack and the opcode histogram of the walk loop of lic_sample_kernel
(python scripts/sass_loop_stats.py [path/to/lib-or-object] [mangled-name-substring]); needs only cuobjdump, no GPU.
The walk loop is taken to be the largest backward-branch span that contains no other equally large loop (i.e. the
largest innermost loop)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'vectorvisualization_b200', 'libvv_b200.so')
pat = sys.argv[2] if len(sys.argv) > 2 else "lic_sample_kernelILi1ELi1ELb0ELb0ELi2E"


def stats(lib, pat):
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.splitlines()
    i = next(k for k, l in enumerate(res) if pat in l)
    name = re.search(r'Function (\S+?):', res[i]).group(1)
    usage = res[i + 1].strip()
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, lib], capture_output=True, text=True).stdout
    ins = []
    for l in sass.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for a, t in ins:
        m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:P\d,\s*)?(0x[0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a:
                loops.append((tgt, a))
    # innermost loops: contain no other loop
    inner = [(lo, hi) for lo, hi in loops if not any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in loops)]
    inner.sort(key=lambda x: x[1] - x[0], reverse=True)

    def op(t):
        t = re.sub(r'^@!?U?P\d\s+', '', t)
        return t.split()[0].split('.')[0]
    out = {"name": name, "usage": usage, "total": len(ins), "loops": []}
    for lo, hi in inner[:3]:
        body = [t for ad, t in ins if lo <= ad <= hi]
        c = collections.Counter(op(t) for t in body)
        out["loops"].append({"n": len(body), "ops": dict(c.most_common(40))})
    return out


if __name__ == "__main__":
    o = stats(lib, pat)
    print(o["name"])
    print(o["usage"])
    print("total", o["total"])
    for L in o["loops"]:
        print("loop", L["n"], L["ops"])
