"""executed-instruction mix of one kernel from an ncu report's source page:
    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python scripts/ncu_opcode_mix.py src.csv [warp_double_steps]
Prints executed warp instructions and stall samples per opcode, and -- given the number of warp double steps of the frame
(ray samples x max(steps) / 32) -- instructions per double step."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, st = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iE:
        continue
    t = re.sub(r'^\s*@!?U?P\d\s+', '', r[iS].strip())
    op = t.split()[0].split('.')[0] if t else "?"
    ex[op] += int(r[iE] or 0)
    st[op] += int(r[iSamp] or 0)
tot, stot = sum(ex.values()), sum(st.values())
norm = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("executed warp instructions: %.4g, stall samples %d" % (tot, stot))
packed = {"FFMA2", "FADD2", "FMUL2"}
fma = {"FFMA", "FADD", "FMUL", "FHADD", "IMAD", "HFMA2", "HADD2", "FFMA2", "FADD2", "FMUL2"}
cyc = sum(v * (2 if k in packed else 1) for k, v in ex.items() if k in fma)
print("FMA-pipe instructions %.4g (%.1f %%), FMA-pipe cycles (packed = 2) %.4g = %.3f per issued instruction" % (
    sum(v for k, v in ex.items() if k in fma), 100.0 * sum(v for k, v in ex.items() if k in fma) / tot, cyc, cyc / tot))
for k, v in ex.most_common(28):
    print("  %-8s %12d  %5.1f %%   samples %5.1f %%%s" % (k, v, 100.0 * v / tot, 100.0 * st[k] / max(stot, 1),
                                                          ("   %.1f / double step" % (v / norm)) if norm else ""))
