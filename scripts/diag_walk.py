"""diagnostic: the walk of one LIC-volume voxel, CUDA vs oracle, step by step (python scripts/diag_walk.py n x y z [field])"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vectorvisualization_b200 as vv
from vectorvisualization_b200 import configs, fields as F
from oracle import vvo
n, x, y, z = (int(v) for v in sys.argv[1:5])
s = configs.cfg5(n=n, size=256)
if len(sys.argv) > 5:
    s.field = getattr(F, sys.argv[5])(n)
s.licvol_fp16 = 0
pos = np.array([(x + 0.5) / n, (y + 0.5) / n, (z + 0.5) / n], dtype=np.float32)
r = vv.Renderer(0)
configs.apply_scene(r, s)
np.set_printoptions(linewidth=200, precision=9)
for wb in (-1, 0):
    o = vvo.OracleScene(s, weight_bits=wb)
    for d in (-1, 1):
        a = o.debug_walk(pos, d, 32)
        for variant in (0, 1):
            b = r.debugWalk(pos, d, 32, variant)
            ulp = (a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
            first = np.argwhere(ulp[:, :6] != 0)
            print("weight_bits %d dir %d variant %d: first position/field difference at step %s; ulp differences per step (pos xyz, field rgb, tap):"
                  % (wb, d, variant, first[0] if len(first) else None))
            print(ulp[:, :7].T)
            if len(first) and wb == -1 and variant == 0:
                k = int(first[0][0])
                for kk in range(max(0, k - 1), k + 1):
                    print("  step %d oracle: pos %s field %s | Pos2 %s step2 %s" % (kk, a[kk, :3], a[kk, 3:6], a[kk, 8:11], a[kk, 11:14]))
                    print("  step %d cuda  : pos %s field %s | Pos2 %s step2 %s" % (kk, b[kk, :3], b[kk, 3:6], b[kk, 8:11], b[kk, 11:14]))
                    print("  ulp Pos2 %s step2 %s" % (ulp[kk, 8:11], ulp[kk, 11:14]))
