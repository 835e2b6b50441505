"""diagnostic: where the CUDA LIC volume and the oracle's differ most (python scripts/diag_licvol.py n [field] [filter] [noise_n] [h])"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vectorvisualization_b200 as vv
from vectorvisualization_b200 import configs, fields as F
from oracle import vvo
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
field = sys.argv[2] if len(sys.argv) > 2 else "curl"
filt = sys.argv[3] if len(sys.argv) > 3 else "triangle"
noise_n = int(sys.argv[4]) if len(sys.argv) > 4 else 256
h = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01
s = configs.cfg5(n=n, size=256, noise_n=noise_n)
if field != "curl":
    s.field = getattr(F, field)(n)
s.filter_row = F.filter_kernel(filt) if filt != "box" else None
s.params["stepSizeLIC"] = h
s.licvol_fp16 = 0
w0 = None
for wb in (-1, 0):
    want = vvo.OracleScene(s, weight_bits=wb).lic_volume()
    if w0 is not None:
        d = np.abs(want - w0) / np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
        print("oracle fused vs GL formula: max rel %.3g, > 1e-4: %d, > 1e-5: %d, > 1e-6: %d" % (d.max(), (d > 1e-4).sum(), (d > 1e-5).sum(), (d > 1e-6).sum()))
    w0 = want
    r = vv.Renderer(0)
    configs.apply_scene(r, s)
    r.updateLICVolume()
    got = r.readLICVolume().copy()
    err = np.abs(got - want)
    rel = err / np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
    print("n %d %s %s noise %d h %g weight_bits %d: max |want| %.4g, max abs err %.3g, max rel %.3g, voxels rel > 1e-4: %d, > 1e-5: %d, > 1e-6: %d"
          % (n, field, filt, noise_n, h, wb, np.abs(want).max(), err.max(), rel.max(), (rel > 1e-4).sum(), (rel > 1e-5).sum(), (rel > 1e-6).sum()))
    idx = np.argsort(rel.ravel())[::-1][:3]
    for i in idx:
        z, y, x = np.unravel_index(i, rel.shape)
        print("   voxel (x %d, y %d, z %d): got %.8g want %.8g rel %.3g" % (x, y, z, got[z, y, x], want[z, y, x], rel[z, y, x]))
