"""CUDA path vs oracle on the seeded random scenes of tests/test_oracle_vs_ref.py (_random_scene): run on a B200 before promoting
it to a `-m gpu` test.   python scripts/random_parity.py [first_seed] [count]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vectorvisualization_b200 as vv  # noqa: E402
from oracle import vvo  # noqa: E402
from test_oracle_vs_ref import _random_scene  # noqa: E402
from util import compare_images, render_cuda  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 32
tables = vvo.illum_tables(40.0)
worst = 0
for seed in range(first, first + count):
    s = _random_scene(seed)
    need = "MALLO" in s.defines or "ZOECKLER" in s.defines
    ref, ref_cnt, ref_tot = vvo.OracleScene(s, illum_tables=tables if need else None).raycast()
    _, img, _, cnt, tot = render_cuda(vv, s)
    md, ps, mf = compare_images(vvo, img, ref)
    worst = max(worst, md)
    print("seed %d %-24s samples %6d / %6d  count map %s  max 8-bit diff %d  PSNR %.1f  max float diff %.3g"
          % (seed, s.defines.replace("#define ", "") or "plain", tot, ref_tot, "equal" if np.array_equal(cnt, ref_cnt) else "DIFFERS", md, ps, mf), flush=True)
print("worst 8-bit difference over %d scenes: %d" % (count, worst))
