"""Debug counters of the warp-per-cell FAST kernel (build with ORBB_NVCC_EXTRA=-DORBB_FW_STATS)."""
import ctypes, os, sys
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
import orbb200
from orbb200.synth import synth_frame
L = orbb200.lib()
NB = int(os.environ.get("FW_STATS_FRAMES", "4"))
def stats(reset=1):
    a = (ctypes.c_ulonglong * 8)()
    L.orbx_debug_fast_stats(a, reset)
    return list(a)
for name, w, h, nf, noise in (("euroc", 752, 480, 1000, False), ("kitti", 1241, 376, 2000, False), ("noise", 752, 480, 1000, True)):
    ex = orbb200.Extractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=NB)
    imgs = np.stack([synth_frame(s, w, h, noise_only=noise) for s in range(NB)])
    stats()
    ex.extract_batch(imgs)
    c = stats()
    print(name, dict(zip(["cells", "rounds", "second_rounds", "overflows", "queue_entries", "corners", "survivors"], c)),
          "entries/cell-round %.1f" % (c[4] / max(c[1] - c[3], 1)))
    ex.close()
