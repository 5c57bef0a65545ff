import os, sys
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
import orbb200
from orbb200.synth import synth_frame
img = synth_frame(0, 752, 480)
ex = orbb200.Extractor(1000, 1.2, 8, 20, 7, max_width=752, max_height=480, max_batch=1)
k, d = ex(img)
print("ok", len(k))
