"""Graded-cloud k-NN probe: per-level failure counts of the tiled passes and the kernel timeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import __graft_entry__ as g
import synth
pkg = g.load_package()
ctx = pkg.Context(0)
ctx.set_timing(True)
pts, nw, hw = synth.graded_cube(2_000_000, np.float32)
idx = np.empty((len(pts), 21), dtype=np.int64)
os.environ["WTP_TILE_DEBUG"] = "1"
for _ in range(2):
    ctx.knn(pts, 21, out_idx=idx)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in ctx.timing().items()}, flush=True)
