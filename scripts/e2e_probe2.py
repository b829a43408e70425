"""e2e k-NN through the host ABI at the bench size; prints wall time per call (the library prints its pipeline line
with WTP_PIPE_DEBUG=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
n, k = 10_000_000, 21
pts = np.random.default_rng(0).random((n, 3)).astype(np.float32)
hp = torch.from_numpy(pts).pin_memory().numpy()
out = torch.empty((n, k), dtype=torch.int64).pin_memory().numpy()
ctx = pkg.Context(0)
for _ in range(3):
    ctx.knn(hp, k, out_idx=out)
ts = []
for _ in range(5):
    t = time.perf_counter(); ctx.knn(hp, k, out_idx=out); ts.append(time.perf_counter() - t)
print(f"stage_mb={os.environ.get('WTP_STAGE_MB')} slots={os.environ.get('WTP_STAGE_SLOTS')} threads={os.environ.get('WTP_HOST_THREADS')} "
      f"e2e={min(ts)*1e3:.2f} ms (median {sorted(ts)[2]*1e3:.2f})  {n/min(ts)/1e6:.1f} Mq/s", flush=True)
ctx.close()
