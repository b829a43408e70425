"""BASELINE config #5's size on one GPU: k-NN k = 21 on U3(100 M) float32 with the device entry point — time per step
(CUDA events) and a brute-force check of 32 random rows on the device (torch: all 100 M squared distances per row)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
import __graft_entry__ as g

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
K = 21
pkg = g.load_package()
ctx = pkg.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
ctx.set_stream(stream.cuda_stream)
ctx.set_timing(True)
gen = torch.Generator(device=dev); gen.manual_seed(0x57545035)
d_pts = torch.rand((n, 3), generator=gen, device=dev, dtype=torch.float32)        # generated on the device: 100 M points
d_idx = torch.empty((n, K), dtype=torch.int32, device=dev)                          # the int32 device table (8.4 GB)
step = lambda: ctx.knn_dev(d_pts.data_ptr(), n, 3, K, np.float32, d_idx.data_ptr(), idx32=True)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3): step()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
t = ctx.timing()
print(f"U3({n}) f32 k=21, int32 device table: {ms:.2f} ms per step = {n / ms / 1e3:.0f} Mq/s; phases "
      f"bbox {t['ms_bbox']:.2f} keys {t['ms_cellkey']:.2f} scatter {t['ms_sort']:.2f} place {t['ms_reorder']:.2f} queries {t['ms_query']:.2f}; "
      f"leftovers {int(t['n_leftover_sparse'])}/{int(t['n_leftover_dense'])}/{int(t['n_leftover_other'])}", flush=True)
rows = torch.randint(0, n, (32,), generator=torch.Generator().manual_seed(7)).tolist()
bad = 0
for i in rows:
    q = d_pts[i]
    dx = d_pts[:, 0] - q[0]; d2 = dx * dx
    dy = d_pts[:, 1] - q[1]; d2 = d2 + dy * dy
    dz = d_pts[:, 2] - q[2]; d2 = d2 + dz * dz                                      # same operation order as dist2_rn (no FMA in eager mode)
    d2[i] = float("inf")
    vals, idx = torch.topk(d2, K + 1, largest=False, sorted=True)
    mine = d_idx[i].to(torch.int64) - 1
    kth = vals[K - 1]
    ok = bool((d2[mine] <= kth).all()) and len(set(mine.tolist())) == K and bool(torch.equal(torch.sort(d2[mine]).values, vals[:K]))
    ordered = bool((d2[mine][1:] >= d2[mine][:-1]).all())
    if not (ok and ordered):
        bad += 1
        print("row", i, "differs", mine.tolist(), idx[:K].tolist(), flush=True)
print("CHECK_100M", "OK" if bad == 0 else f"FAILED ({bad} of 32 rows)", flush=True)
ctx.close()
