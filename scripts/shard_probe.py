"""One GPU, one shard at a time: what a rank of a G-way sharded k-NN call costs (windowed index build + its run of
the queries), and that its compact table equals the rows of the unsharded call. No communicator is needed for k-NN,
so the ranks are shard-only contexts on the same device."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
import bench
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = 21
dev = torch.device("cuda", 0)
pts_h = bench.synth_uniform(n)
d_pts = torch.from_numpy(pts_h).to(dev)
stream = torch.cuda.current_stream(dev)
full = None
for world in (1, 2, 4, 8):
    for rank in sorted({0, world // 2, world - 1}):
        for window in ((True, False) if world > 1 else (True,)):
            if window: os.environ.pop("WTP_NO_WINDOW", None)
            else: os.environ["WTP_NO_WINDOW"] = "1"
            ctx = pkg.Context(0)
            ctx.set_stream(stream.cuda_stream)
            ctx.set_timing(True)
            if world > 1:
                ctx.comm_init(rank, world, None)
            b, e = ctx.shard(n)
            d_idx = torch.empty((e - b, K), dtype=torch.int64, device=dev)
            for _ in range(3):
                ctx.knn_dev(d_pts.data_ptr(), n, 3, K, np.float32, d_idx.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                ctx.knn_dev(d_pts.data_ptr(), n, 3, K, np.float32, d_idx.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            t = ctx.timing()
            ok = None
            if world == 1:
                full = d_idx.clone()
            else:
                own = torch.from_numpy(ctx.owned() - 1).to(dev)
                ok = bool(torch.equal(full[own], d_idx))
            print(json.dumps({"world": world, "rank": rank, "window": window and world > 1, "ms_step": e0.elapsed_time(e1) / 10,
                              "phases": {k: round(t[k], 4) for k in ("ms_bbox", "ms_cellkey", "ms_sort", "ms_reorder", "ms_query")},
                              "window_points": t["n_window_points"], "missed": t["n_window_missed"], "rows_equal_unsharded": ok}), flush=True)
            ctx.close()
