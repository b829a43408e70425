"""Host-pipeline probe on the GPU box: e2e k-NN timing vs host threads, pinned vs pageable output."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
os.system("nproc; lscpu | egrep 'Model name|Socket|Thread|NUMA node\\(s\\)|^CPU\\(s\\)'; free -g | head -2")
pkg = g.load_package()
n, k = 10_000_000, 21
pts = np.random.default_rng(0).random((n, 3)).astype(np.float32)
hp = torch.from_numpy(pts).pin_memory().numpy()
out_pinned = torch.empty((n, k), dtype=torch.int64).pin_memory().numpy()
out_page = np.empty((n, k), dtype=np.int64)
for th in (sys.argv[1:] or ["4", "8", "16", "32"]):
    os.environ["WTP_HOST_THREADS"] = th
    ctx = pkg.Context(0)
    for name, o in (("pinned", out_pinned), ("pageable", out_page)):
        for _ in range(2):
            ctx.knn(hp, k, out_idx=o)
        os.environ["WTP_PIPE_DEBUG"] = "1"
        t = time.perf_counter(); ctx.knn(hp, k, out_idx=o); dt = time.perf_counter() - t
        os.environ.pop("WTP_PIPE_DEBUG")
        print(f"threads={th} out={name} e2e={dt*1e3:.2f} ms  {n/dt/1e6:.1f} Mq/s", flush=True)
    ctx.close()
