"""GPU (-m gpu), two or more GPUs on the box: parity of the sharded device paths against the CPU oracle — every rank's
k-NN rows (windowed index), radius shards, the run-sharded repel with the NVLink peer-memory exchange, identical stop
decisions on every rank. One process per GPU under torch.distributed.run, rendezvous on 127.0.0.1. Skipped (not
failed) on a single-GPU box; the same checks run inside bench.py under world > 1 (`parity_check` in its JSON line)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_paths_match_oracle(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    port = 29700 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "multigpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0 and "MULTIGPU_CHECK OK" in r.stdout, tail


@pytest.mark.parametrize("world", [2])
def test_single_process_multi_device_context(world):
    """wtp_create_multi: ONE process (a Julia session is one process), one context over several GPUs; set_topology /
    repel calls on it are answered by all devices and fill the caller's arrays like a single-device context."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs on one box")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "multidevice_check.py"), str(world)], cwd=ROOT, capture_output=True,
                       text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0 and "MULTIDEVICE_CHECK OK" in r.stdout, tail
