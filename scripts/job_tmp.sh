timeout 900 python -m pytest tests -m gpu -x -q -k "radius or topology or smoke or entry" > gpurun_out/pytest_v37.log 2>&1; echo "pytest_rc=$?"; tail -3 gpurun_out/pytest_v37.log
timeout 600 python scripts/bench_extra.py radius > gpurun_out/extra_radius_v37.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/extra_radius_v37.log
