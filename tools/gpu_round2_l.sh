#!/bin/bash
# GPU call L (round 2): RANSAC-PnP on the device vs cv2; then the whole suite (wgrad pair kernel now off by default) + quick bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pnp_ransac_gpu.py -m gpu -x -q -s > gpurun_out/r2l_pytest_pnp.log 2>&1; echo "pytest pnp rc=$?"
grep -v "^$" gpurun_out/r2l_pytest_pnp.log | tail -60 | cut -c1-220
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2l_pytest.log
timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed quick', d['value'], d['ms_per_step'])"
