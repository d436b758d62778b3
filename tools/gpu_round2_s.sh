#!/bin/bash
# GPU call S (round 2): BN forward / backward-apply walking their tensors back to front (L2 reuse between consecutive kernels)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "bn" > gpurun_out/r2s_pytest_bn.log 2>&1; echo "pytest bn rc=$?"
tail -2 gpurun_out/r2s_pytest_bn.log
for v in 0 1 0 1; do
  GDRN_BN_REVERSE=$v timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed quick bnrev=$v', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
for v in 0 1; do
  GDRN_BN_REVERSE=$v GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half quick bnrev=$v', d['value'], d['ms_per_step'])"
done
