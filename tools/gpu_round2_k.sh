#!/bin/bash
# GPU call K (round 2): cta_group::2 weight-gradient kernel: parity / bit-equality tests, in-situ A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "wgrad" > gpurun_out/r2k_pytest_wgrad.log 2>&1; echo "pytest wgrad rc=$?"
tail -3 gpurun_out/r2k_pytest_wgrad.log
for v in 0 1 0 1; do
  GDRN_WGRAD_2CTA=$v timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed quick wgrad2cta=$v', d['value'], d['ms_per_step'], d['clocks'])"
done
for v in 0 1; do
  GDRN_WGRAD_2CTA=$v GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half quick wgrad2cta=$v', d['value'], d['ms_per_step'])"
done
GDRN_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --kernel-name-base demangled -k regex:"wgrad" --csv --log-file gpurun_out/r2k_wgrad_tp.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r2k_ncu.log 2>&1; echo "ncu rc=$?"
python tools/tensor_pipe_summary.py gpurun_out/r2k_wgrad_tp.csv | head -30
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2k_pytest.log
