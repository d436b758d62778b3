#!/bin/bash
# GPU call N (round 2): programmatic dependent launch of the GEMM kernels: bit-equality chain test, in-situ A/B, whole suite
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "pdl or 2cta or conv_fwd" > gpurun_out/r2n_pytest_ops.log 2>&1; echo "pytest ops rc=$?"
tail -3 gpurun_out/r2n_pytest_ops.log
for v in 0 1 0 1; do
  GDRN_PDL=$v timeout 300 python bench.py --quick --steps 30 --warmup 5 2>gpurun_out/r2n_bench_err_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed quick pdl=$v', d['value'], d['ms_per_step'], d['clocks'])"
done
for v in 0 1; do
  GDRN_PDL=$v GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half quick pdl=$v', d['value'], d['ms_per_step'])"
done
tail -5 gpurun_out/r2n_bench_err_1.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.log 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2n_bench.log') if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity_b64']['pass'], d['parity_b64']['head_rel_l2'], d['clocks'], 'infer', d.get('inference'))
"
