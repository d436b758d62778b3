#!/bin/bash
# GPU call Q (round 2): PDL mode 3 (forward elementwise kernels too) vs the default mode 1; validation of the final tree
set -x
mkdir -p gpurun_out
for v in 1 3 1 3; do
  GDRN_PDL=$v timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed quick pdl=$v', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
for v in 1 3; do
  GDRN_PDL=$v GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half quick pdl=$v', d['value'], d['ms_per_step'])"
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2q_pytest.log
GDRN_PDL=3 timeout 600 python -m pytest tests/test_parity_b64_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2q_pytest_pdl3.log 2>&1; echo "pytest pdl3 rc=$?"
tail -2 gpurun_out/r2q_pytest_pdl3.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.log 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2q_bench.log') if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity_b64']['pass'], d['parity_b64']['head_rel_l2'], d['clocks'])
"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | tail -c 300
