#!/bin/bash
# GPU call J (round 2): validation of the final tree: tests, quick benches, full driver-style bench, smoke
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2j_pytest.log
for i in 1 2; do
  timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed quick', d['value'], d['ms_per_step'])"
done
GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half quick', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.log 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench.log') if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity_b64']['pass'], d['parity_b64']['head_rel_l2'], d['clocks'])
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | tail -c 200
GDRN_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --kernel-name-base demangled -k regex:"gemm" --csv --log-file gpurun_out/r2j_tensor_pipe.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r2j_ncu_tp.log 2>&1; echo "ncu tensor-pipe rc=$?"
python tools/tensor_pipe_summary.py gpurun_out/r2j_tensor_pipe.csv | head -30
