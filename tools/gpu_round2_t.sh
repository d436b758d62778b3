#!/bin/bash
# GPU call T (round 2): validation of the final tree (tests, smoke, driver-style bench, reference arm)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench.log 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench.log') if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['regions_ms'], 'parity', d['parity_b64']['pass'], d['parity_b64']['head_rel_l2'], d['clocks'])
"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | tail -c 200
