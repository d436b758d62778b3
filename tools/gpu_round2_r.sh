#!/bin/bash
# GPU call R (round 2): bench lines of the final build (headline with the 3-region e2e, YCB-V config, Patch-PnP config)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2r_bench.log 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench.log') if l.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['regions_ms'], 'parity', d['parity_b64']['pass'], d['clocks'])
"
timeout 300 python bench.py --config ycbv --steps 20 --warmup 5 > gpurun_out/r2r_bench_ycbv.log 2>/dev/null; echo "ycbv rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench_ycbv.log') if l.startswith('{')][-1])
print('ycbv value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
"
timeout 300 python bench.py --config pnp --nin 69 --steps 20 --warmup 5 > gpurun_out/r2r_bench_pnp.log 2>/dev/null; echo "pnp rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench_pnp.log') if l.startswith('{')][-1])
print('pnp value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
"
