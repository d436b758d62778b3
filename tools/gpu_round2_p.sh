#!/bin/bash
# GPU call P (round 2, 2 GPUs): the N = 2 bench exactly as the driver launches it, with PDL on (default) and off
set -x
mkdir -p gpurun_out
for v in 1 0; do
  start=$(date +%s)
  GDRN_PDL=$v timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$v bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n2_pdl$v.log 2> gpurun_out/r2p_bench_n2_pdl$v.err; echo "bench N=2 pdl=$v rc=$? wall=$(( $(date +%s) - start ))s"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2p_bench_n2_pdl$v.log') if l.startswith('{')][-1])
print('N=2 pdl=$v value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
"
done
tail -3 gpurun_out/r2p_bench_n2_pdl1.err
