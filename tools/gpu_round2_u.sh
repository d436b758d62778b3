#!/bin/bash
# GPU call U (round 2, 8 GPUs): the N = 8 line of the final build, and the same with NCCL limited to fewer CTAs (A/B)
set -x
mkdir -p gpurun_out
run() { # tag, extra env
  start=$(date +%s)
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2u_bench_n8_$1.log 2> gpurun_out/r2u_bench_n8_$1.err; echo "bench N=8 $1 rc=$? wall=$(( $(date +%s) - start ))s"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2u_bench_n8_$1.log') if l.startswith('{')][-1])
print('N=8 $1 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('regions_ms'), d['clocks'])
"
}
run default "GDRN_DUMMY=1" 29531
run maxctas8 "NCCL_MAX_CTAS=8" 29532
run maxctas4 "NCCL_MAX_CTAS=4" 29533
tail -3 gpurun_out/r2u_bench_n8_default.err
