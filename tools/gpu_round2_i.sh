#!/bin/bash
# GPU call I (round 2, 8 GPUs): ROI / recent-kernel tests on GPU 0, then the N = 8 and N = 4 benches exactly as the driver launches them
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -s -k "roi or groupnorm or pose_errors" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2i_pytest.log
for n in 8 4; do
  start=$(date +%s)
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2i_bench_n$n.log 2> gpurun_out/r2i_bench_n$n.err; echo "bench N=$n rc=$? wall=$(( $(date +%s) - start ))s"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2i_bench_n$n.log') if l.startswith('{')][-1])
    print('N=$n value', d['value'], 'per_gpu', d['per_gpu'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
except Exception as e:
    print('parse failed', e)
PY
  tail -3 gpurun_out/r2i_bench_n$n.err
done
timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 quick', d['value'], d['ms_per_step'])"
