#!/bin/bash
# GPU call G (round 2, 2 GPUs): full test suite on GPU 0, then the N=2 bench (graph with NCCL + wgrad side stream) incl. clean exit
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2g_pytest.log
timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | tail -c 300
start=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.log 2> gpurun_out/r2g_bench_n2.err; echo "bench2 rc=$? wall=$(( $(date +%s) - start ))s"
tail -c 700 gpurun_out/r2g_bench_n2.log; tail -4 gpurun_out/r2g_bench_n2.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2g_bench_ref_n2.log 2>&1; echo "ref2 rc=$?"
tail -c 300 gpurun_out/r2g_bench_ref_n2.log
