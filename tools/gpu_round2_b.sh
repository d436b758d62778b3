#!/bin/bash
# GPU call B (round 2, 2 GPUs): new tests + the N=2 bench with the NCCL all-reduces captured in the step graph
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_b64_gpu.py tests/test_patch_pnp_gpu.py tests/test_ops_gpu.py -m gpu -x -q -s -k "mixed_gradients or graph_train or symmetric or nin67 or integer_label or ycbv or patch_pnp or ranger or scale_f32" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.log 2> gpurun_out/r2b_bench_n2.err; echo "bench2 rc=$?"
tail -c 2500 gpurun_out/r2b_bench_n2.log; tail -12 gpurun_out/r2b_bench_n2.err
timeout 300 python bench.py --config pnp --steps 20 --warmup 5 > gpurun_out/r2b_bench_pnp.log 2> gpurun_out/r2b_bench_pnp.err; echo "pnp rc=$?"
tail -c 2500 gpurun_out/r2b_bench_pnp.log; tail -5 gpurun_out/r2b_bench_pnp.err
timeout 300 python bench.py --config ycbv --steps 20 --warmup 5 > gpurun_out/r2b_bench_ycbv.log 2> gpurun_out/r2b_bench_ycbv.err; echo "ycbv rc=$?"
tail -c 1500 gpurun_out/r2b_bench_ycbv.log; tail -5 gpurun_out/r2b_bench_ycbv.err
