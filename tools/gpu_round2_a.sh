#!/bin/bash
# GPU call A (round 2): tests, bench in the new headline mode, launch list + in-situ ablation of the mixed-mode step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.log; tail -5 gpurun_out/r2a_bench.err
GDRN_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_mixed.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
timeout 400 python tools/ablate_step.py 64 mixed > gpurun_out/r2a_ablate_mixed.txt 2>&1; echo "ablate rc=$?"
tail -20 gpurun_out/r2a_ablate_mixed.txt
