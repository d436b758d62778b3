#!/bin/bash
# GPU call C (round 2): 2-CTA kernel validation + A/B, full tests, bench, ablation
set -x
mkdir -p gpurun_out
GDRN_2CTA=1 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "2cta or conv_fwd" > gpurun_out/r2c_pytest_2cta.log 2>&1; rc2=$?; echo "pytest-2cta rc=$rc2"
tail -4 gpurun_out/r2c_pytest_2cta.log
if [ $rc2 -ne 0 ]; then export GDRN_2CTA=0; echo "2-CTA kernel failed its tests: continuing with GDRN_2CTA=0"; fi
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2c_pytest.log
for mode in mixed half; do
  for v in 0 1; do
    if [ $rc2 -ne 0 ] && [ $v -eq 1 ]; then continue; fi
    echo "== quick bench mode=$mode 2cta=$v"
    GDRN_2CTA=$v GDRN_BENCH_MODE=$mode timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
  done
done
echo "== quick bench mixed, BN bitmask off"
GDRN_BN_BITMASK=0 timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.log 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/r2c_bench.log; tail -3 gpurun_out/r2c_bench.err
timeout 400 python tools/ablate_step.py 64 mixed > gpurun_out/r2c_ablate_mixed.txt 2>&1; echo "ablate rc=$?"
tail -12 gpurun_out/r2c_ablate_mixed.txt
GDRN_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_mixed.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
