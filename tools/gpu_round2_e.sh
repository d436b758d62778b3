#!/bin/bash
# GPU call E (round 2): GEMM ncu --set full captures (demangled kernel names), bn_bwd_reduce tweak check, tests, full bench, launch list
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2e_pytest.log
for i in 1 2; do
  echo "== quick bench mixed (bn_bwd_reduce 4 rows in flight)"
  timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -f"
RUN="python bench.py --quick --no-graph --steps 1 --warmup 3"
export GDRN_PROFILE=1
timeout 400 $NCU -k regex:"gemm_fwd2_kernel<128" --launch-skip 35 --launch-count 1 -o gpurun_out/r2e_ncu_gemm_fwd2_128x3 $RUN > gpurun_out/r2e_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 400 $NCU -k regex:"gemm_fwd2_kernel<256" --launch-count 1 -o gpurun_out/r2e_ncu_gemm_fwd2_256x1 $RUN > gpurun_out/r2e_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 400 $NCU -k regex:"gemm_wgrad_kernel<256" --launch-skip 1 --launch-count 1 -o gpurun_out/r2e_ncu_wgrad_256x1 $RUN > gpurun_out/r2e_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 400 $NCU -k regex:"gemm_fwd_kernel<64, 3>" --launch-skip 1 --launch-count 1 -o gpurun_out/r2e_ncu_gemm_fwd_64x3 $RUN > gpurun_out/r2e_ncu7.log 2>&1; echo "ncu7 rc=$?"
timeout 400 $NCU -k regex:"bn_bwd_reduce_kernel" --launch-count 1 -o gpurun_out/r2e_ncu_bn_bwd_reduce $RUN > gpurun_out/r2e_ncu8.log 2>&1; echo "ncu8 rc=$?"
unset GDRN_PROFILE
ls -la gpurun_out/r2e*.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2e_bench.log
GDRN_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_mixed.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r2e_ncu.log 2>&1; echo "ncu list rc=$?"
