#!/bin/bash
# GPU call F (round 2): GEMM ncu --set full captures, tests of the GN cluster kernels / pack overlap, quick benches
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2f_pytest.log
for i in 1 2; do
  echo "== quick bench mixed"
  timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
echo "== quick bench half"
GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -f"
RUN="python bench.py --quick --no-graph --steps 1 --warmup 3"
export GDRN_PROFILE=1
timeout 400 $NCU -k regex:"gemm_fwd2_kernel.*128.*3" --launch-skip 35 --launch-count 1 -o gpurun_out/r2f_ncu_gemm_fwd2_128x3 $RUN > gpurun_out/r2f_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 400 $NCU -k regex:"gemm_fwd2_kernel.*256.*1" --launch-count 1 -o gpurun_out/r2f_ncu_gemm_fwd2_256x1 $RUN > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 400 $NCU -k regex:"gemm_wgrad_kernel.*256.*1" --launch-skip 1 --launch-count 1 -o gpurun_out/r2f_ncu_wgrad_256x1 $RUN > gpurun_out/r2f_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 400 $NCU -k regex:"gemm_fwd_kernel.*64.*3" --launch-skip 1 --launch-count 1 -o gpurun_out/r2f_ncu_gemm_fwd_64x3 $RUN > gpurun_out/r2f_ncu7.log 2>&1; echo "ncu7 rc=$?"
unset GDRN_PROFILE
ls -la gpurun_out/r2f*.ncu-rep
timeout 300 python tools/ablate_step.py 64 mixed > gpurun_out/r2f_ablate_mixed.txt 2>&1; echo "ablate rc=$?"
tail -12 gpurun_out/r2f_ablate_mixed.txt
