#!/bin/bash
# GPU call D (round 2): ncu --set full captures of the dominant kernels of the mixed-mode step, test re-run, config benches
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2d_pytest.log
for v in 0 1; do
  echo "== quick bench mixed, wgrad side stream=$v"
  GDRN_WGRAD_STREAM=$v timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
echo "== quick bench half, wgrad side stream=1 / 0"
GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
GDRN_WGRAD_STREAM=0 GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
RUN="python bench.py --quick --no-graph --steps 1 --warmup 3"
export GDRN_PROFILE=1
timeout 400 $NCU -k regex:"gemm_fwd2_kernel<128" --launch-skip 35 --launch-count 1 -o gpurun_out/r2d_ncu_gemm_fwd2_128x3 $RUN > gpurun_out/r2d_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 400 $NCU -k regex:"gemm_fwd2_kernel<256" --launch-count 1 -o gpurun_out/r2d_ncu_gemm_fwd2_256x1 $RUN > gpurun_out/r2d_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 400 $NCU -k regex:"gemm_wgrad_kernel<256" --launch-skip 1 --launch-count 1 -o gpurun_out/r2d_ncu_wgrad_256x1 $RUN > gpurun_out/r2d_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 400 $NCU -k regex:"bn_fwd_kernel" --launch-count 1 -o gpurun_out/r2d_ncu_bn_fwd_stem $RUN > gpurun_out/r2d_ncu4.log 2>&1; echo "ncu4 rc=$?"
timeout 400 $NCU -k regex:"head_glue_fwd_kernel|pixel_loss_fwd_kernel" --launch-count 2 -o gpurun_out/r2d_ncu_pixel_fwd $RUN > gpurun_out/r2d_ncu5.log 2>&1; echo "ncu5 rc=$?"
timeout 400 $NCU -k regex:"head_bwd_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel" --launch-count 3 -o gpurun_out/r2d_ncu_head_bn_bwd $RUN > gpurun_out/r2d_ncu6.log 2>&1; echo "ncu6 rc=$?"
timeout 400 $NCU -k regex:"gemm_fwd_kernel<64, 3>" --launch-skip 1 --launch-count 1 -o gpurun_out/r2d_ncu_gemm_fwd_64x3 $RUN > gpurun_out/r2d_ncu7.log 2>&1; echo "ncu7 rc=$?"
unset GDRN_PROFILE
ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py --config ycbv --steps 20 --warmup 5 > gpurun_out/r2d_bench_ycbv.log 2> gpurun_out/r2d_bench_ycbv.err; echo "ycbv rc=$?"
timeout 300 python bench.py --config pnp --nin 67 --steps 20 --warmup 5 > gpurun_out/r2d_bench_pnp67.log 2> gpurun_out/r2d_bench_pnp67.err; echo "pnp67 rc=$?"
timeout 300 python bench.py --config pnp --nin 69 --steps 20 --warmup 5 > gpurun_out/r2d_bench_pnp69.log 2> gpurun_out/r2d_bench_pnp69.err; echo "pnp69 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2d_bench.log
