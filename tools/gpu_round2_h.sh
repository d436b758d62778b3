#!/bin/bash
# GPU call H (round 2): tests (deconv by phases, ROI targets), A/B, full bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2h_pytest.log
for v in 0 1; do
  echo "== quick bench mixed deconv_phases=$v"
  GDRN_DECONV_PHASES=$v timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.log 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2h_bench.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2h_smoke.log
