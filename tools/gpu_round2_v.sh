#!/bin/bash
# GPU call V (round 2): ncu launch list (durations + tensor pipe) of one step of the FINAL build, for profiles/
set -x
mkdir -p gpurun_out
GDRN_PROFILE=1 timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2v_launches_final.csv python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/r2v_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py gpurun_out/r2v_launches_final.csv 2>/dev/null | head -12
