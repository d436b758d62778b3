#!/bin/bash
# GPU call O (round 2): PDL variants in situ.  A = elementwise kernels trigger their dependents at entry, B = no explicit trigger
# (library built with -DGDRN_PDL_EW_NO_TRIGGER=1); mode 0 off, 1 GEMM kernels launched with PDL, 2 elementwise kernels as well
set -x
mkdir -p gpurun_out
B=$PWD/gdr_net_b200/lib/libgdrn_b200_ewnotrig.so
run() { # tag, mode, lib
  if [ -n "$3" ]; then export GDRN_LIB_PATH=$3; else unset GDRN_LIB_PATH; fi
  GDRN_PDL=$2 timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed quick $1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for rep in 1 2; do
  run A0 0 ""
  run A1 1 ""
  run A2 2 ""
  run B1 1 $B
  run B2 2 $B
done
unset GDRN_LIB_PATH
for v in 0 1 2; do
  GDRN_PDL=$v GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half quick A$v', d['value'], d['ms_per_step'])"
done
export GDRN_LIB_PATH=$B
for v in 1 2; do
  GDRN_PDL=$v GDRN_BENCH_MODE=half timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('half quick B$v', d['value'], d['ms_per_step'])"
done
