#!/bin/bash
# GPU call W (round 2): last check of the rebuilt library (launch macro in gdrn_internal.h): ROI + RANSAC-PnP tests, smoke
set -x
timeout 150 python -m pytest tests/test_ops_gpu.py tests/test_pnp_ransac_gpu.py -m gpu -x -q -k "roi or pnp_ransac or pdl" 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
