#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -3
echo "== sweep"; timeout 900 python tools/fused_sweep.py 0 640/f 640 2>&1 | tail -4 | cut -c1-150 | tee $O/r2o_sweep.log
