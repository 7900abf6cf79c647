#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
echo "== sampling tests"; timeout 60 python -m pytest tests/test_gpu_sampling.py -q -x -k "pick_for_pick" 2>&1 | tail -4 | tee $O/test_side.log
