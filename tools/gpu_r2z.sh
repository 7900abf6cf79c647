#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in 1 0 1 0; do
RGCN_FUSED_NARROW=$v timeout 300 python bench.py --no-subrecords --no-cpu-baseline 2> $O/r2z.err | python tools/benchline.py
done
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -2
