#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core_gemm" 2>&1 | tail -12
timeout 300 python tools/umma_vs_fma.py 10000000 2>&1 | tail -3 | tee $O/r2w_umma_vs_fma.json
