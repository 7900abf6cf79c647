#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== fused tests (2 turns)"; timeout 900 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -3
echo "== sweep"; timeout 900 python tools/fused_sweep.py 0 640/f 640/f/t3 2>&1 | tail -8 | cut -c1-150 | tee $O/r2n_sweep.log
echo "== ncu"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_rowblock -s 2 -c 1 -f -o $O/r2n_rowblock_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > $O/r2n_ncu_full.log 2>&1
timeout 120 ncu -i $O/r2n_rowblock_full.ncu-rep --page raw --csv > $O/r2n_rowblock_full_raw.csv 2>/dev/null
timeout 120 ncu -i $O/r2n_rowblock_full.ncu-rep --page source --csv > $O/r2n_rowblock_full_source.csv 2>/dev/null
rm -f $O/r2n_rowblock_full.ncu-rep
