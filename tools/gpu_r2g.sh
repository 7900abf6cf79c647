#!/bin/bash
# round 2, call C: software-pipelined fused row-block kernel: parity tests, A/B sweep, one full ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -5 | tee $O/r2g_test_fused.log
echo "== sweep"; timeout 900 python tools/fused_sweep.py 0 640 640/rdv 576 2>&1 | tail -12 | tee $O/r2g_sweep.log
echo "== ncu"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_rowblock -s 2 -c 1 -f -o $O/r2g_rowblock_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2g_ncu_full.log 2>&1
timeout 120 ncu -i $O/r2g_rowblock_full.ncu-rep --page raw --csv > $O/r2g_rowblock_full_raw.csv 2>/dev/null
timeout 120 ncu -i $O/r2g_rowblock_full.ncu-rep --page source --csv > $O/r2g_rowblock_full_source.csv 2>/dev/null
rm -f $O/r2g_rowblock_full.ncu-rep
