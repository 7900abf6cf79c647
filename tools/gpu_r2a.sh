#!/bin/bash
# round 2, call A: fused row-block kernel (TMA gather4): parity tests, A/B sweep, launch list, one full ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -40 | tee $O/r2a_test_fused.log
echo "== sweep"; timeout 600 python tools/fused_sweep.py 0 640 640/bulk 512 640/s4 320/c2 2>&1 | tail -12 | tee $O/r2a_sweep.log
echo "== bench"; timeout 400 python bench.py --no-cpu-baseline > $O/r2a_bench_fused.json 2> $O/r2a_bench_fused.err; tail -c 1500 $O/r2a_bench_fused.json
echo "== ncu"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches_am64.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2a_ncu_b.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_rowblock -s 2 -c 2 -f -o $O/r2a_rowblock_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2a_ncu_full.log 2>&1
timeout 120 ncu -i $O/r2a_rowblock_full.ncu-rep --page raw --csv > $O/r2a_rowblock_full_raw.csv 2>/dev/null
timeout 120 ncu -i $O/r2a_rowblock_full.ncu-rep --page source --csv > $O/r2a_rowblock_full_source.csv 2>/dev/null
ls -la $O | tail -20
