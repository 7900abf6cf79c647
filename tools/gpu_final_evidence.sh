#!/bin/bash
# final evidence: launch list of the default bench command, one full capture of the forward kernel
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_am64_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > $O/r2_ncu_final.log 2>&1
python tools/launch_summary.py $O/r2_launches_am64_final.csv 2>&1 | sort -t= -k4 -r | head -12
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_rowblock -c 1 -o $O/r2_rowblock_final_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords > /dev/null 2>&1
ls -la $O/r2_rowblock_final_full.ncu-rep
