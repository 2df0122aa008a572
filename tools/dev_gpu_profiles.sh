#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/prof_bench.log 2> gpurun_out/prof_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/prof_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:deflate_trials -c 2 -o gpurun_out/prof_trials -f python bench.py --steps 1 --warmup 3 > gpurun_out/prof_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_rows -c 1 -o gpurun_out/prof_rows -f python bench.py --steps 1 --warmup 3 > gpurun_out/prof_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 1 -c 1 -o gpurun_out/prof_inflate -f python bench.py --steps 1 --warmup 3 > gpurun_out/prof_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_count -c 1 -o gpurun_out/prof_scan -f python bench.py --steps 1 --warmup 3 > gpurun_out/prof_ncu4.log 2>&1
tail -c 400 gpurun_out/prof_bench.log
