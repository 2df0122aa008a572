#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 > gpurun_out/s13_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 1 -c 1 -o gpurun_out/s13_inflate -f python bench.py --steps 1 --warmup 3 > gpurun_out/s13_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:deflate_trials -c 2 -o gpurun_out/s13_trials -f python bench.py --steps 1 --warmup 3 > gpurun_out/s13_ncu2.log 2>&1
tail -2 gpurun_out/s13_ncu2.log
