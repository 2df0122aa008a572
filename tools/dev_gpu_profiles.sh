#!/bin/bash
# one profiling session: plain bench first (exit 0), then the ncu launch list, then one --set full capture per hot kernel
mkdir -p gpurun_out
T=${1:-r1f}
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err || exit 1
export ATZ_BENCH_NO_CPU=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/${T}_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:deflate_trials -c 2 -o gpurun_out/${T}_trials -f python bench.py --steps 1 --warmup 3 > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_rows -c 1 -o gpurun_out/${T}_rows -f python bench.py --steps 1 --warmup 3 > gpurun_out/${T}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 1 -c 1 -o gpurun_out/${T}_inflate -f python bench.py --steps 1 --warmup 3 > gpurun_out/${T}_ncu3.log 2>&1
tail -c 600 gpurun_out/${T}_bench.log
