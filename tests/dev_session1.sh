#!/bin/bash
# dev GPU session: tests, full-size bench lines, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 3 --warmup 3 > gpurun_out/s1_bench_c2.log 2> gpurun_out/s1_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --steps 2 --warmup 3 --workload c3 --streams 120 > gpurun_out/s1_bench_c3.log 2> gpurun_out/s1_bench_c3.err; echo "bench c3 rc=$?"
python bench.py --steps 2 --warmup 3 --workload c4 --streams 20000 > gpurun_out/s1_bench_c4.log 2> gpurun_out/s1_bench_c4.err; echo "bench c4 rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s1_ref_c2.log 2>&1; echo "ref rc=$?"
tail -c 3000 gpurun_out/s1_bench_c2.log
