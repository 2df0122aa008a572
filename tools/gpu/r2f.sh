mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2f_pytest.log; tail -12 gpurun_out/r2f_pytest.log
export ATZ_BENCH_NO_CPU=1
python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r2f_c3.log 2> gpurun_out/r2f_c3.err
python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2f_c5.log 2> gpurun_out/r2f_c5.err
python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/r2f_c2.log 2> gpurun_out/r2f_c2.err
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/r2f_c4.log 2> gpurun_out/r2f_c4.err
python bench.py --workload c1 --steps 3 --warmup 3 > gpurun_out/r2f_c1.log 2> gpurun_out/r2f_c1.err
ATZ_DEBUG_TRIALS=1 ATZ_DEBUG_LANES=1 python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2f_c5_dbg.log 2> gpurun_out/r2f_c5_dbg.err
