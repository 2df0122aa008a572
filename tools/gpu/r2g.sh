mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2g_pytest.log; tail -5 gpurun_out/r2g_pytest.log
export ATZ_BENCH_NO_CPU=1
ATZ_DENSE_MINB=4 python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2g_c5_minb4.log 2> gpurun_out/r2g_c5_minb4.err
ATZ_DENSE_MINB=4 python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r2g_c3_minb4.log 2> gpurun_out/r2g_c3_minb4.err
unset ATZ_BENCH_NO_CPU
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2g_c5_1000.log 2> gpurun_out/r2g_c5_1000.err
tail -c 1500 gpurun_out/r2g_c5_1000.log; tail -5 gpurun_out/r2g_c5_1000.err
