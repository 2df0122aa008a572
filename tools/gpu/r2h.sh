mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2h_pytest.log; tail -5 gpurun_out/r2h_pytest.log
export ATZ_BENCH_NO_CPU=1
python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2h_c5_128.log 2> gpurun_out/r2h_c5_128.err
python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r2h_c3.log 2> gpurun_out/r2h_c3.err
python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/r2h_c2.log 2> gpurun_out/r2h_c2.err
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/r2h_c4.log 2> gpurun_out/r2h_c4.err
python bench.py --steps 2 --warmup 3 > gpurun_out/r2h_c5_1000.log 2> gpurun_out/r2h_c5_1000.err
ATZ_TIERS=1 python bench.py --steps 2 --warmup 3 > gpurun_out/r2h_c5_1000_t1.log 2> gpurun_out/r2h_c5_1000_t1.err
tail -3 gpurun_out/r2h_*.err
