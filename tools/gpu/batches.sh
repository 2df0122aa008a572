mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/batches_pytest.log; tail -5 gpurun_out/batches_pytest.log
export ATZ_BENCH_NO_CPU=1
timeout 300 python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/batches_c5_128.log 2> gpurun_out/batches_c5_128.err
timeout 300 python bench.py --streams 256 --steps 2 --warmup 3 > gpurun_out/batches_c5_256.log 2> gpurun_out/batches_c5_256.err
timeout 300 python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/batches_c3.log 2> gpurun_out/batches_c3.err
