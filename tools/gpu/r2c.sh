mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2c_bench_c5_128.log 2> gpurun_out/r2c_bench_c5_128.err
tail -c 3000 gpurun_out/r2c_bench_c5_128.log; tail -5 gpurun_out/r2c_bench_c5_128.err
python bench.py --impl reference --streams 128 --steps 1 --warmup 0 > gpurun_out/r2c_ref_c5_128.log 2> gpurun_out/r2c_ref_c5_128.err
tail -c 1500 gpurun_out/r2c_ref_c5_128.log; tail -5 gpurun_out/r2c_ref_c5_128.err
