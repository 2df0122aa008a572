mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "small_budget or c5_mixed" 2>&1 | tail -3 > gpurun_out/final_pytest_subset.log; cat gpurun_out/final_pytest_subset.log
ATZ_BENCH_NO_CPU=1 timeout 500 python bench.py --steps 3 --warmup 3 --no-c2 > gpurun_out/final2_c5_1000.log 2> gpurun_out/final2_c5_1000.err
ATZ_BENCH_NO_CPU=1 timeout 300 python bench.py --streams 512 --steps 2 --warmup 3 > gpurun_out/final2_c5_512.log 2> gpurun_out/final2_c5_512.err
