mkdir -p gpurun_out
export ATZ_BENCH_NO_CPU=1
ATZ_BURST=0 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2b_c3_noburst.log 2> gpurun_out/r2b_c3_noburst.err
ATZ_BURST=0 ATZ_ALL_ROWS=1 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2b_c3_noburst_allrows.log 2> gpurun_out/r2b_c3_noburst_allrows.err
ATZ_BURST=0 ATZ_ALL_ROWS=1 ATZ_DEBUG_TRIALS=1 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2b_c3_noburst_allrows_dbg.log 2> gpurun_out/r2b_c3_noburst_allrows_dbg.err
ATZ_ALL_ROWS=1 ATZ_DEBUG_LANES=1 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2b_c3_allrows_lanes.log 2> gpurun_out/r2b_c3_allrows_lanes.err
