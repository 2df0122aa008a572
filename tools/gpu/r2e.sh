mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_pytest.log
export ATZ_BENCH_NO_CPU=1
python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r2e_c3.log 2> gpurun_out/r2e_c3.err
python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2e_c5.log 2> gpurun_out/r2e_c5.err
ATZ_BG_B=0 python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2e_c5_fg.log 2> gpurun_out/r2e_c5_fg.err
ATZ_TRIAL_ORDER=1 python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2e_c5_ord1.log 2> gpurun_out/r2e_c5_ord1.err
ATZ_DEBUG_TRIALS=1 ATZ_DEBUG_LANES=1 python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2e_c5_dbg.log 2> gpurun_out/r2e_c5_dbg.err
python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/r2e_c2.log 2> gpurun_out/r2e_c2.err
ATZ_WALK_BURST=0 python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r2e_c3_nowb.log 2> gpurun_out/r2e_c3_nowb.err
ATZ_WAVE_GROWTH=16 python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2e_c5_g16.log 2> gpurun_out/r2e_c5_g16.err
for f in r2e_c3 r2e_c5 r2e_c5_fg r2e_c5_ord1 r2e_c2; do tail -c 600 gpurun_out/$f.log | head -c 400; echo; tail -3 gpurun_out/$f.err; done
