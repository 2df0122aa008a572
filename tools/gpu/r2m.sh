mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2m_pytest.log; tail -5 gpurun_out/r2m_pytest.log
export ATZ_BENCH_NO_CPU=1
python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2m_c5_128.log 2> gpurun_out/r2m_c5_128.err
python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r2m_c3.log 2> gpurun_out/r2m_c3.err
python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/r2m_c2.log 2> gpurun_out/r2m_c2.err
ATZ_DEBUG_TRIALS=1 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2m_c3_dbg.log 2> gpurun_out/r2m_c3_dbg.err
