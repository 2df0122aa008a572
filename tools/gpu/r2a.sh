mkdir -p gpurun_out
set -x
export ATZ_BENCH_NO_CPU=1
python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2a_c3_base.log 2> gpurun_out/r2a_c3_base.err
ATZ_DEBUG_TRIALS=1 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2a_c3_dbg.log 2> gpurun_out/r2a_c3_dbg.err
ATZ_ALL_ROWS=1 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2a_c3_allrows.log 2> gpurun_out/r2a_c3_allrows.err
ATZ_ALL_ROWS=1 ATZ_DEBUG_TRIALS=1 python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r2a_c3_allrows_dbg.log 2> gpurun_out/r2a_c3_allrows_dbg.err
python bench.py --workload c5 --streams 128 --steps 1 --warmup 3 > gpurun_out/r2a_c5_base.log 2> gpurun_out/r2a_c5_base.err
ATZ_ALL_ROWS=1 python bench.py --workload c5 --streams 128 --steps 1 --warmup 3 > gpurun_out/r2a_c5_allrows.log 2> gpurun_out/r2a_c5_allrows.err
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt; nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
