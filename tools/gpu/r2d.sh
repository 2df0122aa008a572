mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2d_smi.txt
python -m pytest tests -m gpu -x -q -k "uncomp_gpus or sharded" 2>&1 | tail -8 > gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --streams 256 --steps 2 --warmup 3 > gpurun_out/r2d_bench_n2.log 2> gpurun_out/r2d_bench_n2.err
tail -c 2500 gpurun_out/r2d_bench_n2.log; tail -5 gpurun_out/r2d_bench_n2.err
ATZ_BENCH_NO_CPU=1 python bench.py --gpus 1 --streams 256 --steps 2 --warmup 3 > gpurun_out/r2d_bench_n1.log 2> gpurun_out/r2d_bench_n1.err
tail -c 1500 gpurun_out/r2d_bench_n1.log
ATZ_BENCH_NO_CPU=1 ATZ_DEBUG_TRIALS=1 ATZ_DEBUG_LANES=1 python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2d_c5_dbg.log 2> gpurun_out/r2d_c5_dbg.err
