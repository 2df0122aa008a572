mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2i_ngpu.txt; nproc >> gpurun_out/r2i_ngpu.txt
python -m pytest tests -m gpu -x -q -k "uncomp_gpus" 2>&1 | tail -5 > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-c2 > gpurun_out/r2i_bench_n$n.log 2> gpurun_out/r2i_bench_n$n.err
tail -c 900 gpurun_out/r2i_bench_n$n.log | head -c 700; echo; tail -3 gpurun_out/r2i_bench_n$n.err
done
PARITY_GPUS=8 python tools/fullsize_parity.py c5:32 > gpurun_out/r2i_parity_c5_32_8gpu.jsonl 2> gpurun_out/r2i_parity.err
cat gpurun_out/r2i_parity_c5_32_8gpu.jsonl | cut -c1-900
