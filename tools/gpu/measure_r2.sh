mkdir -p gpurun_out /tmp/prof
S="python tools/dev_profile_summary.py"
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2n_bench_default.log 2> gpurun_out/r2n_bench_default.err
tail -c 1200 gpurun_out/r2n_bench_default.log; tail -4 gpurun_out/r2n_bench_default.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2n_ref.log 2> gpurun_out/r2n_ref.err
tail -c 900 gpurun_out/r2n_ref.log
export ATZ_BENCH_NO_CPU=1
ATZ_LANES=2 python bench.py --streams 128 --steps 2 --warmup 3 > gpurun_out/r2n_c5_128_lanes2.log 2> gpurun_out/r2n_c5_128_lanes2.err
python tools/fullsize_parity.py c1 c2 c3 c4 > gpurun_out/r2n_fullsize_parity.jsonl 2> gpurun_out/r2n_fullsize_parity.err
cut -c1-400 gpurun_out/r2n_fullsize_parity.jsonl
# (a --set full capture of the trial kernel at 1 GB was tried here and ran into the box's time limit: profiles are taken on 128 MB, tools/gpu/prof_r2.sh)
