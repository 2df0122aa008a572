mkdir -p gpurun_out
export ATZ_BENCH_NO_CPU=1
for t in 1 2; do ATZ_TIERS=$t timeout 400 python bench.py --streams 256 --steps 2 --warmup 3 > gpurun_out/tiers_c5_256_t$t.log 2> gpurun_out/tiers_c5_256_t$t.err; done
for t in 1 2; do ATZ_TIERS=$t timeout 400 python bench.py --streams 512 --steps 2 --warmup 3 > gpurun_out/tiers_c5_512_t$t.log 2> gpurun_out/tiers_c5_512_t$t.err; done
