# ncu evidence for round 2 (B200_PROFILING.md recipe): launch list first, then --set full captures of the kernels DESIGN.md names.
# Every command below has exited 0 without ncu before (tools/gpu/r2*.sh); numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export ATZ_BENCH_NO_CPU=1
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2_launches_c5_128.csv python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
$NCU --set full --import-source on -k regex:deflate_trials_kernel -c 8 -f -o gpurun_out/r2_trials python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_trials.log 2>&1
$NCU --set full --import-source on -k regex:build_rows_kernel -c 5 -f -o gpurun_out/r2_rows python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_rows.log 2>&1
$NCU --set full --import-source on -k regex:inflate_kernel -c 3 -f -o gpurun_out/r2_inflate python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_inflate.log 2>&1
$NCU --set full --import-source on -k "regex:chain_.*_kernel" -c 6 -f -o gpurun_out/r2_chains python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_chains.log 2>&1
$NCU --set full --import-source on -k regex:scan_kernel -c 2 -f -o gpurun_out/r2_scan python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_scan.log 2>&1
cat > /tmp/diffdrv.py <<'PY'
import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import antiz_b200 as az, corpus
data = b"".join(corpus.imperfect(100 + i, i % 2 == 1) for i in range(6))
c = az.Context(0); c.load(data); c.scan(); c.search(az.Options(recompTresh=1000, sizediffTresh=1000, shortcutLength=4000))
print(sum(s.ndiff for s in c.streams()), "diff bytes in", sum(1 for s in c.streams() if s.ndiff), "streams")
PY
$NCU --set full --import-source on -k regex:diff_kernel -c 1 -f -o gpurun_out/r2_diff python /tmp/diffdrv.py > gpurun_out/r2_ncu_diff.log 2>&1
ls -la gpurun_out/*.ncu-rep
