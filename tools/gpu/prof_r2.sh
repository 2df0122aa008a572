# ncu evidence for round 2 (B200_PROFILING.md recipe): launch list first, then --set full captures of the kernels DESIGN.md names.
# Every command below has exited 0 without ncu before (tools/gpu/r2*.sh); numbers printed under ncu are never bench values.
# The .ncu-rep files stay on the box (/tmp/prof: gpurun copies at most 64 MiB back); their summaries go to gpurun_out/.
mkdir -p gpurun_out /tmp/prof
export ATZ_BENCH_NO_CPU=1
NCU="ncu --clock-control none"
S="python tools/dev_profile_summary.py"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2_launches_c5_128.csv python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
$S launches gpurun_out/r2_launches_c5_128.csv gpurun_out/r2_launches_c5_128.md
# the --brute-window launch (the 4th foreground launch of a step; with phase B in the foreground for a fixed launch order: the 7th)
ATZ_BG_B=0 $NCU --set full --import-source on -k regex:deflate_trials_kernel --launch-skip 6 -c 1 -f -o /tmp/prof/r2_trials_brute python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_trials.log 2>&1
$S kernel /tmp/prof/r2_trials_brute.ncu-rep gpurun_out/r2_deflate_trials_kernel_brute.md
ATZ_BG_B=0 $NCU --set full --import-source on -k regex:deflate_trials_kernel --launch-skip 5 -c 1 -f -o /tmp/prof/r2_trials_b python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_trials_b.log 2>&1
$S kernel /tmp/prof/r2_trials_b.ncu-rep gpurun_out/r2_deflate_trials_kernel_phaseB.md
$NCU --set full --import-source on -k regex:build_rows_kernel --launch-skip 2 -c 1 -f -o /tmp/prof/r2_rows python bench.py --streams 128 --steps 1 --warmup 3 > gpurun_out/r2_ncu_rows.log 2>&1
$S kernel /tmp/prof/r2_rows.ncu-rep gpurun_out/r2_build_rows_kernel.md
cat > /tmp/scandrv.py <<'PY'
import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import antiz_b200 as az, corpus
rng = np.random.default_rng(1)
data = rng.integers(0, 256, size=1 << 30, dtype=np.uint8)
blob = np.frombuffer(corpus.c2(20, 3), dtype=np.uint8)
data[1000:1000 + blob.size] = blob
c = az.Context(0); c.load(data)
for _ in range(3):
    n = c.scan(524288); st = c.stats()
    print(n, "streams,", st.n_candidates, "candidates, scan ms", st.ms_scan, "->", data.size / st.ms_scan / 1e6, "GB/s")
PY
python /tmp/scandrv.py > gpurun_out/r2_scan_1gb_plain.log 2>&1
$NCU --set full --import-source on -k regex:scan_kernel --launch-skip 1 -c 1 -f -o /tmp/prof/r2_scan python /tmp/scandrv.py > gpurun_out/r2_ncu_scan.log 2>&1
$S kernel /tmp/prof/r2_scan.ncu-rep gpurun_out/r2_scan_kernel_1gb.md
cat > /tmp/diffdrv.py <<'PY'
import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import antiz_b200 as az, corpus
data = b"".join(corpus.imperfect(100 + i, i % 2 == 1) for i in range(6))
c = az.Context(0); c.load(data); c.scan(); c.search(az.Options(recompTresh=1000, sizediffTresh=1000, shortcutLength=4000))
print(sum(s.ndiff for s in c.streams()), "diff bytes in", sum(1 for s in c.streams() if s.ndiff), "streams; ms_diff", c.stats().ms_diff)
PY
python /tmp/diffdrv.py > gpurun_out/r2_diff_plain.log 2>&1
$NCU --set full --import-source on -k regex:diff_kernel -c 1 -f -o /tmp/prof/r2_diff python /tmp/diffdrv.py > gpurun_out/r2_ncu_diff.log 2>&1
$S kernel /tmp/prof/r2_diff.ncu-rep gpurun_out/r2_diff_kernel.md
ls -la /tmp/prof gpurun_out | tail -30
du -sh gpurun_out
