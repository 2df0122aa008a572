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
NCU="ncu --clock-control none"
ATZ_BG_B=0 $NCU --set full --import-source on -k regex:deflate_trials_kernel --launch-skip 9 -c 3 -f -o /tmp/prof/r2_trials_1gb python bench.py --steps 1 --warmup 3 --no-c2 > gpurun_out/r2n_ncu_trials_1gb.log 2>&1
$S kernel /tmp/prof/r2_trials_1gb.ncu-rep gpurun_out/r2_deflate_trials_kernel_1gb.md
cat > /tmp/scandrv.py <<'PY'
import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import antiz_b200 as az, corpus
rng = np.random.default_rng(1)
data = rng.integers(0, 256, size=1 << 30, dtype=np.uint8)
blob = np.frombuffer(corpus.c2(20, 3), dtype=np.uint8)
data[1000:1000 + blob.size] = blob
c = az.Context(0); c.load(data)
prev = 0.0
for _ in range(3):
    n = c.scan(524288); st = c.stats()
    print(n, "streams,", st.n_candidates, "candidates, scan phase ms (K1 count + write + host round trip)", st.ms_scan - prev); prev = st.ms_scan
PY
python /tmp/scandrv.py > gpurun_out/r2_scan_1gb_plain.log 2>&1
$NCU --set full --import-source on -k "regex:scan_(count|write)_kernel" --launch-skip 2 -c 2 -f -o /tmp/prof/r2_scan python /tmp/scandrv.py > gpurun_out/r2n_ncu_scan.log 2>&1
$S kernel /tmp/prof/r2_scan.ncu-rep gpurun_out/r2_scan_kernels_1gb.md
du -sh gpurun_out
