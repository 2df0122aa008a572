# the -m gpu suite, smoke(), K1 timing and ncu capture on 1 GiB, a short default bench: the last build of round 2
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2p_pytest.log; tail -5 gpurun_out/r2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -2 gpurun_out/r2p_smoke.log
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
timeout 200 python /tmp/scandrv.py > gpurun_out/r2_scan_1gb_plain.log 2>&1; cat gpurun_out/r2_scan_1gb_plain.log
timeout 400 ncu --clock-control none --set full --import-source on -k "regex:scan_(count|write)_kernel" --launch-skip 2 -c 2 -f -o /tmp/prof/r2_scan python /tmp/scandrv.py > gpurun_out/r2p_ncu_scan.log 2>&1
timeout 120 python tools/dev_profile_summary.py kernel /tmp/prof/r2_scan.ncu-rep gpurun_out/r2_scan_kernels_1gb.md
grep -A14 "^kernel" gpurun_out/r2_scan_kernels_1gb.md | grep "kernel\|duration\|inst_executed.sum\|dram__bytes\|issue_active"
ATZ_BENCH_NO_CPU=1 timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/r2p_c2.log 2> gpurun_out/r2p_c2.err; tail -c 300 gpurun_out/r2p_c2.log
