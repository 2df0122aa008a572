#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s4_pytest.log
python bench.py --steps 2 --warmup 3 > gpurun_out/s4_bench_c2.log 2> gpurun_out/s4_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --steps 2 --warmup 3 --workload c4 --streams 20000 > gpurun_out/s4_bench_c4.log 2> gpurun_out/s4_bench_c4.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for w in ("c2","c4"):
    try:
        d=json.loads(open(f"gpurun_out/s4_bench_{w}.log").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, "trials", d["ref_equivalent_trials_per_step"], d["gpu_trials_per_step"])
    except Exception as e:
        print(w, "failed", e)
PY
python bench.py --steps 1 --warmup 3 --streams 300 > gpurun_out/s4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:build_chains -s 3 -c 1 -o gpurun_out/s4_chains -f python bench.py --steps 1 --warmup 3 --streams 300 > gpurun_out/s4_ncu1.log 2>&1
