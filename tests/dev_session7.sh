#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s7_pytest.log
python tests/dev_trial_cycles.py > gpurun_out/s7_cycles.log 2>&1; grep "rec=2" gpurun_out/s7_cycles.log | head -6
python bench.py --steps 2 --warmup 3 > gpurun_out/s7_bench_c2.log 2> gpurun_out/s7_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --steps 2 --warmup 3 --workload c4 --streams 20000 > gpurun_out/s7_bench_c4.log 2> gpurun_out/s7_bench_c4.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
for w in ("c2","c4"):
    try:
        d=json.loads(open(f"gpurun_out/s7_bench_{w}.log").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, "trials", d["ref_equivalent_trials_per_step"], d["gpu_trials_per_step"])
    except Exception as e:
        print(w, "failed", e)
PY
python bench.py --steps 1 --warmup 3 --streams 400 > gpurun_out/s7_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s7_launches.csv python bench.py --steps 1 --warmup 3 --streams 400 > gpurun_out/s7_ncu3.log 2>&1
