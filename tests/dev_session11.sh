#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s11_pytest.log
python bench.py --steps 2 --warmup 3 > gpurun_out/s11_bench_c2.log 2> gpurun_out/s11_bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for w in ("c2",):
    d=json.loads(open(f"gpurun_out/s11_bench_{w}.log").read().strip().splitlines()[-1])
    print(w, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, "trials", d["ref_equivalent_trials_per_step"], d["gpu_trials_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s11_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/s11_ncu.log 2>&1
