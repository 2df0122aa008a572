#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s10_pytest.log
for w in "c2 0" "c4 20000"; do set -- $w; python bench.py --steps 2 --warmup 3 --workload $1 --streams $2 > gpurun_out/s10_bench_$1.log 2> gpurun_out/s10_bench_$1.err; echo "bench $1 rc=$?"; done
python - <<'PY'
import json
for w in ("c2","c4"):
    try:
        d=json.loads(open(f"gpurun_out/s10_bench_{w}.log").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, "trials", d["ref_equivalent_trials_per_step"], d["gpu_trials_per_step"])
    except Exception as e:
        print(w, "failed", e)
PY
