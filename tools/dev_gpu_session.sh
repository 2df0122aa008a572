#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/dev_pytest.log
run() { tag=$1; wl=$2; ns=$3; shift 3; env "$@" python bench.py --steps 2 --warmup 3 --workload $wl --streams $ns > gpurun_out/dev_$tag.log 2> gpurun_out/dev_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dev_$tag.log").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, d["gpu_trials_per_step"])
except Exception as e: print("$tag failed", e)
PY
}
run c2 c2 0 X=1
run c3 c3 120 X=1
run c5 c5 48 X=1
