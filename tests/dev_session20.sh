#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; wl=$2; ns=$3; shift 3; env "$@" python bench.py --steps 2 --warmup 3 --workload $wl --streams $ns > gpurun_out/s20_$tag.log 2> gpurun_out/s20_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s20_$tag.log").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, d["gpu_trials_per_step"])
except Exception as e: print("$tag failed", e)
PY
}
run c3_sparse c3 120 ATZ_DENSE=0
run c3_dense c3 120 ATZ_DENSE=1
run c2_sparse c2 0 ATZ_DENSE=0
