#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 > gpurun_out/s12_$tag.log 2> gpurun_out/s12_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s12_$tag.log").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()})
except Exception as e: print("$tag failed", e)
PY
}
run default X=1
run inf8 ATZ_INFLATE_WARPS=8
run inf12 ATZ_INFLATE_WARPS=12
run sparse ATZ_DENSE=0
