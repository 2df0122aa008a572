#!/bin/bash
# development sweep: bench.py under different env settings, one summary line each.  usage: tools/dev_sweep.sh <workload> <streams> "ENV1=.. ENV2=.." "ENV.." ...
mkdir -p gpurun_out
wl=$1; ns=$2; shift 2
i=0
for envs in "$@"; do
  i=$((i+1)); tag="sw_${wl}_$i"
  env ATZ_BENCH_NO_CPU=1 $envs python bench.py --steps 3 --warmup 3 --workload $wl --streams $ns > gpurun_out/$tag.log 2> gpurun_out/$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$tag.log").read().strip().splitlines()[-1])
    print("$wl [$envs]", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k[3:]:round(v,1) for k,v in d["phase_ms_per_step"].items()}, "trials", d["gpu_trials_per_step"], "streams", d["config"]["streams_per_gpu"], d["config"]["recompressed_per_gpu"])
except Exception as e: print("$wl [$envs] failed", e); print(open("gpurun_out/$tag.err").read()[-1500:])
PY
done
