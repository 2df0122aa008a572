#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/s6_bench_c2.log 2> gpurun_out/s6_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --steps 2 --warmup 3 --workload c4 --streams 20000 > gpurun_out/s6_bench_c4.log 2> gpurun_out/s6_bench_c4.err; echo "bench c4 rc=$?"
python bench.py --steps 2 --warmup 3 --workload c3 --streams 120 > gpurun_out/s6_bench_c3.log 2> gpurun_out/s6_bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for w in ("c2","c4","c3"):
    try:
        d=json.loads(open(f"gpurun_out/s6_bench_{w}.log").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, "trials", d["ref_equivalent_trials_per_step"], d["gpu_trials_per_step"])
    except Exception as e:
        print(w, "failed", e)
PY
export ATZ_FORCE_REC=2
python tests/dev_one_trial.py 6 > gpurun_out/s6_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:deflate_trials -c 1 -o gpurun_out/s6_trial_res -f python tests/dev_one_trial.py 6 > gpurun_out/s6_ncu1.log 2>&1
cat gpurun_out/s6_plain1.log
