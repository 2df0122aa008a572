#!/bin/bash
mkdir -p gpurun_out
python tests/dev_make_corpus.py c2 1 /dev/shm/one.bin >/dev/null
python - <<'PY'
import sys; sys.path.insert(0,'tests')
import corpus, zref
d = corpus.text(110000, 5)
open('/dev/shm/f1.bin','wb').write(corpus.container([zref.ref_deflate(d, 1, 15, 8)], 3)[0])
PY
export ATZ_FORCE_REC=2
python tests/dev_one_trial.py 6 > gpurun_out/s3_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:deflate_trials -c 1 -o gpurun_out/s3_trial_slow -f python tests/dev_one_trial.py 6 > gpurun_out/s3_ncu1.log 2>&1
unset ATZ_FORCE_REC
antiz_b200/uncomp -i /dev/shm/f1.bin --notest > gpurun_out/s3_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:deflate_trials -c 1 -o gpurun_out/s3_trial_fast -f antiz_b200/uncomp -i /dev/shm/f1.bin --notest > gpurun_out/s3_ncu2.log 2>&1
python bench.py --steps 1 --warmup 3 --streams 400 > gpurun_out/s3_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 1 --warmup 3 --streams 400 > gpurun_out/s3_ncu3.log 2>&1
tail -3 gpurun_out/s3_plain2.log
