mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench
path, data, parts = bench.shared_container("c5", 32, bench.SEED, leader=True)
print(path, len(data))
path, data, parts = bench.shared_container("c2", 2000, bench.SEED, leader=True)
print(path, len(data))
PY
for w in c5_32 c2_2000; do
F=/dev/shm/atz_bench_${w}_2.bin
FL=""; if [ $w = c5_32 ]; then FL="--brute-window"; fi
( time ./antiz_b200/uncomp -i $F -o $F.atz --notest --stats $FL ) > gpurun_out/r2j_${w}_pre.log 2>&1
( time ./antiz_b200/uncomp -r -i $F.atz -o $F.rec --stats ) > gpurun_out/r2j_${w}_rec.log 2>&1
( time ATZ_DEBUG_TRIALS=1 ./antiz_b200/uncomp -r -i $F.atz -o $F.rec --stats ) > gpurun_out/r2j_${w}_rec_dbg.log 2>&1
cmp $F $F.rec && echo roundtrip ok >> gpurun_out/r2j_${w}_rec.log
( time ./oracle/_ref/uncomp_ref -r -i $F.atz -o $F.rec2 ) > gpurun_out/r2j_${w}_refrec.log 2>&1
tail -12 gpurun_out/r2j_${w}_pre.log; tail -8 gpurun_out/r2j_${w}_rec.log; tail -4 gpurun_out/r2j_${w}_refrec.log
done
