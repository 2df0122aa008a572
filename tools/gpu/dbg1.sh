mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import corpus
open('/dev/shm/c1.bin','wb').write(corpus.c1())
open('/dev/shm/c2s.bin','wb').write(corpus.c2(12, 2))
PY
for v in "" "ATZ_BG_B=0" "ATZ_BURST=0" "ATZ_BG_B=0 ATZ_BURST=0"; do
  echo "== env: $v"; env $v ./antiz_b200/uncomp -i /dev/shm/c1.bin -o /dev/shm/c1.atz --notest 2>&1 | tail -3
  echo "== c2s env: $v"; env $v ./antiz_b200/uncomp -i /dev/shm/c2s.bin -o /dev/shm/c2s.atz --notest 2>&1 | tail -3
done > gpurun_out/dbg1.log 2>&1
timeout 600 compute-sanitizer --tool memcheck ./antiz_b200/uncomp -i /dev/shm/c1.bin -o /dev/shm/c1.atz --notest > gpurun_out/dbg1_san.log 2>&1
head -60 gpurun_out/dbg1_san.log
cat gpurun_out/dbg1.log
