mkdir -p gpurun_out
g++ -std=c++17 -O2 -I include -I antiz_b200/host tests/zlibwrapper_scan.cpp -o /tmp/zws -L antiz_b200 -lantiz_b200 -Wl,-rpath,$PWD/antiz_b200
python - <<'PY' > gpurun_out/dbg3.log 2>&1
import sys, subprocess
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import corpus, antiz_b200 as az
data = corpus.c2(10, 55, 1 << 10, 40 << 10) + corpus.c4(40, 56) + corpus.junk(5000, 57)
open('/dev/shm/zw.bin','wb').write(data)
for cs in (20000,):
    out = subprocess.run(['/tmp/zws','/dev/shm/zw.bin',str(cs)],capture_output=True,text=True)
    got=[tuple(int(x) for x in l.split()) for l in out.stdout.splitlines()]
    c=az.Context(0); c.load(data); c.scan(cs)
    want=[(s.offset,s.offsetType,s.streamLength,s.inflatedLength) for s in c.streams()]
    print('got ',got[:8]); print('want',want[:8]); print(out.stderr[-500:])
PY
cat gpurun_out/dbg3.log
