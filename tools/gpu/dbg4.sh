mkdir -p gpurun_out
g++ -std=c++17 -O2 -I include -I antiz_b200/host tools/gpu/tmp/zws_dbg.cpp -o /tmp/zwsd -L antiz_b200 -lantiz_b200 -Wl,-rpath,$PWD/antiz_b200
python - <<'PY' > gpurun_out/dbg4.log 2>&1
import sys, subprocess
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import corpus, zref
data = corpus.c2(10, 55, 1 << 10, 40 << 10) + corpus.c4(40, 56) + corpus.junk(5000, 57)
open('/dev/shm/zw.bin','wb').write(data)
# the stream that starts between 14303 and 19999
import re
cands=[i for i in range(14303,19999) if data[i] in (0x78,) and ((data[i]<<8)|data[i+1])%31==0 and not data[i+1]&0x20]
print('cands',cands[:5])
for c in cands[:2]:
    print(subprocess.run(['/tmp/zwsd','/dev/shm/zw.bin',str(c)],capture_output=True,text=True).stdout)
    print('ref scan:', zref.ref_inflate_scan(data[:20000], c, 20000))
PY
cat gpurun_out/dbg4.log
