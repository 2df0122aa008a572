mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "strategies" 2>&1 | tail -15 | cut -c1-600 > gpurun_out/dbg5_k.log; cat gpurun_out/dbg5_k.log
python - <<'PY' > gpurun_out/dbg5.log 2>&1
import sys, random
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import corpus, zref, antiz_b200 as az
r = random.Random(77)
ss = []; meta = []
for i in range(24):
    d = corpus.binaryish(r.randint(3000, 60000), 700 + i) if i % 2 else corpus.text(r.randint(3000, 90000), 700 + i, 300)
    strat = [1, 2, 3, 4][i % 4]
    lvl = r.randint(4, 9) if strat == 1 else r.randint(1, 9)
    w = r.choice([12, 15]); m = r.choice([8, 9, 4])
    ss.append(zref.ref_deflate(d, lvl, w, m, strat)); meta.append((strat, lvl, w, m, len(d)))
data, offs = corpus.container(ss, 78)
c = az.Context(0); c.load(data); n = c.scan(); c.search(az.Options(flags=az.ATZ_F_STRATEGIES | az.ATZ_F_EXACT_RECORDS))
got = {s.offset: s for s in c.streams()}
for o, z, mt in zip(offs, ss, meta):
    s = got.get(o)
    print(mt, 'C', len(z), 'found' if s else 'NOT FOUND', (s.recomp, s.clevel & 15, s.clevel >> 4, s.window, s.memlevel, s.identBytes, s.streamLength) if s else '')
PY
cat gpurun_out/dbg5.log
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import antiz_b200 as az, corpus
rng = np.random.default_rng(1)
data = rng.integers(0, 256, size=1 << 30, dtype=np.uint8)
blob = np.frombuffer(corpus.c2(20, 3), dtype=np.uint8)
data[1000:1000 + blob.size] = blob
c = az.Context(0); c.load(data)
for _ in range(3):
    n = c.scan(524288); st = c.stats()
    print(n, "streams,", st.n_candidates, "candidates, scan ms", st.ms_scan, "->", data.size / st.ms_scan / 1e6, "GB/s")
PY
