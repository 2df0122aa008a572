import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(_R, "tests")); sys.path.insert(0, _R)
import antiz_b200 as az, corpus, zref
ctx = az.Context(0)
for name, plain in (("text110k", corpus.text(110000, 5)), ("bin110k", corpus.binaryish(110000, 6)), ("text1M", corpus.text(1 << 20, 7))):
    orig_f = zref.ref_deflate(plain, 6, 15, 8, 1)   # Z_FILTERED: nothing matches
    orig_m = zref.ref_deflate(plain, 6, 15, 8)
    for (lvl, w, m) in ((6, 15, 8), (9, 15, 9), (9, 15, 1), (1, 15, 8), (3, 15, 9)):
        for tag, orig in (("matched", orig_m),):
            for force in ("0", "2"):
                os.environ["ATZ_FORCE_REC"] = force
                r = ctx.trial(plain, orig, lvl, w, m, az.Options())
                print(f"{name} l{lvl} w{w} m{m} {tag} rec={force}: status {r.status} in_consumed {r.in_consumed} out {r.out_len} kcyc {r.kcycles} flush {r.kcycles_flush} -> {r.kcycles/1965:.2f} ms, {r.in_consumed/max(r.kcycles,1)/1.024*1.965:.2f} MB/s")
