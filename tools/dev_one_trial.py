import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(_R, "tests")); sys.path.insert(0, _R)
import antiz_b200 as az, corpus, zref
ctx = az.Context(0)
plain = corpus.text(110000, 5)
orig = zref.ref_deflate(plain, 6, 15, 8)
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 6
r = ctx.trial(plain, orig, lvl, 15, 8, az.Options())
print(r.status, r.in_consumed, r.kcycles, r.kcycles_flush)
