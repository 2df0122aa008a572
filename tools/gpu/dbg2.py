import ctypes, os, sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import antiz_b200 as az
if len(sys.argv) > 1 and sys.argv[1] != "-":
    az.LIB_PATH = os.path.abspath(sys.argv[1])
import corpus
which = sys.argv[2] if len(sys.argv) > 2 else "c1"
data = corpus.c1() if which == "c1" else corpus.c2(12, 2)
try:
    c = az.Context(0); c.load(data); n = c.scan(); c.search(az.Options())
    print(sys.argv[1:], "OK", n, sum(s.recomp for s in c.streams()))
except Exception as e:
    print(sys.argv[1:], "FAIL", str(e)[-120:])
