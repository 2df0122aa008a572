import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import corpus
kind, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
d = {"c1": lambda: corpus.c1(), "c2": lambda: corpus.c2(n, 2), "c3": lambda: corpus.c3(n, 3), "c4": lambda: corpus.c4(n, 4)}[kind]()
open(out, "wb").write(d); print(kind, n, len(d))
