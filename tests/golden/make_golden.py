"""Generates tests/golden/deflate_digests.json and inflate_vectors.json from the reference's own zlib 1.2.8
(oracle/_ref/libz128.so, compiled from /root/reference by oracle/build_ref.sh).  Run: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import zref  # noqa: E402
from test_oracle_deflate import _inputs  # noqa: E402
from test_oracle_inflate import INFCOVER_RAW, wrap  # noqa: E402

R = random.Random(99)
dig = {}
for name, d in _inputs().items():
    for lvl in range(10):
        for (w, m) in [(15, 8), (10, 1), (R.randint(10, 15), R.randint(1, 9))]:
            dig[f"{name}:{lvl}:{w}:{m}"] = hashlib.sha256(zref.ref_deflate(d, lvl, w, m)).hexdigest()
json.dump(dig, open(os.path.join(HERE, "deflate_digests.json"), "w"), indent=0, sort_keys=True)

vec = {}
for hexs, what in INFCOVER_RAW:
    z = wrap(hexs)
    first, ret, tin, tout, avail = zref.ref_inflate_scan(z, 0, 1 << 16)
    vec[hexs] = {"what": what, "ret": ret, "total_in": tin, "total_out": tout}
json.dump(vec, open(os.path.join(HERE, "inflate_vectors.json"), "w"), indent=1, sort_keys=True)
print(len(dig), "deflate digests,", len(vec), "inflate vectors")
