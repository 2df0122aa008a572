"""Adversarial scan inputs against the reference binary: a file that is nothing but zlib headers, random bytes, and a valid
stream repeated back to back without gaps."""
import os, subprocess, sys, random, time
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(_R, "tests")); sys.path.insert(0, _R)
import corpus, zref
UNCOMP = os.path.join(_R, "antiz_b200", "uncomp")
R = random.Random(1)
z = zref.ref_deflate(corpus.text(30000, 3), 6, 15, 8)
cases = {"all_magic_1MB": b"\x78\x9c" * (1 << 19), "random_32MB": R.randbytes(32 << 20), "back_to_back": z * 200,
         "magic_then_stream": b"\x78\x9c" * 5000 + z + b"\x78\x01" * 3000 + z[:-1] + b"x"}
for name, data in cases.items():
    f = f"/dev/shm/adv_{name}.bin"; open(f, "wb").write(data)
    t0 = time.time(); a = subprocess.run([zref.REF_BIN, "-i", f, "-o", f + ".ref", "--notest"], capture_output=True, text=True); t1 = time.time()
    b = subprocess.run([UNCOMP, "-i", f, "-o", f + ".gpu", "--notest", "--stats"], capture_output=True, text=True); t2 = time.time()
    same = a.returncode == b.returncode == 0 and open(f + ".ref", "rb").read() == open(f + ".gpu", "rb").read()
    stats = [l for l in b.stderr.splitlines() if l.startswith("[gpu 0]")]
    print(name, len(data), "identical" if same else f"DIFFERENT rc {a.returncode} {b.returncode} {b.stdout[-200:]} {b.stderr[-300:]}", f"ref {t1-t0:.2f}s gpu {t2-t1:.2f}s", stats[0][:150] if stats else "")
