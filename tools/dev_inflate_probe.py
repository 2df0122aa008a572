import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import corpus, zref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def run(tag, streams):
    data, _ = corpus.container(streams, 5)
    f = "/dev/shm/ip.bin"; open(f, "wb").write(data)
    r = subprocess.run([os.path.join(ROOT, "antiz_b200", "uncomp"), "-i", f, "--notest", "--stats"], capture_output=True, text=True)
    line = [l for l in r.stderr.splitlines() if l.startswith("[gpu 0]")]
    print(tag, line[0].split("| ms:")[1].split("| algo")[0] if line else r.stderr[-300:])
d = corpus.text(256 << 10, 77)
for lvl in (1, 3, 6, 9):
    run(f"1 x 256KB level {lvl}:", [zref.ref_deflate(d, lvl, 15, 8)])
run("64 x 256KB level 6:", [zref.ref_deflate(corpus.text(256 << 10, 100 + i), 6, 15, 8) for i in range(64)])
run("1500 x 256KB level 6:", [zref.ref_deflate(corpus.text(256 << 10, 100 + i % 50), 6, 15, 8) for i in range(1500)])
b = corpus.binaryish(256 << 10, 5)
run("1 x 256KB binaryish level 6:", [zref.ref_deflate(b, 6, 15, 8)])
