"""Developer run on a GPU box: whole-file parity of antiz_b200/uncomp against oracle/_ref/uncomp_ref."""
import os, sys, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import corpus, zref

UNCOMP = os.path.join(ROOT, "antiz_b200", "uncomp")


def run(cmd):
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True)
    return p.returncode, p.stdout, p.stderr, time.time() - t0


def compare(name, data, flags, tmp):
    f = os.path.join(tmp, name + ".bin")
    open(f, "wb").write(data)
    rc1, o1, e1, t1 = run([zref.REF_BIN, "-i", f, "-o", f + ".ref.atz", "--notest"] + flags)
    rc2, o2, e2, t2 = run([UNCOMP, "-i", f, "-o", f + ".gpu.atz", "--stats"] + flags)
    a = open(f + ".ref.atz", "rb").read() if os.path.exists(f + ".ref.atz") else None
    b = open(f + ".gpu.atz", "rb").read() if os.path.exists(f + ".gpu.atz") else None
    same = a is not None and a == b
    l1 = [l for l in o1.splitlines() if l.startswith(("Total zlib", "recompressed", "Total bytes"))]
    l2 = [l for l in o2.splitlines() if l.startswith(("Total zlib", "recompressed", "Total bytes", "Testing"))]
    print(f"{name} {flags}: N={len(data)} ref {t1:.2f}s rc{rc1} | gpu {t2:.2f}s rc{rc2} | atz identical: {same}")
    print("   ref:", l1); print("   gpu:", l2)
    if e2.strip(): print("   gpu stderr:", e2.strip()[-600:])
    if not same and a and b:
        k = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), -1)
        print("   first diff at", k, "sizes", len(a), len(b))
    # reconstruct with our binary from the reference's atz
    if a is not None:
        rc3, o3, e3, t3 = run([UNCOMP, "-r", "-i", f + ".ref.atz", "-o", f + ".rec"])
        ok = os.path.exists(f + ".rec") and open(f + ".rec", "rb").read() == data
        print(f"   reconstruct(ref atz) {t3:.2f}s ok={ok}")
        same = same and ok
    return same


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    ok = True
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        if which == "small":
            ok &= compare("c1", corpus.c1(), [], tmp)
            ok &= compare("c2s", corpus.c2(60, 2), [], tmp)
            ok &= compare("c2chunk", corpus.c2(60, 22), ["--chunksize", "65536"], tmp)
            ok &= compare("c3s", corpus.c3(12, 3), ["--brute-window"], tmp)
            ok &= compare("c4s", corpus.c4(1500, 4), [], tmp)
            ok &= compare("c4tol", corpus.c4(300, 44), ["--mismatch-tol", "0", "--shortcut-len", "256", "--recomp-tresh", "16"], tmp)
        elif which == "c3":
            ok &= compare("c3", corpus.c3(25, 3), ["--brute-window"], tmp)
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
