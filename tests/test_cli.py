"""The `uncomp` host program's argument handling and ATZ1 reader errors (no GPU needed for these paths)."""
import os
import struct
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNCOMP = os.path.join(ROOT, "antiz_b200", "uncomp")


def run(*a):
    return subprocess.run([UNCOMP] + list(a), capture_output=True, text=True)


def test_missing_input_is_a_parse_error():
    p = run()
    assert p.returncode == 1 and "AntiZ 0.1.6-git" in p.stdout
    assert "PARSE ERROR" in p.stderr and "Required argument missing: input" in p.stderr


def test_unknown_flag_and_version():
    assert run("--bogus").returncode == 1
    v = run("--version")
    assert v.returncode == 0 and "0.1.6-git" in v.stdout
    h = run("--help")
    assert h.returncode == 0 and "--brute-window" in h.stdout and "--shortcut-len" in h.stdout


def test_reconstruct_rejects_bad_atz():
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "x.atz")
        open(f, "wb").write(b"NOPE" + bytes(40))
        p = run("-r", "-i", f)
        assert p.returncode == 255 and "Invalid file: ATZ1 header not found" in p.stdout          # main.cpp:1018-1021, main() returns -1
        open(f, "wb").write(b"ATZ\x01" + struct.pack("<QQQ", 999, 10, 0) + b"0123456789")
        p = run("-r", "-i", f)
        assert p.returncode == 255 and "Invalid file: ATZ file size mismatch" in p.stdout        # main.cpp:1022-1025
        # an ATZ with no recompressed streams is a plain copy from offset 28 (main.cpp:941-948); needs no GPU
        body = b"hello world, no streams here"
        open(f, "wb").write(b"ATZ\x01" + struct.pack("<QQQ", 28 + len(body), len(body), 0) + body)
        p = run("-r", "-i", f, "-o", f + ".rec")
        assert p.returncode == 0 and open(f + ".rec", "rb").read() == body
        assert "reconstructing from" in p.stdout and "Original file size: %d" % len(body) in p.stdout
