"""The `uncomp` host program's argument handling and ATZ1 reader errors (no GPU needed for these paths)."""
import os
import struct
import subprocess
import tempfile

import pytest

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


REF = os.path.join(ROOT, "oracle", "_ref", "uncomp_ref")
CLI_CASES = ["", "--bogus", "--version", "-i", "--input", "-r", "-i x --recomp-tresh abc", "-i x --recomp-tresh 12abc", "-i x --chunksize", "-i x -i y",
             "-i x extra", "-i x --notest --notest", "-i nosuchfile -- foo bar", "-o y", "-i x -o", "--shortcut-len 5 --shortcut-len 6 -i x"]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/uncomp_ref not built")
def test_cli_texts_equal_the_reference_binary():
    """usage, help, version and parse-error texts (stdout, stderr, exit code) against the reference's TCLAP front end, both
    programs started under the same argv[0] (the usage line is wrapped relative to the program name)"""
    with tempfile.TemporaryDirectory() as tmp:
        for d, exe in (("a", REF), ("b", UNCOMP)):
            os.mkdir(os.path.join(tmp, d)); os.symlink(exe, os.path.join(tmp, d, "uncomp"))
        def both(args):
            return [subprocess.run(["./uncomp"] + args.split(), cwd=os.path.join(tmp, d), capture_output=True, text=True) for d in ("a", "b")]
        for args in CLI_CASES:
            r, u = both(args)
            assert (r.returncode, r.stdout, r.stderr) == (u.returncode, u.stdout, u.stderr), args
        # --help: the reference's text, then the extensions of this implementation
        r, u = both("--help")
        assert r.returncode == u.returncode == 0 and u.stdout.startswith(r.stdout) and r.stderr == u.stderr
        rest = u.stdout[len(r.stdout):]
        assert rest.startswith("antiz_b200 extensions: ") and all(f in rest for f in ("--gpus <integer>", "--device <integer>", "--exact-records", "--stats"))


def test_cli_layout_without_the_reference():
    """the same layout rules, pinned on literal text (runs where oracle/_ref is absent)"""
    p = run("--bogus")
    assert p.returncode == 1 and p.stdout == "AntiZ 0.1.6-git\n"
    lines = p.stderr.split("\n")
    assert lines[0] == "PARSE ERROR: Argument: --bogus" and lines[1] == "             Couldn't find match for argument" and lines[3] == "Brief USAGE: "
    assert all(len(l) <= 75 for l in lines) and lines[4].startswith("   " + UNCOMP + "  [--brute-window] [--notest] [-r]")
    assert p.stderr.endswith("For complete USAGE and HELP type: \n   " + UNCOMP + " --help\n\n")
    assert run().stderr.startswith("PARSE ERROR:  \n             Required argument missing: input\n")
    assert run("-i").stderr.startswith("PARSE ERROR: Argument: -i (--input)\n             Missing a value for this argument!\n")
    assert run("-i", "x", "--mismatch-tol", "q").stderr.startswith("PARSE ERROR: Argument: (--mismatch-tol)\n             Couldn't read argument value from string 'q'\n")
    h = run("--help").stdout
    assert h.startswith("AntiZ 0.1.6-git\n\nUSAGE: \n\n   ") and "\n\nWhere: \n\n   --brute-window\n     Bruteforce deflate window size" in h
    assert "   -i <string>,  --input <string>\n     (required)  Input file name\n" in h and all(len(l) <= 75 for l in h.split("\n"))
