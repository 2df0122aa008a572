"""Full-size parity of the BASELINE configs: the product's `uncomp` against the unmodified reference binary
(oracle/_ref/uncomp_ref) on the same container and flags - .atz sha256, stream / recompressed counts, wall times - and the round trip
back through `uncomp -r`.  Runs on a GPU box; writes one JSON line per config.   usage: python tools/fullsize_parity.py [c1 c2 c3 c4 c5:64 ...]"""
import hashlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, zref

def sha(p):
    h = hashlib.sha256()
    with open(p, "rb") as f:
        for b in iter(lambda: f.read(1 << 22), b""):
            h.update(b)
    return h.hexdigest()

def run(cmd):
    t = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True); return r, time.perf_counter() - t

def main():
    uncomp = os.path.join(ROOT, "antiz_b200", "uncomp")
    for spec in sys.argv[1:] or ["c1", "c2", "c3", "c4"]:
        kind, _, n = spec.partition(":")
        desc, ndef, flags, _ = bench.WORKLOADS[kind]
        n = int(n) if n else ndef
        path, data, parts = bench.shared_container(kind, n, bench.SEED, leader=True)
        gpus = int(os.environ.get("PARITY_GPUS", "1"))
        ref, t_ref = run([zref.REF_BIN, "-i", path, "-o", path + ".ref.atz", "--notest"] + flags)
        gpu, t_gpu = run([uncomp, "-i", path, "-o", path + ".gpu.atz", "--notest", "--stats", "--gpus", str(gpus)] + flags)
        rec, t_rec = run([uncomp, "-r", "-i", path + ".gpu.atz", "-o", path + ".rec"])
        refrec, t_refrec = run([zref.REF_BIN, "-r", "-i", path + ".ref.atz", "-o", path + ".refrec"])
        keep = ("Total zlib headers found", "recompressed:", "Total bytes written")
        line = {"config": desc, "streams_or_mb": n, "container_bytes": len(data), "flags": flags, "gpus": gpus,
                "reference_stdout": [l for l in ref.stdout.splitlines() if l.startswith(keep)], "gpu_stdout": [l for l in gpu.stdout.splitlines() if l.startswith(keep)],
                "rc": [ref.returncode, gpu.returncode, rec.returncode],
                "atz_sha256_reference": sha(path + ".ref.atz") if ref.returncode == 0 else None, "atz_sha256_gpu": sha(path + ".gpu.atz") if gpu.returncode == 0 else None,
                "round_trip_ok": rec.returncode == 0 and sha(path + ".rec") == sha(path),
                "seconds": {"uncomp_ref --notest (1 thread)": round(t_ref, 2), "uncomp --notest (wall, incl. context creation)": round(t_gpu, 2),
                            "uncomp -r": round(t_rec, 2), "uncomp_ref -r": round(t_refrec, 2)},
                "MB_per_s": {"reference": round(len(data) / t_ref / 1e6, 2), "gpu_cli": round(len(data) / t_gpu / 1e6, 2),
                             "reconstruct_gpu_cli_out": round(len(data) / t_rec / 1e6, 2), "reconstruct_reference_out": round(len(data) / t_refrec / 1e6, 2)},
                "gpu_stats": [l for l in gpu.stderr.splitlines() if l.startswith("[")][:12]}
        line["identical"] = line["atz_sha256_reference"] is not None and line["atz_sha256_reference"] == line["atz_sha256_gpu"]
        print(json.dumps(line), flush=True)
        for ext in (".ref.atz", ".gpu.atz", ".rec", ".refrec"):
            try:
                os.unlink(path + ext)
            except OSError:
                pass

if __name__ == "__main__":
    main()
