"""File -> file timing of the `gotranseq` CLI (main.go:52-64: os.Open / os.Create around Translate)
on tmpfs, next to the restated reference's CLI; prints one JSON line."""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402

shm = "/dev/shm"
src, dst, ref = (os.path.join(shm, x) for x in ("gts_in.fna", "gts_out.faa", "gts_ref.faa"))
s = workloads.readme_stream()
with open(src, "wb") as f:
    f.write(s.generate().tobytes())
nt = s.nucleotides()
exe = os.path.join(ROOT, "gotranseq_b200", "bin", "gotranseq")
res = {"in_bytes": os.path.getsize(src), "nucleotides": nt, "runs": []}
for threads in ("1", "4", "8"):
    env = dict(os.environ, GOTRANSEQ_B200_IO_THREADS=threads, GOTRANSEQ_B200_TIMING="1")
    best, phases = None, None
    for _ in range(3):
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-s", src, "-o", dst, "-f", "6"], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, env=env)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best = dt
            try:
                phases = json.loads(r.stderr.decode().strip().splitlines()[-1])
            except Exception:
                phases = None
    stream_s = phases["stream_ms"] / 1e3 if phases else None
    res["runs"].append({"io_threads": int(threads), "wall_s": best, "Gbp_per_s_wall": nt / best / 1e9, "phases_ms": phases,
                        "Gbp_per_s_read_translate_write": nt / stream_s / 1e9 if stream_s else None})
res["out_bytes"] = os.path.getsize(dst)
res["out_sha256"] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
oc = os.path.join(ROOT, "oracle", "transeq_oracle_cli")
t0 = time.perf_counter()
subprocess.run([oc, "-s", src, "-o", ref, "-f", "6", "-n", "1"], check=True)
res["oracle_cli_1_worker_s"] = time.perf_counter() - t0
res["matches_oracle_cli"] = hashlib.sha256(open(ref, "rb").read()).hexdigest() == res["out_sha256"]
t0 = time.perf_counter()
subprocess.run([oc, "-s", src, "-o", ref, "-f", "6"], check=True)
res["oracle_cli_all_cores_s"] = time.perf_counter() - t0
res["cores"] = os.cpu_count()
print(json.dumps(res))
for p in (src, dst, ref):
    os.remove(p)
