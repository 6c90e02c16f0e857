"""A/B the gather build variants on the GPU box: per-stage ms/step of a short bench for each library
(default build + every build_variants/*.so, selected through RFB200_LIB)."""
import glob
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [None] + sorted(glob.glob(os.path.join(root, "build_variants", "*.so")))
extra = sys.argv[1:] or ["--steps", "3", "--warmup", "3"]
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["RFB200_LIB"] = lib
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline", "--no-e2e"] + extra,
                         env=env, capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not lines:
        print(os.path.basename(lib or "default"), "FAILED", out.stderr[-300:])
        continue
    d = json.loads(lines[-1])
    print("%-18s value=%8.0f ms/step=%7.2f %s" % (os.path.basename(lib or "default"), d["value"], d["ms_per_step"],
                                                 {k: round(v, 2) for k, v in d["config"]["stage_ms_per_step"].items()}), flush=True)
